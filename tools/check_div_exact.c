#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
static uint64_t s=88172645463325252ULL;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
static inline double u01(){return (rnd()>>11)*(1.0/9007199254740992.0);}
int main(){
  long bad=0,n=0,badq=0;
  for(long i=0;i<400000000L;i++){
    double a,b;
    int mode=i&3;
    if(mode==0){ a=u01()*5000; b=(double)(float)(u01()*3000); }
    else if(mode==1){ a=ldexp(1+u01(), (int)(rnd()%80)-40); b=ldexp(1+u01(), (int)(rnd()%80)-40); }
    else if(mode==2){ a=u01()*1e4; b=6.0; }
    else { a=u01()*100; b=u01()*100+1e-9; }
    if(b==0) continue;
    double inv=1.0/b;
    double q=a*inv;
    double r=fma(-q,b,a);
    double q2=fma(r,inv,q);
    double t=a/b;
    n++;
    if(q!=t) badq++;
    if(q2!=t){bad++; if(bad<5) printf("a=%a b=%a q2=%a t=%a\n",a,b,q2,t);}
  }
  printf("n=%ld uncorrected-differs=%ld corrected-differs=%ld\n",n,badq,bad);
}
