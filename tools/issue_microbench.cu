// issue_microbench.cu — what bounds a kernel that mixes FP64-pipe instructions (DFMA/DADD/DMUL/DSETP) with
// 32-bit ALU instructions (FSEL, IMAD.MOV, FFMA) on sm_100a? Each kernel runs `iters` iterations of a straight-line
// body of independent chains; 8 warps per SM sub-partition. Per-iteration instruction counts come from the SASS
// (`cuobjdump -sass`, the loop body is not unrolled across iterations); the program prints time per iteration in
// cycles per warp-iteration per sub-partition so that rates follow by division.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_microbench tools/issue_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

// KIND 0: 8 DFMA                         1: 8 DADD                      2: 8 DMUL
//      3: 8 DSETP + 16 FSEL (select between doubles)                    4: 8 DFMA + (8 DSETP + 16 FSEL)
//      5: 8 DFMA + 16 FFMA               6: 8 DFMA + 8 DSETP + 16 FSEL + 16 FFMA
//      7: 8 DADD + 8 DMUL + 8 DSETP + 16 FSEL + 16 FFMA (the model kernel's mix, roughly)
template <int KIND>
__global__ void __launch_bounds__(256, 4) mix(double* out, int iters, double seed) {
    double d[8], e[8];
    float f[16];
    int n[8];
    const int hi = (int)seed + 1073741824;
    for (int k = 0; k < 8; ++k) n[k] = k;
    const double m = seed * 1e-9 + 1.0, c = seed * 1e-12, thr = seed * 3.0;
    for (int k = 0; k < 8; ++k) {
        d[k] = seed + k + threadIdx.x;
        e[k] = seed * 2 + k;
    }
    for (int k = 0; k < 16; ++k) f[k] = (float)seed + k;
    const float fm = (float)m, fc = (float)c;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0 || KIND == 4 || KIND == 5 || KIND == 6) d[k] = fma(d[k], m, c);
            if (KIND == 1 || KIND == 7) d[k] = __dadd_rn(d[k], c);
            if (KIND == 2) d[k] = __dmul_rn(d[k], m);
            if (KIND == 7) e[k] = __dmul_rn(e[k], m);
            if (KIND == 3 || KIND == 4 || KIND == 6) e[k] = (d[k] > thr) ? e[k] : d[(k + 1) & 7];
            if (KIND == 7) e[k] = (d[k] > thr) ? e[k] : c;
            if (KIND == 3) d[k] = (e[k] > thr) ? d[k] : e[(k + 3) & 7];
            if (KIND >= 8 && KIND <= 12) d[k] = __dadd_rn(d[k], c);
            if (KIND == 8) n[k] += (d[k] > thr) ? 1 : 0;                       // DSETP + predicated integer add
            if (KIND == 9) e[k] = (d[k] > thr) ? e[k] : c;                     // DSETP + 2 FSEL
            if (KIND == 10) e[k] = fmax(e[k], d[k]);                           // double max
            if (KIND == 11) n[k] += (__double2hiint(d[k]) > hi) ? 1 : 0;       // integer compare of the high word
            if (KIND == 12) e[k] = (__double2hiint(d[k]) > hi) ? e[k] : c;     // ISETP + 2 FSEL
            if (KIND == 5 || KIND == 6 || KIND == 7) {
                f[2 * k] = fmaf(f[2 * k], fm, fc);
                f[2 * k + 1] = fmaf(f[2 * k + 1], fm, fc);
            }
        }
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += d[k] + e[k];
    for (int k = 0; k < 16; ++k) s += f[k];
    for (int k = 0; k < 8; ++k) s += n[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double* out, int sms, double clockGHz) {
    const int iters = 20000, blocks = sms * 4, threads = 256;
    mix<KIND><<<blocks, threads>>>(out, 100, 1.5);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    mix<KIND><<<blocks, threads>>>(out, iters, 1.5);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warpsPerSmsp = blocks * (threads / 32.0) / (sms * 4.0);
    const double cycles = ms * 1e-3 * clockGHz * 1e9;
    printf("kind %d %-46s %8.3f ms  %.2f cycles per warp-iteration per sub-partition\n", KIND, name, ms,
           cycles / (iters * warpsPerSmsp));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal max; cycle counts assume the SM runs at it)\n", p.name, p.multiProcessorCount, ghz);
    double* out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 256);
    const int n = p.multiProcessorCount;
    run<0>("8 DFMA", out, n, ghz);
    run<1>("8 DADD", out, n, ghz);
    run<2>("8 DMUL", out, n, ghz);
    run<3>("16 x (DSETP + 2 FSEL)", out, n, ghz);
    run<4>("8 DFMA + 8 x (DSETP + 2 FSEL)", out, n, ghz);
    run<5>("8 DFMA + 16 FFMA", out, n, ghz);
    run<6>("8 DFMA + 8 x (DSETP + 2 FSEL) + 16 FFMA", out, n, ghz);
    run<7>("8 DADD + 8 DMUL + 8 x (DSETP + 2 FSEL) + 16 FFMA", out, n, ghz);
    run<8>("8 DADD + 8 x (DSETP + int add)", out, n, ghz);
    run<9>("8 DADD + 8 x (DSETP + 2 FSEL)", out, n, ghz);
    run<10>("8 DADD + 8 fmax(double)", out, n, ghz);
    run<11>("8 DADD + 8 x (ISETP hi word + int add)", out, n, ghz);
    run<12>("8 DADD + 8 x (ISETP hi word + 2 FSEL)", out, n, ghz);
    cudaFree(out);
    return 0;
}
