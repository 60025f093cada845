// issue_microbench2.cu — isolates what the 32-bit ALU consumers of FP64 results cost on sm_100a (see
// issue_microbench.cu). 8 independent chains per thread, 8 warps per sub-partition, loop body not unrolled.
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256, 4) mix(double* out, int iters, double seed) {
    double d[8], e[8];
    const double m = seed * 1e-9 + 1.0, c = seed * 1e-12, thr = seed * 3.0;
    for (int k = 0; k < 8; ++k) {
        d[k] = seed + k + threadIdx.x;
        e[k] = seed * 2 + k;
    }
    int flip = (int)seed;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        flip = flip * 1103515245 + 12345;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            d[k] = fma(d[k], m, c);
            const bool pi = (flip >> k) & 1;  // predicate from the integer pipe
            if (KIND == 1) e[k] = pi ? e[k] : e[(k + 1) & 7];          // FSEL on old data
            if (KIND == 2) e[k] = pi ? e[k] : d[k];                    // FSEL reads the fresh DFMA result
            if (KIND == 3) e[k] = (d[k] > thr) ? e[k] : e[(k + 1) & 7];  // DSETP on fresh, FSEL on old data
            if (KIND == 4) e[k] = (e[(k + 3) & 7] > thr) ? e[k] : e[(k + 1) & 7];  // DSETP on old data
            if (KIND == 5) e[k] = (d[k] > thr) ? e[k] : d[k];          // DSETP on fresh + FSEL of fresh
            if (KIND == 6) e[k] = (d[(k + 4) & 7] > thr) ? e[k] : d[(k + 4) & 7];  // same, consumer 4 chains away
        }
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += d[k] + e[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + flip;
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
    printf("kind %d %-52s %8.3f ms  %.2f cycles per warp-iteration per sub-partition\n", KIND, name, ms,
           ms * 1e-3 * clockGHz * 1e9 / (iters * warpsPerSmsp));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    double* out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 256);
    const int n = p.multiProcessorCount;
    run<0>("8 DFMA", out, n, ghz);
    run<1>("8 DFMA + 8 double selects of old data, int predicate", out, n, ghz);
    run<2>("8 DFMA + 8 double selects of the fresh result, int pred", out, n, ghz);
    run<3>("8 DFMA + 8 DSETP(fresh) + selects of old data", out, n, ghz);
    run<4>("8 DFMA + 8 DSETP(old) + selects of old data", out, n, ghz);
    run<5>("8 DFMA + 8 DSETP(fresh) + selects of fresh", out, n, ghz);
    run<6>("8 DFMA + 8 DSETP(fresh, 4 away) + selects of fresh", out, n, ghz);
    cudaFree(out);
    return 0;
}
