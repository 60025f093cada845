// tests/host_emul/hbk_host_shims.h — lets g++ compile hydrobricks_b200/csrc/hbk_device.cuh for the CPU.
// TEST INFRASTRUCTURE ONLY (-DHBK_HOST_EMULATION): the per-(parameter set, hydro unit) arithmetic of the CUDA kernels,
// the very same source, run on the host so that the CPU test suite can check it against the reference's golden vectors.
// Differences to the device build: libm exp / log2f / exp2f instead of the CUDA versions (<= a few ulp, and the fp32
// seed of pow53 is Newton-corrected anyway); everything else (IEEE add/mul/div/sqrt/fma, -ffp-contract=off = -fmad=false)
// is bit-identical.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define HBK_NOINLINE __attribute__((noinline))

static inline int __double2hiint(double x) {
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (int)(b >> 32);
}
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
