// hbk_engine.cu — B200 (sm_100a) batched simulation engine behind the C-ABI of include/hbk.h.
//
// Replaces ModelHydro::Run() (reference core/src/base/ModelHydro.cpp:100-133) for Socont-family
// structures: ONE launch integrates every (parameter set, hydro unit) pair over a whole time segment.
//
//  * one thread per (parameter set p, hydro unit u); storages live in registers, the time loop is inside
//    the kernel (hbk_device.cuh holds the per-step arithmetic);
//  * a block covers PB parameter sets x UB units. PB = 32 puts the parameter axis on the lanes (calibration
//    ensembles: every lane of a warp reads the same forcing value -> shared-memory broadcast);
//    PB = 1 puts the unit axis on the lanes (large distributed basins);
//  * forcing is staged per block in chunks of TC steps by TMA bulk copies (cp.async.bulk + mbarrier),
//    double buffered: either the block's tile of pre-spatialised per-unit series, or the station series
//    with the elevation lapse rates applied in registers (fused spatialisation, forcing.py:1061-1113);
//  * the per-step outlet contributions (and the glacier deposits into the two sub-basin stores) are kept
//    in shared memory for a chunk and reduced over the block's units in a fixed order, then written once;
//  * the two sub-basin linear stores (glacier_area_*_storage) only depend on the per-step deposit sums, so
//    a second tiny kernel integrates them per parameter set after the unit pass;
//  * state is spilled to a struct-of-arrays in HBM at segment boundaries, which is where dated actions
//    (glacier evolution, land-cover change) get applied by their own kernels.
//
// There is no CPU fallback anywhere in this file.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hbk.h"
#include "hbk_device.cuh"

namespace hbk {

#ifndef HBK_ENS_TC
#define HBK_ENS_TC 16  // steps per chunk of the compile-time ensemble shape
#endif
#ifndef HBK_MAX_THREADS
#define HBK_MAX_THREADS 128  // largest block the host may launch (PB * UB); 128 x 6 blocks = 24 warps/SM measured best
#endif
#ifndef HBK_MIN_BLOCKS
#define HBK_MIN_BLOCKS 6  // register budget = 65536 / (HBK_MAX_THREADS * HBK_MIN_BLOCKS)
#endif
#ifndef HBK_MIN_BLOCKS_SOIL2
// two soil stores without glacier: at 6 blocks (80 registers) the spill frames (312 B x 768 threads) overflow the L1 left
// beside the shared memory; 5 blocks (96 registers, 240 B) measured +4 % on C4's plain tiles, 4 blocks no better
#define HBK_MIN_BLOCKS_SOIL2 5
#endif
#ifndef HBK_ENS_TC_GLACIER
// glacier variants: three outlet quantities per step in the chunk slots. 16 steps x 4 blocks = 200 KB of shared memory
// leave 23 KB of L1 for 143 KB of spill frames (ncu r02s: 5 % local-load hit rate); 8 steps measured +4 % on C3 and C4
#define HBK_ENS_TC_GLACIER 8
#endif
#ifndef HBK_MIN_BLOCKS_GLACIER
#define HBK_MIN_BLOCKS_GLACIER 4  // the glacier variants carry ~60 more live registers: 128 regs x 16 warps (C3: +8.5 % over 6)
#endif
constexpr int kBlock = HBK_MAX_THREADS;
constexpr int kNH = 4;  // per-unit constants: area fraction of basin, area [m2], pow(slope,0.5), elevation

struct KArgs {
    int P, U, tBegin, tEnd, PB, UB, TC, nUTiles, G, nGroups;
    // the unit tiles [tile0, tile1) of this launch, cut into nGroups groups of G tiles; grpBase = index of its first group
    // among the partial buffers (one launch covers all tiles unless the tiles without glacier units run the plain kernel)
    int tile0, tile1, grpBase;
    int clusterMode;  // 1: one thread-block cluster per parameter-set group, one CTA per unit tile, DSMEM reduction
    Flags flags;
    double dt;
    const float* params;  // [HBK_N_PARAMS][ldp]
    int ldp;
    const double* hru;  // [kNH][ldu]
    const int* hruFlags;
    int ldu;
    const double* frac;  // [2][fracLd]: ground, glacier
    long long fracLd;
    int fracPerSet;
    int forcingMode;            // 0: tiled per-unit series, 1: station series + fused spatialisation
    const double* forcing[HBK_N_VARS];  // mode 0: [nUTiles][tld][UB]; mode 1: [tld]
    int nVars;                  // 3, or 4 with the solar radiation of melt:temperature_index
    long long tld;
    const double* gradT;  // per set or nullptr
    const double* gradP;
    const double* corrP;
    double gradTs, gradPs, corrPs, zrefT, zrefP;
    double* state;  // [HBK_N_STATES][lds], element (p, u) at p*sP + u*sU
    long long lds, sP, sU;
    double* outQ;  // [nUTiles][P][ldt]
    double* outD1;
    double* outD2;
    long long ldt;
    int writeOutputs;
    unsigned long long recordMask;
    double* rec;  // [popcount][P][T][U]
    int T;
    int* err;
    unsigned long long* unconverged;
    int* unconvergedPerSet;  // [P], internal set order: sweeps of that set's units that did not stabilise
    // work queue (persistent blocks): items = (time slice, virtual block), slice-major; nullptr = one block per virtual
    // block over the whole segment. progress[vb] = slices of that virtual block that are complete.
    int* workCounter;
    int* progress;
    int nVBlocks, nItems, sliceChunks;
    unsigned long long* blockTimes;  // tuning aid (HBK_BLOCK_TIMES): [grid][3] start ns, end ns, SM id; else nullptr
    // Parameter sets are stored in an engine-chosen order (similar sets share a warp, hbk_set_parameters);
    // perm[internal index] = caller's index, nullptr = identity. Only what leaves the engine is mapped back:
    // final outlet rows, recorder rows, per-set spatialisation inputs.
    const int* perm;
    int outQFinal;  // outQ is the caller-visible outlet (no tile partials)
    const double* swatFactor;  // [T]: 1 + sin(2 pi (day of year - ref) / 365), transform:snow_ice_swat; else nullptr
};

// ---- TMA bulk copy + mbarrier (PTX) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP). 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tmaLoad1D(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

// Snowpack sublimation rates of one step (sublimation:constant: the parameter; sublimation:pet: PET x the parameter,
// float32 promoted; ProcessSublimation{Constant,PET}.cpp). Out of line and read from global memory at use, so that the
// kernels of structures without the option carry neither the registers nor predicated instructions for it.
template <bool GLACIER>
__device__ __noinline__ double2 sublimationRates(const float* params, int ldp, int kind, int p, double pet, bool twin = false) {
    const double scale = (kind == HBK_SUBLIMATION_PET) ? pet : 1.0;
    double2 r;
    r.x = scale * (double)params[p + (size_t)HBK_P_GROUND_SUBLIMATION * ldp];
    r.y = GLACIER ? scale * (double)params[p + (size_t)(twin ? HBK_P_GLACIER2_SUBLIMATION : HBK_P_GLACIER_SUBLIMATION) * ldp]
                  : 0.0;
    return r;
}

// The three melt factors of one (set, unit) under melt:degree_day_aspect (aspect: HBK_ASPECT_*) resp. of one step under
// melt:temperature_index. Out of line and read from global memory at use like the sublimation rates: structures with the
// plain melt:degree_day carry neither registers nor instructions for them in the time loop.
struct AspectFactors {
    double ground, glacierSnow, ice;
};
template <bool GLACIER>
__device__ __noinline__ AspectFactors aspectFactors(const float* params, int ldp, int p, int aspect, bool twin = false) {
    const int g = aspect == 0 ? HBK_P_GROUND_SNOW_DDF : (aspect == 1 ? HBK_P_GROUND_SNOW_MELT_B : HBK_P_GROUND_SNOW_MELT_C);
    int s = aspect == 0 ? HBK_P_GLACIER_SNOW_DDF : (aspect == 1 ? HBK_P_GLACIER_SNOW_MELT_B : HBK_P_GLACIER_SNOW_MELT_C);
    int i = aspect == 0 ? HBK_P_ICE_DDF : (aspect == 1 ? HBK_P_ICE_MELT_B : HBK_P_ICE_MELT_C);
    if (twin) {  // the second glacier cover's own factors
        s = aspect == 0 ? HBK_P_GLACIER2_SNOW_DDF : (aspect == 1 ? HBK_P_GLACIER2_SNOW_MELT_B : HBK_P_GLACIER2_SNOW_MELT_C);
        i = aspect == 0 ? HBK_P_ICE2_DDF : (aspect == 1 ? HBK_P_ICE2_MELT_B : HBK_P_ICE2_MELT_C);
    }
    AspectFactors f;
    f.ground = (double)params[p + (size_t)g * ldp];
    f.glacierSnow = GLACIER ? (double)params[p + (size_t)s * ldp] : 0.0;
    f.ice = GLACIER ? (double)params[p + (size_t)i * ldp] : 0.0;
    return f;
}
template <bool GLACIER>
__device__ __noinline__ AspectFactors radiationFactors(const float* params, int ldp, int p, double radiation,
                                                       bool twin = false) {
    AspectFactors f;  // float32 parameters promoted, the product and the sum in double like the reference's expression
    f.ground = (double)params[p + (size_t)HBK_P_GROUND_SNOW_DDF * ldp] +
               (double)params[p + (size_t)HBK_P_GROUND_SNOW_MELT_B * ldp] * radiation;
    const int sA = twin ? HBK_P_GLACIER2_SNOW_DDF : HBK_P_GLACIER_SNOW_DDF, sB = twin ? HBK_P_GLACIER2_SNOW_MELT_B : HBK_P_GLACIER_SNOW_MELT_B;
    const int iA = twin ? HBK_P_ICE2_DDF : HBK_P_ICE_DDF, iB = twin ? HBK_P_ICE2_MELT_B : HBK_P_ICE_MELT_B;
    f.glacierSnow = GLACIER ? (double)params[p + (size_t)sA * ldp] + (double)params[p + (size_t)sB * ldp] * radiation : 0.0;
    f.ice = GLACIER ? (double)params[p + (size_t)iA * ldp] + (double)params[p + (size_t)iB * ldp] * radiation : 0.0;
    return f;
}

// The glacier parameters of the unit being loaded when the structure has twin units (hbk_set_unit_twins): an ordinary
// unit takes the first cover's slots, a twin the second cover's. Out of line, read from global memory at use.
struct GlacierPars {
    double snowDdf, snowTm, iceDdf, iceTm, snowIce;
};
__device__ __noinline__ GlacierPars glacierPars(const float* params, int ldp, int p, bool twin) {
    GlacierPars g;
    g.snowDdf = (double)params[p + (size_t)(twin ? HBK_P_GLACIER2_SNOW_DDF : HBK_P_GLACIER_SNOW_DDF) * ldp];
    g.snowTm = (double)params[p + (size_t)(twin ? HBK_P_GLACIER2_SNOW_TMELT : HBK_P_GLACIER_SNOW_TMELT) * ldp];
    g.iceDdf = (double)params[p + (size_t)(twin ? HBK_P_ICE2_DDF : HBK_P_ICE_DDF) * ldp];
    g.iceTm = (double)params[p + (size_t)(twin ? HBK_P_ICE2_TMELT : HBK_P_ICE_TMELT) * ldp];
    g.snowIce = (double)params[p + (size_t)(twin ? HBK_P_SNOW_ICE2 : HBK_P_SNOW_ICE) * ldp];
    return g;
}

// Per-parameter-set values of one set (pp = its column of the float32 parameter table), promoted to double where the
// reference's expressions promote them; pv[field * ps].
__device__ __forceinline__ void fillPar(double* pv, int ps, const float* pp, int ldp) {
    const float t0f = pp[HBK_P_TRANSITION_START * ldp], t1f = pp[HBK_P_TRANSITION_END * ldp];
    pv[0 * ps] = t0f;
    pv[1 * ps] = t1f;
    pv[2 * ps] = 1.0 / (double)(t1f - t0f);
    pv[3 * ps] = pp[HBK_P_RAIN_CORRECTION * ldp];
    pv[4 * ps] = pp[HBK_P_SNOW_CORRECTION * ldp];
    pv[5 * ps] = pp[HBK_P_GROUND_SNOW_DDF * ldp];
    pv[6 * ps] = pp[HBK_P_GROUND_SNOW_TMELT * ldp];
    pv[7 * ps] = pp[HBK_P_GLACIER_SNOW_DDF * ldp];
    pv[8 * ps] = pp[HBK_P_GLACIER_SNOW_TMELT * ldp];
    pv[9 * ps] = pp[HBK_P_ICE_DDF * ldp];
    pv[10 * ps] = pp[HBK_P_ICE_TMELT * ldp];
    const double capacity = pp[HBK_P_CAPACITY * ldp];
    pv[11 * ps] = capacity;
    pv[12 * ps] = (capacity == 0.0) ? kInvZeroCapacity : 1.0 / capacity;
    pv[13 * ps] = pp[HBK_P_K_SLOW_1 * ldp];
    pv[14 * ps] = pp[HBK_P_PERCOLATION * ldp];
    pv[15 * ps] = pp[HBK_P_K_SLOW_2 * ldp];
    pv[16 * ps] = pp[HBK_P_SNOW_ICE * ldp];
}

// HBK_RETENTION_*: where the extra state of one (set, unit) lives and its four parameters (read at use, like the options above)
__device__ __noinline__ RetentionCtx retentionCtx(double* state, long long lds, long long si, const float* params, int ldp,
                                                  int p, int bits) {
    RetentionCtx c;
    const float* pp = params + p;
    c.g.water = state + HBK_S_GROUND_SNOWPACK_WATER * lds + si;
    c.g.refreeze = state + HBK_S_AMT_REFREEZE_GROUND * lds + si;
    c.g.snowmelt = state + HBK_S_AMT_SNOWMELT_GROUND * lds + si;
    c.g.whc = pp[(size_t)HBK_P_GROUND_SNOW_WHC * ldp];
    c.g.cfr = pp[(size_t)HBK_P_GROUND_SNOW_CFR * ldp];
    c.gl.water = state + HBK_S_GLACIER_SNOWPACK_WATER * lds + si;
    c.gl.refreeze = state + HBK_S_AMT_REFREEZE_GLACIER * lds + si;
    c.gl.snowmelt = state + HBK_S_AMT_SNOWMELT_GLACIER * lds + si;
    c.gl.whc = pp[(size_t)HBK_P_GLACIER_SNOW_WHC * ldp];
    c.gl.cfr = pp[(size_t)HBK_P_GLACIER_SNOW_CFR * ldp];
    c.refreeze = (bits & HBK_RETENTION_REFREEZE) != 0;
    c.rainToSnowpack = (bits & HBK_RETENTION_RAIN_TO_SNOWPACK) != 0;
    return c;
}

// Logger::Record (Logger.cpp:93-120) of one (set, unit, step): contents after Finalize and flux amounts, into the slots
// of the record mask [slot][P][T][U]; NaN where the unit's structure variant has no such brick.
__device__ __forceinline__ void recordStep(const KArgs& a, int pOut, int t, int u, const Hru& h, const StepOut& o, double fG,
                                           double fGl, bool gv, const RetentionCtx* ret = nullptr) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double* base = a.rec + ((size_t)pOut * a.T + t) * a.U + u;
    const size_t stride = (size_t)a.P * a.T * a.U;
    int k = 0;
#define HBK_REC(slot, value)                                   \
    if (a.recordMask & (1ull << (slot))) {                     \
        base[(size_t)k * stride] = (value);                    \
        ++k;                                                   \
    }
    HBK_REC(HBK_R_GROUND_WATER, h.wG)
    HBK_REC(HBK_R_GROUND_INFILTRATION, h.amtInfil)
    HBK_REC(HBK_R_GROUND_RUNOFF, h.amtRest)
    HBK_REC(HBK_R_GLACIER_WATER, gv ? h.wGl : nan)
    HBK_REC(HBK_R_GLACIER_ICE, gv ? h.ice : nan)
    HBK_REC(HBK_R_GLACIER_OUTFLOW, gv ? h.amtDep1 : nan)
    HBK_REC(HBK_R_GLACIER_MELT, gv ? h.amtDep2 : nan)
    // with liquid water in the snowpacks h.amtMelt* are the melt WATER released to the covers; content, melt and refreeze
    // amounts are in the state arrays (RetentionCtx)
    HBK_REC(HBK_R_GROUND_SNOWPACK_WATER, ret ? *ret->g.water : 0.0)
    HBK_REC(HBK_R_GROUND_SNOWPACK_SNOW, h.snowG)
    HBK_REC(HBK_R_GROUND_SNOWPACK_MELT, ret ? *ret->g.snowmelt : h.amtMeltG)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_WATER, gv ? (ret ? *ret->gl.water : 0.0) : nan)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_SNOW, gv ? h.snowGl : nan)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_MELT, gv ? (ret ? *ret->gl.snowmelt : h.amtMeltGl) : nan)
    HBK_REC(HBK_R_SLOW_WATER, h.slow)
    HBK_REC(HBK_R_SLOW_ET, o.amtEt)
    HBK_REC(HBK_R_SLOW_OUTFLOW, o.amtOut)
    HBK_REC(HBK_R_SLOW_OVERFLOW, o.amtOvf)
    HBK_REC(HBK_R_SLOW_PERCOLATION, o.amtPerc)
    HBK_REC(HBK_R_SLOW2_WATER, h.slow2)
    HBK_REC(HBK_R_SLOW2_OUTFLOW, o.amtOut2)
    HBK_REC(HBK_R_QUICK_WATER, h.quick)
    HBK_REC(HBK_R_QUICK_RUNOFF, o.amtQ)
    HBK_REC(HBK_R_FRACTION_GROUND, fG)
    HBK_REC(HBK_R_FRACTION_GLACIER, gv ? fGl : nan)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_TRANSFO, gv ? h.amtSnowIce : nan)
    HBK_REC(HBK_R_GROUND_SNOWPACK_SUBLIMATION, o.amtSubG)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_SUBLIMATION, gv ? o.amtSubGl : nan)
    HBK_REC(HBK_R_GROUND_SNOWPACK_MELTWATER, h.amtMeltG)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_MELTWATER, gv ? h.amtMeltGl : nan)
    HBK_REC(HBK_R_GROUND_SNOWPACK_REFREEZE, ret ? *ret->g.refreeze : 0.0)
    HBK_REC(HBK_R_GLACIER_SNOWPACK_REFREEZE, gv ? (ret ? *ret->gl.refreeze : 0.0) : nan)
#undef HBK_REC
}

// resident blocks per multiprocessor a kernel variant is compiled for (= its register budget) and the chunk length of
// the compile-time ensemble shape
__host__ __device__ constexpr int hbkMinBlocks(int nSoil, bool glacier) {
    return glacier ? HBK_MIN_BLOCKS_GLACIER : (nSoil == 2 ? HBK_MIN_BLOCKS_SOIL2 : HBK_MIN_BLOCKS);
}
__host__ __device__ constexpr int hbkEnsTC(bool glacier) { return glacier ? HBK_ENS_TC_GLACIER : HBK_ENS_TC; }

// ---- main kernel ------------------------------------------------------------------------------------
// Block = PB parameter sets x UB units (threads), time-multiplexed over a GROUP of G consecutive unit tiles:
//   for each chunk of TC steps: for each tile of the group: [load state] TC steps [store state], add the tile's
//   outlet contributions to the chunk accumulator; after the last tile write outlet[p][t0..t0+TC).
// With G = all tiles (ensembles) one block sums every unit of its parameter sets in unit order: no partial
// buffers, no second pass. With G = 1 (few sets, many units) the state never leaves the registers.
// DAILY: the time step is exactly one day (the reference's usual case); x * 1.0 and x / 1.0 are exact, so the
// multiplications by dt fold away without changing a bit.
template <int SOLVER, int NSOIL, bool GLACIER, bool DAILY, int SHAPE>
__global__ void __launch_bounds__(HBK_MAX_THREADS, hbkMinBlocks(NSOIL, GLACIER))
    hbkRunUnits(const KArgs a) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    constexpr int NQ = GLACIER ? 3 : 1;
    // SHAPE 1 / 2: a block shape fixed at compile time — 1: calibration ensembles, 32 parameter sets (lanes) x 4 units;
    // 2: distributed basins, one parameter set x 128 units (lanes = units) — with the variant's chunk length, station
    // forcing with fused spatialisation, no cluster, so that the address arithmetic of the time loop folds into
    // immediates (at 80 registers ptxas otherwise re-derives it from the kernel arguments at every step). 0: any shape.
    constexpr bool ENS = SHAPE != 0;
    const int PB = SHAPE == 1 ? 32 : (SHAPE == 2 ? 1 : a.PB), UB = SHAPE == 1 ? 4 : (SHAPE == 2 ? 128 : a.UB);
    const int TC = ENS ? hbkEnsTC(GLACIER) : a.TC;
    const int forcingMode = ENS ? 1 : a.forcingMode;
    const int NV = ENS ? 3 : a.nVars;  // staged forcing variables (the ensemble shape never carries the radiation)
    const bool clusterMode = ENS ? false : (a.clusterMode != 0);
    const int tid = threadIdx.x;
    const int nThreads = ENS ? 128 : blockDim.x;  // PB * UB
    // cluster mode: the cluster's CTAs are the unit tiles of one parameter-set group; every CTA keeps its tile's
    // state in registers for the whole segment and the outlet is reduced through distributed shared memory.
    const int pl = tid % PB;
    const int ul = tid / PB;
    // identity of the (virtual) block being integrated: set by setBlock() — once for a plain launch, once per work item
    // when the blocks are persistent and pull (time slice, virtual block) items from the queue
    int grp = 0, pg = 0, crank = 0, p = 0, pOut = 0, tileBegin = 0, tileEnd = 0;
    bool validP = false, multi = false;

    // shared memory carve-up
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemRaw);  // 2 mbarriers
    const int fRow = (forcingMode == 0) ? TC * UB : (TC + 2);  // doubles per variable per stage
    double* fbuf = reinterpret_cast<double*>(smemRaw + 128);             // [2][NV][fRow]
    double* qbuf = fbuf + 2 * NV * fRow;                                 // [NQ][TC][nThreads]
    double* accb = qbuf + (size_t)NQ * TC * nThreads;                  // cluster mode only: [2][NQ][TC][PB]

    // per-parameter-set constants: shared memory [field][PB] (or registers, HBK_PAR_SHARED=0)
    [[maybe_unused]] double* sPar = accb + (size_t)(clusterMode ? 2 : 0) * NQ * TC * PB;
    Par par;
    double quickBase = 0, gT = 0, gP = 0, corr = 1;
    auto setBlock = [&](int vb) {
        // cluster mode: the cluster's CTAs are the unit tiles of one parameter-set group; every CTA keeps its tile's
        // state in registers for the whole segment and the outlet is reduced through distributed shared memory.
        grp = clusterMode ? 0 : vb % a.nGroups;
        pg = clusterMode ? vb / a.nUTiles : vb / a.nGroups;
        crank = clusterMode ? vb % a.nUTiles : 0;
        p = pg * PB + pl;
        validP = p < a.P;
        pOut = (validP && a.perm) ? a.perm[p] : p;  // the caller's index of this parameter set
        tileBegin = clusterMode ? crank : a.tile0 + grp * a.G;
        tileEnd = clusterMode ? crank + 1 : min(tileBegin + a.G, a.tile1);
        multi = !clusterMode && a.G > 1;
#if HBK_PAR_SHARED
        double* pv = sPar + pl;
        const int ps = PB;
        par.b = pv;
        par.s = ps;
        const bool fill = validP && ul == 0;
#else
        double* pv = par.v;
        const int ps = 1;
        const bool fill = validP;
#endif
        if (fill) fillPar(pv, ps, a.params + p, a.ldp);
        if (validP) {
            quickBase = a.params[p + (size_t)HBK_P_QUICK * a.ldp];
            par.quickV = quickBase;
            if (forcingMode == 1) {
                gT = a.gradT ? a.gradT[pOut] : a.gradTs;
                gP = a.gradP ? a.gradP[pOut] : a.gradPs;
                corr = a.corrP ? a.corrP[pOut] : a.corrPs;
            }
        }
    };

    // per-(set, unit) values, (re)loaded per tile when the block time-multiplexes several tiles
    Hru h;
    double fa = 0, faW = 0, fG = 1, fGl = 0, cT = 0, mP = 1;  // faW: weight of the outlet fluxes (Flux.cpp:20-26)
    bool glacierVariant = false, twin = false;
    auto loadUnit = [&](int u) {
        fa = a.hru[0 * a.ldu + u];
        faW = (fa < 1.0) ? fa : 1.0;
        const double area = a.hru[1 * a.ldu + u];
        const double slopeTerm = a.hru[2 * a.ldu + u];
        const double elev = a.hru[3 * a.ldu + u];
        // runoff:socont folded: beta*sqrt(slope)/area * (2/1000)^(5/3) * 86400 * 1000 (ProcessRunoffSocont.cpp:41-56)
        const double socontUnit = 3.174802103936398e-05 * 86400.0 * 1000.0;
        par.quickV = (a.flags.quickKind == 0) ? quickBase * slopeTerm / area * socontUnit : quickBase;
        const int unitFlags = a.hruFlags[u];
        glacierVariant = GLACIER && (unitFlags & 1);
        twin = GLACIER && (unitFlags & 8);
#if !HBK_PAR_SHARED
        if (GLACIER && a.flags.hasTwins) {
            const GlacierPars g = glacierPars(a.params, a.ldp, p, twin);
            par.v[7] = g.snowDdf;
            par.v[8] = g.snowTm;
            par.v[9] = g.iceDdf;
            par.v[10] = g.iceTm;
            par.v[16] = g.snowIce;
        }
        if (a.flags.meltKind == HBK_MELT_DEGREE_DAY_ASPECT) {
            // ProcessMeltDegreeDayAspect::SetParameters (ProcessMeltDegreeDayAspect.cpp:38-58): the melt processes of this
            // unit use the degree-day factor of its aspect class — a per-(set, unit) value, picked when the unit is loaded
            const AspectFactors f3 = aspectFactors<GLACIER>(a.params, a.ldp, p, (unitFlags >> 1) & 3, twin);
            par.v[5] = f3.ground;
            par.v[7] = f3.glacierSnow;
            par.v[9] = f3.ice;
        }
#endif
        const long long fi = a.fracPerSet ? ((long long)p * a.U + u) : u;
        fG = a.frac[fi];
        fGl = a.frac[a.fracLd + fi];
        if (forcingMode == 1) {
            // forcing.py:1073-1075 / 1104-1106: g*(z - zref) -> /100 -> (1 + ...) -> x * ...
            cT = gT * (elev - a.zrefT) / 100;
            mP = 1 + gP * (elev - a.zrefP) / 100;
        }
    };
    // `full` = segment boundary (first load / last store). In between (time-multiplexed tiles) the persistent flux
    // amounts of a non-null brick are dead — every one is rewritten before it is read within the next step — so only
    // the storages travel, plus the amounts of null bricks (which keep their stale values, Processor.cpp:258-260).
    auto loadState = [&](int u, bool full) {
        const long long si = (long long)p * a.sP + (long long)u * a.sU;
        const bool groundAmts = full || nearlyZero(fG, kPrecision);
        const bool glacierAmts = full || !glacierVariant || nearlyZero(fGl, kPrecision);
        h.snowG = __ldcg(&a.state[HBK_S_GROUND_SNOW * a.lds + si]);
        h.wG = __ldcg(&a.state[HBK_S_GROUND_WATER * a.lds + si]);
        h.slow = __ldcg(&a.state[HBK_S_SLOW * a.lds + si]);
        h.slow2 = (NSOIL == 2) ? __ldcg(&a.state[HBK_S_SLOW2 * a.lds + si]) : 0.0;
        h.quick = __ldcg(&a.state[HBK_S_QUICK * a.lds + si]);
        if (groundAmts) {
            h.amtInfil = __ldcg(&a.state[HBK_S_AMT_INFILTRATION * a.lds + si]);
            h.amtRest = __ldcg(&a.state[HBK_S_AMT_REST * a.lds + si]);
            h.amtMeltG = __ldcg(&a.state[HBK_S_AMT_MELT_GROUND * a.lds + si]);
        }
        if (GLACIER) {
            h.snowGl = __ldcg(&a.state[HBK_S_GLACIER_SNOW * a.lds + si]);
            h.wGl = __ldcg(&a.state[HBK_S_GLACIER_WATER * a.lds + si]);
            h.ice = __ldcg(&a.state[HBK_S_ICE * a.lds + si]);
            if (glacierAmts) {
                h.amtMeltGl = __ldcg(&a.state[HBK_S_AMT_MELT_GLACIER * a.lds + si]);
                h.amtDep1 = __ldcg(&a.state[HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT * a.lds + si]);
                h.amtDep2 = __ldcg(&a.state[HBK_S_AMT_DEPOSIT_ICEMELT * a.lds + si]);
                h.amtSnowIce = __ldcg(&a.state[HBK_S_AMT_SNOW_ICE * a.lds + si]);
            }
        } else {
            h.snowGl = h.wGl = h.ice = h.amtMeltGl = h.amtDep1 = h.amtDep2 = h.amtSnowIce = 0.0;
        }
    };
    auto storeState = [&](int u, bool full) {
        const long long si = (long long)p * a.sP + (long long)u * a.sU;
        const bool groundAmts = full || nearlyZero(fG, kPrecision);
        const bool glacierAmts = full || !glacierVariant || nearlyZero(fGl, kPrecision);
        a.state[HBK_S_GROUND_SNOW * a.lds + si] = h.snowG;
        a.state[HBK_S_GROUND_WATER * a.lds + si] = h.wG;
        a.state[HBK_S_SLOW * a.lds + si] = h.slow;
        if (NSOIL == 2) a.state[HBK_S_SLOW2 * a.lds + si] = h.slow2;
        a.state[HBK_S_QUICK * a.lds + si] = h.quick;
        if (groundAmts) {
            a.state[HBK_S_AMT_INFILTRATION * a.lds + si] = h.amtInfil;
            a.state[HBK_S_AMT_REST * a.lds + si] = h.amtRest;
            a.state[HBK_S_AMT_MELT_GROUND * a.lds + si] = h.amtMeltG;
        }
        if (GLACIER) {
            a.state[HBK_S_GLACIER_SNOW * a.lds + si] = h.snowGl;
            a.state[HBK_S_GLACIER_WATER * a.lds + si] = h.wGl;
            a.state[HBK_S_ICE * a.lds + si] = h.ice;
            if (glacierAmts) {
                a.state[HBK_S_AMT_MELT_GLACIER * a.lds + si] = h.amtMeltGl;
                a.state[HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT * a.lds + si] = h.amtDep1;
                a.state[HBK_S_AMT_DEPOSIT_ICEMELT * a.lds + si] = h.amtDep2;
                a.state[HBK_S_AMT_SNOW_ICE * a.lds + si] = h.amtSnowIce;
            }
        }
    };

    const double dt = DAILY ? 1.0 : a.dt;
    const double invDt = DAILY ? 1.0 : 1.0 / a.dt;
    const int nSteps = a.tEnd - a.tBegin;
    const int nChunks = (nSteps + TC - 1) / TC;
    // the chunks [cBegin, cEnd) of the current work item (the whole segment without the queue) and its forcing loads:
    // one per chunk (station series) or per (chunk, tile) (per-unit series), numbered seq = 0.. within the item;
    // loadBase = loads issued by earlier items of this block (the two mbarriers alternate over the running number)
    int cBegin = 0, cEnd = nChunks, nTilesHere = 0, nLoads = 0, loadBase = 0;

    auto issueLoad = [&](int seq) {
        const int stage = (loadBase + seq) & 1;
        double* dst = fbuf + (size_t)stage * NV * fRow;
        if (forcingMode == 0) {
            const int c = cBegin + seq / nTilesHere, tile = tileBegin + seq % nTilesHere;
            const int t0 = a.tBegin + c * TC;
            const int n = min(TC, a.tEnd - t0);
            const uint32_t bytes = (uint32_t)(n * UB * sizeof(double));
            mbarExpectTx(&bars[stage], NV * bytes);
            for (int v = 0; v < NV; ++v) {
                const double* src = a.forcing[v] + ((size_t)tile * a.tld + t0) * UB;
                tmaLoad1D(dst + (size_t)v * fRow, src, bytes, &bars[stage]);
            }
        } else {
            const int t0 = a.tBegin + (cBegin + seq) * TC;
            const int t0a = t0 & ~1;  // 16-byte aligned source
            const uint32_t bytes = (uint32_t)((TC + 2) * sizeof(double));
            mbarExpectTx(&bars[stage], NV * bytes);
            for (int v = 0; v < NV; ++v) tmaLoad1D(dst + (size_t)v * fRow, a.forcing[v] + t0a, bytes, &bars[stage]);
        }
    };

    if (a.blockTimes && tid == 0) {
        unsigned long long t;
        unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        a.blockTimes[3 * (size_t)blockIdx.x] = t;
        a.blockTimes[3 * (size_t)blockIdx.x + 2] = sm;
    }
    if (tid == 0) {
        mbarInit(&bars[0], 1);
        mbarInit(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nRec = __popcll(a.recordMask);
    const bool queued = a.workCounter != nullptr;
    __shared__ int sItem;
    for (;;) {  // work items (exactly one without the queue)
    int slice = 0, vbCur = 0;
    if (queued) {
        // Persistent blocks, items in slice-major order: every virtual block advances slice by slice, whichever block
        // is free takes the next item. A set-group that is slower than the others no longer holds a multiprocessor
        // slot to the end of the launch while the rest of the device drains (profiles/r02_tuning.md: 13 % of the
        // block-time of the C2 launch was such a tail).
        if (tid == 0) sItem = atomicAdd(a.workCounter, 1);
        __syncthreads();
        const int item = sItem;
        __syncthreads();
        if (item >= a.nItems) break;
        slice = item / a.nVBlocks;
        const int vb = item - slice * a.nVBlocks;
        vbCur = vb;
        setBlock(vb);
        cBegin = slice * a.sliceChunks;
        cEnd = min(cBegin + a.sliceChunks, nChunks);
        if (slice > 0) {
            // the previous slice of this virtual block has a smaller item number, so a running block holds it (or has
            // finished it): wait for its state to be published
            if (tid == 0) {
                int done;
                long long spins = 0;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(a.progress + vb) : "memory");
                    if (done < slice) __nanosleep(256);
                } while (done < slice && ++spins < (1LL << 27));  // ~30 s: a logic error must not hang the device
                if (done < slice) atomicOr(a.err, kErrQueueStall);
            }
            __syncthreads();
        }
    } else {
        setBlock(blockIdx.x);
    }
    nTilesHere = tileEnd - tileBegin;
    nLoads = (forcingMode == 0) ? (cEnd - cBegin) * nTilesHere : (cEnd - cBegin);
    if (tid == 0 && nLoads > 0) issueLoad(0);

    int err = 0;
    int unconverged = 0;
    bool loaded = false;

    for (int c = cBegin; c < cEnd; ++c) {
        const int t0 = a.tBegin + c * TC;
        const int n = min(TC, a.tEnd - t0);
        const int off = (forcingMode == 1) ? (t0 & 1) : 0;
        for (int tl = 0; tl < nTilesHere; ++tl) {
            const int tile = tileBegin + tl;
            const int u = tile * UB + ul;
            const bool valid = validP && (u < a.U);
            // ---- forcing for this (chunk, tile) ----
            const int seq = (forcingMode == 0) ? (c - cBegin) * nTilesHere + tl : (c - cBegin);
            const int gseq = loadBase + seq;
            if (forcingMode == 0 || tl == 0) {
                if (tid == 0 && seq + 1 < nLoads) issueLoad(seq + 1);
                mbarWait(&bars[gseq & 1], (gseq >> 1) & 1);
            }
            const double* fs = fbuf + (size_t)(gseq & 1) * NV * fRow;
            if (valid && (multi || !loaded)) {
                loadUnit(u);
                loadState(u, c == cBegin);
            }
            loaded = true;

            // this thread's forcing column in the staged chunk (per-unit series: its unit's column; station series: the
            // shared series, spatialised below) and its outlet slots
            const double* fcol = fs + ((forcingMode == 0) ? ul : off);
            const int fStep = (forcingMode == 0) ? UB : 1;
            double* slotCol = qbuf + tid;
            for (int tc = 0; tc < n; ++tc) {
                StepOut o;
                o.q = 0.0;
                o.dep1 = 0.0;
                o.dep2 = 0.0;
                if (valid) {
                    const double* fp = fcol + tc * fStep;
                    double precip = fp[0], temp = fp[fRow], pet = fp[2 * fRow];
                    if (forcingMode != 0) {  // fused spatialisation (forcing.py:1061-1113)
                        temp = temp + cT;
                        precip = (precip * corr) * mP;
                        precip = (precip < 0.0) ? 0.0 : precip;
                        pet = (pet < 0.0) ? 0.0 : pet;
                    }
#if !HBK_PAR_SHARED
                    if (!ENS && a.flags.meltKind == HBK_MELT_TEMPERATURE_INDEX) {
                        // ProcessMeltTemperatureIndex.cpp:68-71: (T - Tm) * (melt_factor + radiation_coefficient * R); the
                        // factor in brackets takes the place of the degree-day factor for this step
                        const AspectFactors f3 = radiationFactors<GLACIER>(a.params, a.ldp, p, fp[3 * fRow], twin);
                        par.v[5] = f3.ground;
                        par.v[7] = f3.glacierSnow;
                        par.v[9] = f3.ice;
                    }
#endif
                    double sublimG = 0.0, sublimGl = 0.0;
                    if (a.flags.sublimKind != 0) {
                        const double2 su = sublimationRates<GLACIER>(a.params, a.ldp, a.flags.sublimKind, p, pet, twin);
                        sublimG = su.x;
                        sublimGl = su.y;
                    }
                    RetentionCtx rc;
                    const RetentionCtx* ret = nullptr;
                    if (!ENS && a.flags.retention) {  // liquid water kept in the snowpacks: generic-shape kernels only
                        rc = retentionCtx(a.state, a.lds, (long long)p * a.sP + (long long)u * a.sU, a.params, a.ldp, p,
                                          a.flags.retention);
                        ret = &rc;
                    }
                    directPhase<GLACIER>(h, par, precip, temp, dt, invDt, fa, fG, fGl, glacierVariant, a.flags, o, err,
                                         (GLACIER && a.swatFactor) ? a.swatFactor[t0 + tc] : 0.0, sublimG, sublimGl, 0.0, 0.0,
                                         NoSlide(), NoSlide(), ret);
                    if constexpr (SOLVER == 3) {
                        // the ensemble-shape instantiation of the sequential family is crank_nicolson's (the default of the
                        // reference's Python layer): the other members run the generic instantiation
                        solveUnitSequential<NSOIL, (SHAPE == 1 ? (int)kSeqCrank : -1)>(h, par, pet, faW, dt, invDt, a.flags, o,
                                                                                         err);
                    } else {
                        solveUnit<SOLVER, NSOIL>(h, par, pet, faW, dt, invDt, a.flags, o, err, unconverged);
                    }

                    if (nRec && a.writeOutputs) recordStep(a, pOut, t0 + tc, u, h, o, fG, fGl, glacierVariant, ret);
                }
                // this thread's slot accumulates its units of the group's tiles in tile order (no race: own slot)
                double* slot = slotCol + (size_t)tc * nThreads;
                const bool first = clusterMode || tl == 0;
                if (!first) o.q += slot[0];  // slot + x == x + slot
                slot[0] = o.q;
                if (GLACIER) {
                    double* s1 = slot + (size_t)TC * nThreads;
                    double* s2 = s1 + (size_t)TC * nThreads;
                    if (!first) {
                        o.dep1 += s1[0];
                        o.dep2 += s2[0];
                    }
                    s1[0] = o.dep1;
                    s2[0] = o.dep2;
                }
            }
            if (valid && multi) storeState(u, c == cEnd - 1);
            if (clusterMode) {
                __syncthreads();
                // this CTA's tile sum -> its partial buffer (double buffered by chunk parity) ...
                double* part = accb + (size_t)(c & 1) * NQ * TC * PB;
                for (int idx = tid; idx < NQ * n * PB; idx += nThreads) {
                    const int qi = idx / (n * PB);
                    const int rem = idx - qi * n * PB;
                    const int tc = rem / PB;
                    const int lp = rem - tc * PB;
                    const double* src = qbuf + ((size_t)qi * TC + tc) * nThreads + lp;
                    double s = 0.0;
                    for (int k = 0; k < UB; ++k) s += src[(size_t)k * PB];
                    part[((size_t)qi * TC + tc) * PB + lp] = s;
                }
                // ... then the rows (quantity, step) are dealt round-robin to the cluster's CTAs, which add the
                // tiles' partials in tile order through DSMEM and write the final series.
                namespace cg = cooperative_groups;
                cg::cluster_group cluster = cg::this_cluster();
                cluster.sync();
                const int nRows = NQ * n;
                for (int row = crank; row < nRows; row += a.nUTiles) {
                    const int qi = row / n, tc = row - qi * n;
                    for (int lp = tid; lp < PB; lp += nThreads) {
                        const int gp = pg * PB + lp;
                        double s = 0.0;
                        for (int k = 0; k < a.nUTiles; ++k) {
                            const double* remote = cluster.map_shared_rank(part, k);
                            s += remote[((size_t)qi * TC + tc) * PB + lp];
                        }
                        if (gp < a.P && (a.writeOutputs || qi > 0)) {
                            double* dst = (qi == 0) ? a.outQ : (qi == 1 ? a.outD1 : a.outD2);
                            const int row = (qi == 0 && a.perm) ? a.perm[gp] : gp;
                            __stcs(&dst[(size_t)row * a.ldt + (t0 + tc)], s);
                        }
                    }
                }
                continue;  // the next chunk writes the other partial buffer; its cluster.sync orders the reuse
            }
            // SubBasin::ComputeOutletDischarge (SubBasin.cpp:322-329): after the last tile of the group, the UB
            // per-thread sums of every (quantity, step, set) are added in thread order and written out.
            const bool last = (tl == nTilesHere - 1);
            if (last) {
                __syncthreads();
                if (SHAPE != 1 && UB >= 32) {
                    // few sets, many units per block (distributed basins): one WARP per (quantity, step, set) row — the
                    // lanes add strided slices of the row's UB values in unit order, then a butterfly over the lanes
                    // (a single thread per row would walk up to 128 dependent additions while the block waits)
                    const int lane = tid & 31;
                    for (int row = tid >> 5; row < NQ * n * PB; row += nThreads >> 5) {
                        const int qi = row / (n * PB);
                        const int rem = row - qi * n * PB;
                        const int tc = rem / PB;
                        const int lp = rem - tc * PB;
                        const double* src = qbuf + ((size_t)qi * TC + tc) * nThreads + lp;
                        double s = 0.0;
                        for (int k = lane; k < UB; k += 32) s += src[(size_t)k * PB];
                        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                        const int gp = pg * PB + lp;
                        if (lane == 0 && gp < a.P && (a.writeOutputs || qi > 0)) {
                            double* dst = (qi == 0) ? a.outQ : (qi == 1 ? a.outD1 : a.outD2);
                            const int orow = (qi == 0 && a.outQFinal && a.perm) ? a.perm[gp] : gp;
                            __stcs(&dst[((size_t)(a.grpBase + grp) * a.P + orow) * a.ldt + (t0 + tc)], s);
                        }
                    }
                } else
                for (int idx = tid; idx < NQ * n * PB; idx += nThreads) {
                    const int qi = idx / (n * PB);
                    const int rem = idx - qi * n * PB;
                    const int tc = rem / PB;
                    const int lp = rem - tc * PB;
                    const double* src = qbuf + ((size_t)qi * TC + tc) * nThreads + lp;
                    double s = 0.0;
                    for (int k = 0; k < UB; ++k) s += src[(size_t)k * PB];
                    const int gp = pg * PB + lp;
                    if (gp < a.P && (a.writeOutputs || qi > 0)) {  // deposits feed the basin stores during spin-up
                        double* dst = (qi == 0) ? a.outQ : (qi == 1 ? a.outD1 : a.outD2);
                        const int row = (qi == 0 && a.outQFinal && a.perm) ? a.perm[gp] : gp;
                        // streaming store: the 8.8 GB outlet must not evict the L2-resident unit state
                        __stcs(&dst[((size_t)(a.grpBase + grp) * a.P + row) * a.ldt + (t0 + tc)], s);
                    }
                }
            }
            // the slots are rewritten by the next chunk; per-unit forcing stages are recycled tile by tile
            if (last || forcingMode == 0) __syncthreads();
        }
    }

    if (clusterMode) cooperative_groups::this_cluster().sync();  // peers may still be reading this CTA's partials
    if (validP && !multi && nTilesHere > 0) {
        const int u = tileBegin * UB + ul;
        if (u < a.U) storeState(u, true);
    }
    if (err) atomicOr(a.err, err);
    if (unconverged) {
        atomicAdd(a.unconverged, (unsigned long long)unconverged);
        if (validP) atomicAdd(&a.unconvergedPerSet[p], unconverged);
    }
    if (!queued) break;
    loadBase += nLoads;
    // publish the slice: every thread's state stores, then the counter
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int done = slice + 1;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.progress + vbCur), "r"(done) : "memory");
    }
    }  // work items
    if (a.blockTimes && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.blockTimes[3 * (size_t)blockIdx.x + 1] = t;
    }
}

// ---- lateral kernel (transport:snow_slide) -----------------------------------------------------------
// The snow slide couples the units of a parameter set INSIDE a time step: a snowpack that exceeds its holding depth hands
// snow, instantaneously, to the snowpacks of the units it is connected to, and the reference processes the units in basin
// order (Processor.cpp:237-276), so a receiver that comes later sees the snow in the same step. One thread therefore owns
// one parameter set and walks its units in basin order, step by step; the unit state lives in the state arrays (element
// (p, u) at u * P + p: the loads and stores of a warp's 32 sets are contiguous), the per-step arithmetic is the same
// device code as everywhere else (directPhase with the slide hooked into the snowpack step, solveUnit). Parallelism is the
// parameter axis only — the option is for glacier-evolution studies with one or a few hundred sets, not the ensemble path.
struct LatArgs {
    const int* off;        // [U + 1]: the connections of giver u are [off[u], off[u + 1]), in the order they were added
    const int* recv;       // [nConn] receiving unit (> giver)
    const double* weight;  // [nConn] share of the giver's sliding snow
    const double* thr;     // [U][ldp] snow holding threshold [mm of snow depth] of (unit, set), evaluated on the host
    double* pend;          // [2][U][P] snow received this step, not yet taken in (ground / glacier snowpack)
    int skipGlaciers;      // the glacier snowpacks receive but do not give (SettingsModel.cpp:565-567)
    const double* raw[HBK_N_VARS];  // forcing mode 0: the API layout [T][U]
};

template <int SOLVER, int NSOIL, bool GLACIER>
__global__ void __launch_bounds__(128) hbkRunLateral(const KArgs a, const LatArgs L) {
#if HBK_PAR_SHARED  // tuning builds with the parameters in shared memory have no lateral kernel (refused at engine creation)
    (void)a;
    (void)L;
#else
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    const int pOut = a.perm ? a.perm[p] : p;
    Par par;
    fillPar(par.v, 1, a.params + p, a.ldp);
    const double quickBase = a.params[p + (size_t)HBK_P_QUICK * a.ldp];
    double gT = 0, gP = 0, corr = 1;
    if (a.forcingMode == 1) {
        gT = a.gradT ? a.gradT[pOut] : a.gradTs;
        gP = a.gradP ? a.gradP[pOut] : a.gradPs;
        corr = a.corrP ? a.corrP[pOut] : a.corrPs;
    }
    const double dt = a.dt, invDt = 1.0 / a.dt;
    const int nRec = __popcll(a.recordMask);
    const float maxSnowDepth = a.params[p + (size_t)HBK_P_SLIDE_MAX_SNOW_DEPTH * a.ldp];
    // AvoidUnrealisticAccumulation (ProcessLateralSnowSlide.cpp:120-134): 1.5 * max_snow_depth / (water / snow density)
    const double accumulationLimit = 1.5 * maxSnowDepth / 4.0f;
    const size_t PU = (size_t)a.P * a.U;
    int err = 0, unconverged = 0;
    for (int t = a.tBegin; t < a.tEnd; ++t) {
        double q = 0.0, d1 = 0.0, d2 = 0.0;
        for (int u = 0; u < a.U; ++u) {
            // ---- unit constants (as hbkRunUnits loadUnit)
            const double fa = a.hru[0 * a.ldu + u];
            const double faW = (fa < 1.0) ? fa : 1.0;
            const double area = a.hru[1 * a.ldu + u];
            const double socontUnit = 3.174802103936398e-05 * 86400.0 * 1000.0;
            par.quickV = (a.flags.quickKind == 0) ? quickBase * a.hru[2 * a.ldu + u] / area * socontUnit : quickBase;
            const int unitFlags = a.hruFlags[u];
            const bool gv = GLACIER && (unitFlags & 1);
            if (a.flags.meltKind == HBK_MELT_DEGREE_DAY_ASPECT) {
                const AspectFactors f3 = aspectFactors<GLACIER>(a.params, a.ldp, p, (unitFlags >> 1) & 3);
                par.v[5] = f3.ground;
                par.v[7] = f3.glacierSnow;
                par.v[9] = f3.ice;
            }
            const long long fi = a.fracPerSet ? ((long long)p * a.U + u) : u;
            const double fG = a.frac[fi], fGl = a.frac[a.fracLd + fi];
            // ---- forcing
            double precip, temp, pet, rad = 0.0;
            if (a.forcingMode == 0) {
                const size_t k = (size_t)t * a.U + u;
                precip = L.raw[0][k];
                temp = L.raw[1][k];
                pet = L.raw[2][k];
                if (a.nVars == 4) rad = L.raw[3][k];
            } else {  // station series + fused spatialisation (forcing.py:1061-1113)
                const double elev = a.hru[3 * a.ldu + u];
                temp = a.forcing[1][t] + gT * (elev - a.zrefT) / 100;
                precip = (a.forcing[0][t] * corr) * (1 + gP * (elev - a.zrefP) / 100);
                precip = (precip < 0.0) ? 0.0 : precip;
                pet = a.forcing[2][t];
                pet = (pet < 0.0) ? 0.0 : pet;
                if (a.nVars == 4) rad = a.forcing[3][t];
            }
            if (a.flags.meltKind == HBK_MELT_TEMPERATURE_INDEX) {
                const AspectFactors f3 = radiationFactors<GLACIER>(a.params, a.ldp, p, rad);
                par.v[5] = f3.ground;
                par.v[7] = f3.glacierSnow;
                par.v[9] = f3.ice;
            }
            double sublimG = 0.0, sublimGl = 0.0;
            if (a.flags.sublimKind != 0) {
                const double2 su = sublimationRates<GLACIER>(a.params, a.ldp, a.flags.sublimKind, p, pet);
                sublimG = su.x;
                sublimGl = su.y;
            }
            // ---- state
            const size_t si = (size_t)p * a.sP + (size_t)u * a.sU;
            Hru h;
            h.snowG = a.state[HBK_S_GROUND_SNOW * a.lds + si];
            h.wG = a.state[HBK_S_GROUND_WATER * a.lds + si];
            h.slow = a.state[HBK_S_SLOW * a.lds + si];
            h.slow2 = (NSOIL == 2) ? a.state[HBK_S_SLOW2 * a.lds + si] : 0.0;
            h.quick = a.state[HBK_S_QUICK * a.lds + si];
            h.amtInfil = a.state[HBK_S_AMT_INFILTRATION * a.lds + si];
            h.amtRest = a.state[HBK_S_AMT_REST * a.lds + si];
            h.amtMeltG = a.state[HBK_S_AMT_MELT_GROUND * a.lds + si];
            if (GLACIER) {
                h.snowGl = a.state[HBK_S_GLACIER_SNOW * a.lds + si];
                h.wGl = a.state[HBK_S_GLACIER_WATER * a.lds + si];
                h.ice = a.state[HBK_S_ICE * a.lds + si];
                h.amtMeltGl = a.state[HBK_S_AMT_MELT_GLACIER * a.lds + si];
                h.amtDep1 = a.state[HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT * a.lds + si];
                h.amtDep2 = a.state[HBK_S_AMT_DEPOSIT_ICEMELT * a.lds + si];
                h.amtSnowIce = a.state[HBK_S_AMT_SNOW_ICE * a.lds + si];
            } else {
                h.snowGl = h.wGl = h.ice = h.amtMeltGl = h.amtDep1 = h.amtDep2 = h.amtSnowIce = 0.0;
            }
            // snow that earlier units of this step slid onto this unit's snowpacks
            double* pendG = L.pend + (size_t)u * a.P + p;
            double* pendGl = L.pend + PU + (size_t)u * a.P + p;
            const double receivedG = *pendG, receivedGl = GLACIER ? *pendGl : 0.0;
            *pendG = 0.0;
            if (GLACIER) *pendGl = 0.0;

            // ProcessLateralSnowSlide::GetRates + WaterContainer::ApplyConstraints + Process::ApplyChange for the slide
            // process of one of this unit's snowpacks (fOrigin: area fraction of its land cover)
            auto slideOf = [&](double fOrigin) {
                return [&, fOrigin](double content, double stat, double inputsStatic, double& dyn, bool& tooHigh) {
                    const int c0 = L.off[u], c1 = L.off[u + 1];
                    if (c0 == c1) return;  // no connection: StoreRates({})
                    const double swe = content + stat;  // GetContentWithChanges
                    if (swe <= kPrecision) return;      // Process::GetChangeRates (Process.cpp:480-487): all rates 0
                    const double snowDepth = swe * 4.0f;  // constants::waterDensity / constants::snowDensity, a float
                    const double threshold = L.thr[(size_t)u * a.ldp + p];
                    double excess = 0.0;
                    if (snowDepth > threshold) excess = (snowDepth - threshold) / 4.0f;
                    if (excess <= 0.0) return;
                    if (excess > 1000.0) excess = 1000.0;
                    // one output per (connection, snowpack of the receiving unit): ground, then glacier where the
                    // receiver carries the glacier bricks (ModelBuilder.cpp:476-508); rate < 0 stands for "no such output"
                    auto rateOf = [&](int k, int cover, double& fTarget) -> double {
                        const int r = L.recv[k];
                        if (cover == 1 && !(a.hruFlags[r] & 1)) return -1.0;
                        const long long fr = a.fracPerSet ? ((long long)p * a.U + r) : r;
                        fTarget = a.frac[cover * a.fracLd + fr];
                        if (nearlyZero(fTarget, kPrecision)) return 0.0;
                        double rate = excess * L.weight[k] * fTarget;
                        if (maxSnowDepth > 0.0f) {  // AvoidUnrealisticAccumulation: the receiver's snow content
                            const double targetSwe = a.state[(cover ? HBK_S_GLACIER_SNOW : HBK_S_GROUND_SNOW) * a.lds +
                                                             (size_t)p * a.sP + (size_t)r * a.sU];
                            if (targetSwe > accumulationLimit) rate = 0.0;
                        }
                        return rate;
                    };
                    const int nCovers = GLACIER ? 2 : 1;
                    double outputs = 0.0;
                    for (int k = c0; k < c1; ++k) {
                        for (int cover = 0; cover < nCovers; ++cover) {
                            double fTarget;
                            const double rate = rateOf(k, cover, fTarget);
                            if (rate < 0.0) continue;
                            tooHigh = tooHigh || rate > 10000.0;
                            outputs += rate;
                        }
                    }
                    // WaterContainer::ApplyConstraints (WaterContainer.cpp:117-142) on the snow container
                    const double change = -outputs;
                    const double balance = content + inputsStatic + change * dt;
                    const bool limited = change < 0.0 && balance < 0.0;
                    const double diff = balance / dt;
                    for (int k = c0; k < c1; ++k) {
                        for (int cover = 0; cover < nCovers; ++cover) {
                            double fTarget = 0.0;
                            double rate = rateOf(k, cover, fTarget);
                            if (rate < 0.0) continue;
                            if (limited && !nearlyZero(rate, kEpsD)) {
                                rate = (fabs(diff - change) <= kPrecision) ? 0.0 : rate + diff * fabs(rate / outputs);
                            }
                            rate = (rate < 0.0) ? 0.0 : rate;  // Process::ApplyChange (Process.cpp:495-508)
                            if (rate > kPrecision) {
                                const double amount = rate * dt;
                                dyn -= amount;
                                // ProcessLateral::ComputeFractionAreas (ProcessLateral.cpp:67-76) -> the flux's weight;
                                // FluxToBrickInstantaneous::UpdateFlux books it into the receiver's static change
                                const int r = L.recv[k];
                                const double fractionAreas = (area * fOrigin) / (a.hru[1 * a.ldu + r] * fTarget);
                                const double booked = (fractionAreas != 1.0) ? amount * fractionAreas : amount;
                                L.pend[(size_t)cover * PU + (size_t)r * a.P + p] += booked;
                            }
                        }
                    }
                };
            };
            auto slideG = slideOf(GLACIER ? fG : 1.0);
            auto slideGlAll = slideOf(fGl);
            auto slideGl = [&](double content, double stat, double inputsStatic, double& dyn, bool& tooHigh) {
                if (!L.skipGlaciers) slideGlAll(content, stat, inputsStatic, dyn, tooHigh);
            };
            StepOut o;
            directPhase<GLACIER, true, decltype(slideG), decltype(slideGl)>(
                h, par, precip, temp, dt, invDt, fa, fG, fGl, gv, a.flags, o, err,
                (GLACIER && a.swatFactor) ? a.swatFactor[t] : 0.0, sublimG, sublimGl, receivedG, receivedGl, slideG, slideGl);
            if constexpr (SOLVER == 3) {
                solveUnitSequential<NSOIL>(h, par, pet, faW, dt, invDt, a.flags, o, err);
            } else {
                solveUnit<SOLVER, NSOIL>(h, par, pet, faW, dt, invDt, a.flags, o, err, unconverged);
            }
            if (nRec && a.writeOutputs) recordStep(a, pOut, t, u, h, o, fG, fGl, gv);
            // ---- state back
            a.state[HBK_S_GROUND_SNOW * a.lds + si] = h.snowG;
            a.state[HBK_S_GROUND_WATER * a.lds + si] = h.wG;
            a.state[HBK_S_SLOW * a.lds + si] = h.slow;
            if (NSOIL == 2) a.state[HBK_S_SLOW2 * a.lds + si] = h.slow2;
            a.state[HBK_S_QUICK * a.lds + si] = h.quick;
            a.state[HBK_S_AMT_INFILTRATION * a.lds + si] = h.amtInfil;
            a.state[HBK_S_AMT_REST * a.lds + si] = h.amtRest;
            a.state[HBK_S_AMT_MELT_GROUND * a.lds + si] = h.amtMeltG;
            if (GLACIER) {
                a.state[HBK_S_GLACIER_SNOW * a.lds + si] = h.snowGl;
                a.state[HBK_S_GLACIER_WATER * a.lds + si] = h.wGl;
                a.state[HBK_S_ICE * a.lds + si] = h.ice;
                a.state[HBK_S_AMT_MELT_GLACIER * a.lds + si] = h.amtMeltGl;
                a.state[HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT * a.lds + si] = h.amtDep1;
                a.state[HBK_S_AMT_DEPOSIT_ICEMELT * a.lds + si] = h.amtDep2;
                a.state[HBK_S_AMT_SNOW_ICE * a.lds + si] = h.amtSnowIce;
            }
            // SubBasin::ComputeOutletDischarge (SubBasin.cpp:322-329): units in basin order
            q += o.q;
            d1 += o.dep1;
            d2 += o.dep2;
        }
        if (a.writeOutputs) a.outQ[(size_t)pOut * a.ldt + t] = q;
        if (GLACIER) {
            a.outD1[(size_t)p * a.ldt + t] = d1;
            a.outD2[(size_t)p * a.ldt + t] = d2;
        }
    }
    if (err) atomicOr(a.err, err);
    if (unconverged) {
        atomicAdd(a.unconverged, (unsigned long long)unconverged);
        atomicAdd(&a.unconvergedPerSet[p], unconverged);
    }
#endif
}

// Sum the unit-tile partials in tile order: out[p][t] = sum_tiles part[tile][p][t]  (only when nUTiles > 1).
__global__ void hbkReduceTiles(const double* __restrict__ part, double* __restrict__ out, int nTiles, int P,
                               long long ldt, int tBegin, int tEnd, const int* __restrict__ perm) {
    const long long n = (long long)P * (tEnd - tBegin);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / (tEnd - tBegin));
        const int t = tBegin + (int)(i - (long long)p * (tEnd - tBegin));
        double s = 0.0;
        for (int k = 0; k < nTiles; ++k) s += part[((size_t)k * P + p) * ldt + t];
        out[(size_t)(perm ? perm[p] : p) * ldt + t] = s;
    }
}

// The two sub-basin linear stores, one thread per parameter set, sequential in time
// (modules/glacier.py:108-131; solved with the same solver as the units, Processor.cpp:94-111).
template <int SOLVER>
__global__ void hbkBasinStores(const float* params, int ldp, int P, long long ldt, int tBegin, int tEnd, double dt,
                               const double* d1, const double* d2, const double* stale, double* outlet,
                               double* basinState, int writeOutputs, int* errFlags, unsigned long long* unconv,
                               const int* perm, int seqKind) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const size_t row = perm ? perm[p] : p;  // outlet rows are in the caller's order
    const double kSnow = params[HBK_P_K_SNOW * ldp + p];
    const double kIce = params[HBK_P_K_ICE * ldp + p];
    double c1 = basinState[2 * p], c2 = basinState[2 * p + 1];
    const double st1 = stale ? stale[2 * p] : 0.0, st2 = stale ? stale[2 * p + 1] : 0.0;
    int err = 0, unc = 0;
    for (int t = tBegin; t < tEnd; ++t) {
        const double dep1 = d1[(size_t)p * ldt + t], dep2 = d2[(size_t)p * ldt + t];
        const double o1 = basinStoreStep<SOLVER>(c1, kSnow, dep1, dep1 + st1, dt, err, unc, seqKind);
        const double o2 = basinStoreStep<SOLVER>(c2, kIce, dep2, dep2 + st2, dt, err, unc, seqKind);
        if (writeOutputs) outlet[row * ldt + t] = (o1 + o2) + outlet[row * ldt + t];
    }
    basinState[2 * p] = c1;
    basinState[2 * p + 1] = c2;
    if (err) atomicOr(errFlags, err);
    if (unc) atomicAdd(unconv, (unsigned long long)unc);
}

// [T][U] (API layout of ModelHydro::CreateTimeSeries) -> [tile][tld][UB] so that a block's chunk is contiguous.
__global__ void hbkTileForcing(const double* __restrict__ src, double* __restrict__ dst, int T, int U, int UB,
                               long long tld, int nTiles) {
    const long long n = (long long)nTiles * T * UB;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ul = (int)(i % UB);
        const long long r = i / UB;
        const int t = (int)(r % T);
        const int tile = (int)(r / T);
        const int u = tile * UB + ul;
        dst[((size_t)tile * tld + t) * UB + ul] = (u < U) ? src[(size_t)t * U + u] : 0.0;
    }
}

__global__ void hbkBroadcastState(double* dst, const double* perUnit, double value, int P, int U, int laneP) {
    const long long n = (long long)P * U;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = perUnit ? perUnit[laneP ? i / P : i % U] : value;
}

// ---- dated actions (kernel boundaries) ------------------------------------------------------------------
struct ActArgs {
    int P, U;
    double* frac;  // [2][P*U] per set: ground, glacier
    double* state;
    long long lds, sP, sU;
    const double* hru;  // unit constants (area at [1][u])
    const int* hruFlags;
    int ldu;
    int iceInfinite;
    int retention;  // HBK_RETENTION_*: the snowpacks hold liquid water, which moves with the land like the snow
    int* err;
    // twin units (hbk_set_unit_twins): twinOf[u] = unit whose second glacier cover u carries (-1: an ordinary unit),
    // twinCol[u] = the twin of unit u (-1: none); nullptr without twins
    const int* twinOf;
    const int* twinCol;
};

__device__ __forceinline__ double& stateAt(const ActArgs& a, int slot, int p, int u) {
    return a.state[(size_t)slot * a.lds + (long long)p * a.sP + (long long)u * a.sU];
}

// HydroUnit::ChangeLandCoverAreaFraction for the glacier cover of unit u, parameter set p: the generic (ground)
// cover absorbs the area difference and the water / snow column of the strip that changes hands moves with it
// (HydroUnit.cpp:332-374 ChangeLandCoverAreaFraction, :384-451 TransferLandCoverContent, :459-490 FixLandCoverFractionsTotal).
__device__ void changeGlacierFraction(const ActArgs& a, int p, int u, double fraction) {
    // With twin units the glacier cover being changed lives in u (an ordinary unit: the first glacier cover; a twin: the
    // second), the generic cover that absorbs the difference in the unit itself (uGen), and the unit's other glacier
    // cover (uOther, -1: none) only enters the total of FixLandCoverFractionsTotal.
    const int uGen = (a.twinOf && a.twinOf[u] >= 0) ? a.twinOf[u] : u;
    const int uOther = !a.twinOf ? -1 : (uGen != u ? uGen : a.twinCol[u]);
    const size_t PU = (size_t)a.P * a.U;
    double& fG = a.frac[(size_t)p * a.U + uGen];
    double& fGl = a.frac[PU + (size_t)p * a.U + u];
    if (fraction < 0 || fraction > 1 || !(a.hruFlags[u] & 1)) {
        atomicOr(a.err, kErrActionFraction);
        return;
    }
    const double oldFraction = fGl;
    const double delta = fraction - oldFraction;
    const double genericOld = fG;
    const double genericNew = genericOld - delta;
    if (genericNew < 0 - kEpsD) {
        atomicOr(a.err, kErrActionFraction);
        return;
    }
    if (!nearlyZero(delta, kEpsD)) {
        double& wG = stateAt(a, HBK_S_GROUND_WATER, p, uGen);
        double& wGl = stateAt(a, HBK_S_GLACIER_WATER, p, u);
        double& snowG = stateAt(a, HBK_S_GROUND_SNOW, p, uGen);
        double& snowGl = stateAt(a, HBK_S_GLACIER_SNOW, p, u);
        double& ice = stateAt(a, HBK_S_ICE, p, u);
        // the snowpacks' liquid water (CoverRole::SnowpackWater, HydroUnit.cpp:42-46): a storage only with HBK_RETENTION_*
        double none = 0.0;
        double& swG = a.retention ? stateAt(a, HBK_S_GROUND_SNOWPACK_WATER, p, uGen) : none;
        double& swGl = a.retention ? stateAt(a, HBK_S_GLACIER_SNOWPACK_WATER, p, u) : none;
        if (delta > 0) {  // the glacier grows, taking land (and its water column) from the generic cover
            const double fRO = oldFraction, fRN = fraction, fDN = genericNew, strip = delta;
            if (!nearlyZero(fRN, kEpsD)) {
                wGl = (wGl * fRO + wG * strip) / fRN;
                if (!a.iceInfinite) ice = (ice * fRO + 0.0 * strip) / fRN;
                snowGl = (snowGl * fRO + snowG * strip) / fRN;
                if (a.retention) swGl = (swGl * fRO + swG * strip) / fRN;
                if (nearlyZero(fDN, kEpsD)) {
                    wG = 0;
                    snowG = 0;
                    swG = 0;
                }
            }
        } else {  // the glacier shrinks, giving land to the generic cover; the ice stays with the glacier
            const double fRO = genericOld, fRN = genericNew, fDN = fraction, strip = -delta;
            if (!nearlyZero(fRN, kEpsD)) {
                wG = (wG * fRO + wGl * strip) / fRN;
                snowG = (snowG * fRO + snowGl * strip) / fRN;
                if (a.retention) swG = (swG * fRO + swGl * strip) / fRN;
                if (nearlyZero(fDN, kEpsD)) {
                    wGl = 0;
                    if (!a.iceInfinite) ice = 0;
                    snowGl = 0;
                    swGl = 0;
                }
            }
        }
    }
    fGl = fraction;
    // HydroUnit::FixLandCoverFractionsTotal (HydroUnit.cpp:459-490): the covers in brick order — generic, first glacier,
    // second glacier
    double total = 0;
    total += fG;
    if (uOther < 0) {
        total += fGl;
    } else if (uGen == u) {
        total += fGl;
        total += a.frac[PU + (size_t)p * a.U + uOther];
    } else {
        total += a.frac[PU + (size_t)p * a.U + uOther];
        total += fGl;
    }
    if (!(fabs(total - 1.0) <= kEpsD)) {
        const double diff = total - 1.0;
        if (fG < diff - kEpsD) {
            atomicOr(a.err, kErrActionFraction);
            return;
        }
        fG = (fG - diff > 0) ? fG - diff : 0.0;
    }
}

__global__ void hbkActLandCoverChange(const ActArgs a, int u, double fraction) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.P) changeGlacierFraction(a, p, u, fraction);
}

// ActionGlacierSnowToIceTransformation::Apply (ActionGlacierSnowToIceTransformation.cpp:41-74)
__global__ void hbkActSnowToIce(const ActArgs a, int cover) {  // cover: 0 the first glacier cover, 1 the second (twins)
    const long long n = (long long)a.P * a.U;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / a.U), u = (int)(i % a.U);
        if (!(a.hruFlags[u] & 1)) continue;
        if (((a.hruFlags[u] >> 3) & 1) != cover) continue;
        if (nearlyZero(a.frac[(size_t)a.P * a.U + i], kPrecision)) continue;
        double& snow = stateAt(a, HBK_S_GLACIER_SNOW, p, u);
        if (snow <= 0) continue;
        if (a.iceInfinite) {
            atomicOr(a.err, kErrActionInfiniteIce);
            continue;
        }
        double& ice = stateAt(a, HBK_S_ICE, p, u);
        ice = ice + snow;
        snow = 0;
    }
}

// ActionGlacierEvolutionDeltaH::Apply (ActionGlacierEvolutionDeltaH.cpp:79-142): one thread per parameter set.
// Step 1 (:90-104): water equivalent of the glacier over the units of the lookup table.
__device__ double deltaHGlacierWE(const ActArgs& a, int p, int n, const int* units) {
    double glacierWE = 0.0;
    for (int i = 0; i < n; ++i) {
        const int u = units[i];
        const double f = a.frac[(size_t)a.P * a.U + (size_t)p * a.U + u];
        if (nearlyZero(f, kPrecision)) continue;
        const double area = f * a.hru[1 * a.ldu + u];
        glacierWE += area * stateAt(a, HBK_S_ICE, p, u);
    }
    return glacierWE;
}
// Step 2 (:106-142): retreat -> table row -> new extents, ice redistributed by the row's volume shares. rowSums: the
// volume sum of every row over the WHOLE table when this engine only holds some of its columns (unit shard), else nullptr.
__device__ void deltaHApply(const ActArgs& a, int p, int n, const int* units, int nRows, const double* areas,
                            const double* volumes, const double* rowSums, double initialWE, int* lastRow, double glacierWE) {
    double retreat = (initialWE - glacierWE) / initialWE;
    retreat = (0.0 < retreat) ? retreat : 0.0;
    const double rowInc = 1.0 / (nRows - 1);
    int row = (int)(retreat / rowInc);
    row = min(row, nRows - 1);
    if (row != lastRow[p]) {
        for (int i = 0; i < n; ++i) {
            const int u = units[i];
            double fraction = areas[(size_t)row * n + i] / a.hru[1 * a.ldu + u];
            if (nearlyZero(fraction, kPrecision)) {  // Action::CheckLandCoverAreaFraction (Action.cpp:71-91)
                fraction = 0.0;
            } else if (fabs(1 - fraction) <= kPrecision) {
                fraction = 1.0;
            }
            changeGlacierFraction(a, p, u, fraction);
        }
        lastRow[p] = row;
    }
    double rowVolumeSum = 0;
    if (rowSums) {
        rowVolumeSum = rowSums[row];
    } else {
        for (int i = 0; i < n; ++i) rowVolumeSum += volumes[(size_t)row * n + i];
    }
    for (int i = 0; i < n; ++i) {
        const int u = units[i];
        const double areaGlacier = areas[(size_t)row * n + i];
        double& ice = stateAt(a, HBK_S_ICE, p, u);
        if (nearlyZero(areaGlacier, kPrecision)) {
            ice = 0;
        } else {
            const double volumeRatio = volumes[(size_t)row * n + i] / rowVolumeSum;
            ice = glacierWE * volumeRatio / areaGlacier;
        }
    }
}
__global__ void hbkActDeltaH(const ActArgs a, int n, const int* units, int nRows, const double* areas,
                             const double* volumes, double initialWE, int* lastRow) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    deltaHApply(a, p, n, units, nRows, areas, volumes, nullptr, initialWE, lastRow, deltaHGlacierWE(a, p, n, units));
}
// Unit shards: this block's part of the water equivalent; the sum over the shards comes back through the caller's
// reduction (hbk_set_shard_reduce) before the second kernel applies the row of the whole table.
__global__ void hbkActDeltaHSum(const ActArgs a, int n, const int* units, double* we) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.P) we[p] = deltaHGlacierWE(a, p, n, units);
}
__global__ void hbkActDeltaHApply(const ActArgs a, int n, const int* units, int nRows, const double* areas,
                                  const double* volumes, const double* rowSums, double initialWE, int* lastRow,
                                  const double* we) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.P) deltaHApply(a, p, n, units, nRows, areas, volumes, rowSums, initialWE, lastRow, we[p]);
}

// ActionGlacierEvolutionAreaScaling::Apply (ActionGlacierEvolutionAreaScaling.cpp:75-123): one thread per
// (parameter set, unit of the lookup table); the units are independent.
__global__ void hbkActAreaScaling(const ActArgs a, int n, const int* units, int nRows, const double* areas,
                                  const double* volumes) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)a.P * n) return;
    const int p = (int)(idx / n), i = (int)(idx % n);
    const int u = units[i];
    const double f = a.frac[(size_t)a.P * a.U + (size_t)p * a.U + u];
    if (nearlyZero(f, kPrecision)) return;
    const double unitArea = a.hru[1 * a.ldu + u];
    const double area = f * unitArea;
    double& ice = stateAt(a, HBK_S_ICE, p, u);
    const double iceVolume = area * ice;
    if (nearlyZero(iceVolume, kPrecision)) {
        changeGlacierFraction(a, p, u, 0.0);
        return;
    }
    const double initialWE = volumes[i] * 900.0;  // _tableVolume.row(0) * constants::iceDensity
    double retreat = (initialWE - iceVolume) / initialWE;
    retreat = (0.0 < retreat) ? retreat : 0.0;
    const double rowInc = 1.0 / (nRows - 1);
    int row = (int)(retreat / rowInc);
    row = min(row, nRows - 1);
    const double newArea = areas[(size_t)row * n + i];
    double fraction = newArea / unitArea;
    if (nearlyZero(fraction, kPrecision)) {  // Action::CheckLandCoverAreaFraction (Action.cpp:71-91)
        fraction = 0.0;
    } else if (fabs(1 - fraction) <= kPrecision) {
        fraction = 1.0;
    }
    changeGlacierFraction(a, p, u, fraction);
    ice = iceVolume / newArea;  // +inf when the table's area is 0, as in the reference
}

// Sum, per parameter set, of the stale instantaneous-flux amounts of glacier bricks that are currently null (zero
// area): they no longer deposit water, but FluxToBrickInstantaneous::GetRealAmount() still reports their last
// amount to the sub-basin stores' constraint test (WaterContainer.cpp:98-101, Processor.cpp:258-260). One block
// per parameter set; only launched when actions can make a glacier vanish.
__global__ void hbkStaleDeposits(const ActArgs a, double* stale) {
    const int p = blockIdx.x;
    __shared__ double s1[128], s2[128];
    double v1 = 0, v2 = 0;
    for (int u = threadIdx.x; u < a.U; u += blockDim.x) {
        if (!(a.hruFlags[u] & 1)) continue;
        if (!nearlyZero(a.frac[(size_t)a.P * a.U + (size_t)p * a.U + u], kPrecision)) continue;
        v1 += stateAt(a, HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT, p, u);
        v2 += stateAt(a, HBK_S_AMT_DEPOSIT_ICEMELT, p, u);
    }
    s1[threadIdx.x] = v1;
    s2[threadIdx.x] = v2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0, t2 = 0;
        for (int i = 0; i < blockDim.x; ++i) {
            t1 += s1[i];
            t2 += s2[i];
        }
        stale[2 * p] = t1;
        stale[2 * p + 1] = t2;
    }
}

__global__ void hbkBroadcastFractions(double* dst, const double* init, int P, int U) {
    const long long n = (long long)P * U;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        dst[i] = init[i % U];
        dst[n + i] = init[U + i % U];
    }
}

// ---- batched objective functions -----------------------------------------------------------------------
// One block per parameter set; two passes over its outlet row (means, then centred sums), fixed-order tree
// reductions. obsStats = {n valid, mean_o, sum (o-mean_o)^2}.
__device__ __forceinline__ double blockSum(double v, double* sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__global__ void hbkObsStats(const double* obs, int t0, int T, double* stats) {
    __shared__ double sh[256];
    double n = 0, s = 0;
    for (int t = t0 + threadIdx.x; t < T; t += blockDim.x) {
        const double o = obs[t];
        if (o == o) {
            n += 1;
            s += o;
        }
    }
    n = blockSum(n, sh);
    s = blockSum(s, sh);
    const double mean = s / n;
    double ss = 0;
    for (int t = t0 + threadIdx.x; t < T; t += blockDim.x) {
        const double o = obs[t];
        if (o == o) ss += (o - mean) * (o - mean);
    }
    ss = blockSum(ss, sh);
    if (threadIdx.x == 0) {
        stats[0] = n;
        stats[1] = mean;
        stats[2] = ss;
    }
}

__global__ void hbkScores(const double* __restrict__ outlet, long long ldt, const double* __restrict__ obs, int t0, int T,
                          const double* __restrict__ stats, int metric, double* __restrict__ out) {
    __shared__ double sh[256];
    const int p = blockIdx.x;
    const double* q = outlet + (size_t)p * ldt;
    const double n = stats[0], mo = stats[1], sso = stats[2];
    double a = 0, b = 0, ab = 0;
    for (int t = t0 + threadIdx.x; t < T; t += blockDim.x) {
        const double o = obs[t];
        if (o == o) {
            const double s = q[t];
            a += s;
            b += (s - o) * (s - o);
            ab += fabs(s - o);
        }
    }
    const double sumS = blockSum(a, sh);
    const double sse = blockSum(b, sh);
    const double sae = blockSum(ab, sh);
    const double ms = sumS / n;
    switch (metric) {  // metrics of the first pass
        case HBK_SCORE_NSE:
            if (threadIdx.x == 0) out[p] = 1 - sse / sso;
            return;
        case HBK_SCORE_ME:
            if (threadIdx.x == 0) out[p] = ms - mo;  // mean(s - o) over the same steps
            return;
        case HBK_SCORE_MAE:
            if (threadIdx.x == 0) out[p] = sae / n;
            return;
        case HBK_SCORE_MSE:
            if (threadIdx.x == 0) out[p] = sse / n;
            return;
        case HBK_SCORE_RMSE:
            if (threadIdx.x == 0) out[p] = sqrt(sse / n);
            return;
        case HBK_SCORE_NRMSE_MEAN:
            if (threadIdx.x == 0) out[p] = sqrt(sse / n) / mo;
            return;
        default:
            break;
    }
    double c = 0, d = 0;
    for (int t = t0 + threadIdx.x; t < T; t += blockDim.x) {
        const double o = obs[t];
        if (o == o) {
            const double ds = q[t] - ms;
            c += ds * ds;
            d += ds * (o - mo);
        }
    }
    const double sss = blockSum(c, sh);
    const double sso2 = blockSum(d, sh);
    if (threadIdx.x == 0) {
        const double r = sso2 / sqrt(sss * sso);
        const double be = ms / mo;
        if (metric == HBK_SCORE_R_SQUARED) {
            out[p] = r * r;
        } else if (metric == HBK_SCORE_KGE_2009) {
            const double al = sqrt(sss / (n - 1)) / sqrt(sso / (n - 1));
            out[p] = 1 - sqrt((r - 1) * (r - 1) + (al - 1) * (al - 1) + (be - 1) * (be - 1));
        } else {
            const double cvS = sqrt(sss / (n - 1)) / ms;
            const double cvO = sqrt(sso / (n - 1)) / mo;
            const double g = cvS / cvO;
            out[p] = 1 - sqrt((r - 1) * (r - 1) + (g - 1) * (g - 1) + (be - 1) * (be - 1));
        }
    }
}

// FP64 roofline denominator: 8 independent DFMA chains per thread (2 flops per DFMA).
__global__ void hbkFp64Peak(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

// sqrtUnit against sqrt() on n values spread over (2^-40, 1] (and 1 itself): counts the values whose bits differ.
__global__ void hbkSqrtCheck(long long n, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;  // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double m = 1.0 + (double)(z >> 12) * (1.0 / 4503599627370496.0);  // [1, 2)
        const int ex = (int)(z & 0xfff) % 41;                                       // 2^-1 .. 2^-41
        double x = ldexp(m, -1 - ex);
        if (i == 0) x = 1.0;
        if (__double_as_longlong(sqrtUnit(x)) != __double_as_longlong(sqrt(x))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace hbk

// =====================================================================================================
// Host side
// =====================================================================================================
using namespace hbk;

struct hbk_engine {
    hbk_structure st{};
    int U = 0, T = 0, P = 0, device = 0;
    int PB = 32, UB = 8, TC = 16, nUTiles = 1, G = 1, nGroups = 1;
    bool clusterMode = false;
    bool stateLaneP = false;  // state layout: element (p,u) at u*P + p (parameter axis on the lanes) or p*U + u
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string error;
    // device buffers
    float* dParams = nullptr;
    int ldp = 0;
    double* dHru = nullptr;
    int* dHruFlags = nullptr;
    int ldu = 0;
    double* dFrac = nullptr;      // [2][U] (per unit) or [2][P*U]
    double* dFracInit = nullptr;  // initial fractions [2][U]
    int fracPerSet = 0;
    double* dForcingRaw[HBK_N_VARS] = {};    // [T][U]
    double* dForcingTiled[HBK_N_VARS] = {};  // [nUTiles][tld][UB]
    double* dStation[HBK_N_VARS] = {};       // [tld]
    bool haveRaw[HBK_N_VARS] = {}, haveStation[HBK_N_VARS] = {};
    int nVars() const { return st.melt_kind == HBK_MELT_TEMPERATURE_INDEX ? 4 : 3; }
    // state slots in use: the six of the snowpacks' liquid water only with HBK_RETENTION_*
    int nStates() const { return st.snow_retention ? HBK_N_STATES : HBK_S_AMT_REFREEZE_GROUND; }
    std::vector<int> hAspect;  // HBK_ASPECT_* per unit (hbk_set_unit_aspect); empty = not given
    std::vector<int> hTwinOf;  // unit a twin (second glacier cover) belongs to, -1 = ordinary unit; empty = no twins
    int* dTwinOf = nullptr;    // [U] device copies for the action kernels (nullptr without twins)
    int* dTwinCol = nullptr;   // [U] twin of a unit, -1 = none
    bool hasTwins = false;
    // bit 0: the unit carries the glacier bricks; bits 1-2: HBK_ASPECT_*; bit 3: twin (second glacier cover)
    int unitFlags(int u) const {
        return (hGlacierVariant.empty() ? 0 : hGlacierVariant[u]) | ((hAspect.empty() ? 0 : hAspect[u]) << 1) |
               ((!hTwinOf.empty() && hTwinOf[u] >= 0) ? 8 : 0);
    }
    // transport:snow_slide (hbk_set_lateral_connections)
    std::vector<int> hLatOff, hLatRecv;
    std::vector<double> hLatWeight;
    std::vector<float> hSlopeDeg;
    std::vector<float> hParamsT;  // [HBK_N_PARAMS][ldp], the device layout (internal set order)
    float* hParamStage = nullptr;  // page-locked staging of the same layout for hbk_set_parameters
    size_t paramStageN = 0;
    int* dLatOff = nullptr;
    int* dLatRecv = nullptr;
    double* dLatWeight = nullptr;
    double* dLatThr = nullptr;   // [U][ldp]
    double* dLatPend = nullptr;  // [2][U][P]
    bool latThrDirty = true;
    int tiledUB = 0;
    double* dGradT = nullptr;
    double* dGradP = nullptr;
    double* dCorrP = nullptr;
    int perSetDroppedVars = 0;   // bit var: that variable's per-set arrays were freed
    bool perSetDropped = false;  // per-set gradients / corrections were configured, then freed by a change of n_sets
    double gradTs = 0, gradPs = 0, corrPs = 1, zrefT = 0, zrefP = 0;
    long long tld = 0;
    double* dState = nullptr;      // [HBK_N_STATES][P*U]
    double* dInitUnit = nullptr;   // [HBK_N_STATES][U] initial state per unit (broadcast over sets)
    bool initPerSet = false;
    double* dInitFull = nullptr;   // [HBK_N_STATES][P*U] after save_as_initial_state
    double* dBasinState = nullptr; // [P][2]
    double* dBasinInit = nullptr;
    double* dOutlet = nullptr;     // [P][ldt]: the buffer the current / last run writes (one of dOutletBuf)
    // hbk_get_outlet_async: the outlet of run i travels to the host on its own stream while run i+1 computes into the
    // other buffer
    double* dOutletBuf[2] = {nullptr, nullptr};
    int outletCur = 0;
    bool copyPending[2] = {false, false};
    cudaStream_t copyStream = nullptr;
    cudaEvent_t copyDone[2] = {nullptr, nullptr};
    double* dPartQ = nullptr;      // [nGroups][P][ldt] (nGroups > 1)
    int partGroups = 0;
    double* dD1 = nullptr;         // [P][ldt]
    double* dD2 = nullptr;
    double* dPartD1 = nullptr;
    double* dPartD2 = nullptr;
    long long ldt = 0;
    uint64_t recordMask = 0;
    double* dRec = nullptr;
    size_t recBytes = 0;
    int* dErr = nullptr;
    unsigned long long* dUnconv = nullptr;
    int* dUnconvSet = nullptr;  // [P]
    // work queue of a unit-kernel launch: [0] item counter, [1 + vb] slices done per virtual block; one per launch of a
    // run segment (the launches of a split glacier structure run concurrently, each on its own stream)
    static constexpr int kMaxRuns = 6;
    int* dQueue[kMaxRuns] = {};
    size_t queueInts[kMaxRuns] = {};
    cudaStream_t auxStream[kMaxRuns] = {};  // [0] unused: the first launch runs on `stream`
    cudaEvent_t evFork = nullptr, evJoin[kMaxRuns] = {};
    int spinupSteps = 0;
    bool unitsSet = false, ran = false;
    // dated actions
    struct Event {
        int step, kind, unit;  // kind 0 land-cover change, 1 snow->ice, 2 delta-h, 3 area scaling
        double value;
    };
    std::vector<Event> events;
    double* dFracSet = nullptr;  // [2][P*U]
    int* dDhUnits = nullptr;
    double* dDhAreas = nullptr;
    double* dDhVolumes = nullptr;
    int* dDhLastRow = nullptr;
    double* dSwatFactor = nullptr;  // [T] day-of-year term of transform:snow_ice_swat
    int* dPerm = nullptr;       // [P] internal index -> caller's index; nullptr = identity
    std::vector<int> perm;
    double* dScores = nullptr;  // [P]
    double* dObs = nullptr;     // [T] + 3 stats
    double* dStale = nullptr;  // [P][2] stale deposit amounts of null glacier bricks (actions only)
    int dhUnits = 0, dhRows = 0;
    std::vector<int> hDhUnits;  // the lookup tables' units (host copy: they may gain ice at an update)
    // Unit tiles of a glacier structure whose units can never carry ice (glacier fraction 0, untouched by any action) run
    // the plain kernel variant (80 instead of 128 registers): the launches of one run segment (planRuns)
    struct TileRun {
        int tile0, tile1;  // unit tiles [tile0, tile1)
        bool glacier;      // kernel variant
        int TC, G, nGroups, grpBase;
    };
    std::vector<TileRun> runs;
    int totalGroups = 1, glacierGroups = 1;
    bool dhTablesSet = false;  // since the last hbk_clear_actions (dhUnits may be 0 on a unit shard)
    double dhInitialWE = 0;
    std::vector<double> hArea;
    size_t allocP = 0;
    // stats
    double lastMs = 0;
    int lastLaunches = 0;
    long long lastUnconverged = 0;
    std::vector<double> hFracG, hFracGl;
    std::vector<int> hGlacierVariant;
    // unit sharding (hbk_set_unit_shard): this engine holds a block of the basin's units; hbk_run stops after the
    // unit pass and hbk_finish_sharded_run integrates the sub-basin stores on the reduced deposit sums
    bool shard = false, shardPending = false;
    double shardBasinArea = 0;
    // sums over all shards inside hbk_run (hbk_set_shard_reduce): spin-up, delta-h dates, end of run
    hbk_shard_reduce_fn shardReduce = nullptr;
    void* shardUser = nullptr;
    double* dShardStage = nullptr;  // contiguous staging of the column blocks handed to the callback
    size_t shardStageN = 0;
    double* dStaleSeg = nullptr;    // [segment][P][2]: stale deposit amounts of null glacier bricks per segment
    size_t staleSegN = 0;
    double* dShardWE = nullptr;     // [P] glacier water equivalent of this block at a delta-h date
    double* dDhRowSums = nullptr;   // [dhRows] volume sums of the whole table's rows (hbk_set_delta_h_shard_totals)
    bool dhShardTotals = false;
    struct Segment {
        int tBegin, tEnd;
    };
    std::vector<Segment> shardSegments;  // segments of the current run (shard mode with a callback)
};

namespace {

int fail(hbk_engine* e, int code, const std::string& msg) {
    if (e) e->error = msg;
    return code;
}

#define HBK_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _err = (call);                                                                  \
        if (_err != cudaSuccess)                                                                    \
            return fail(e, HBK_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));      \
    } while (0)

template <class T>
void freeDev(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// Thread-block shape: parameter axis on the lanes when there are enough sets, else the unit axis.
int chunkSteps(const hbk_engine* e, bool glacier);
int tilesPerBlock(const hbk_engine* e, int nTiles, bool glacier);

void chooseShape(hbk_engine* e) {
    int pb = 1;
    while (pb < 32 && pb < e->P) pb *= 2;
    e->PB = pb;
    // units per block: 128..256 threads; pick the width that leaves the fewest idle unit slots in the last
    // tile (U = 50 with 8 units/block wastes 6 of 56 warps; 5 units/block wastes none)
    const int ubMax = kBlock / pb, ubMin = std::max(1, ubMax / 2);
    int best = ubMax;
    double bestUse = 0;
    for (int ub = ubMax; ub >= ubMin; --ub) {
        if ((ub * pb) % 32 != 0 || (e->haveRaw[0] && (ub % 2 != 0))) continue;
        const int tiles = (e->U + ub - 1) / ub;
        const double use = (double)e->U / ((double)tiles * ub);
        // prefer wide blocks: fewer resident blocks keep the time-multiplexed state of all resident blocks inside
        // the 126 MB L2 (888 blocks x 32 sets x 50 units x 8 slots x 8 B = 90 MB); a narrower block must buy
        // more than 5 % of unit-slot utilisation to win
        if (use > bestUse + 0.05) {
            bestUse = use;
            best = ub;
        }
    }
    e->UB = best;
    e->nUTiles = (e->U + e->UB - 1) / e->UB;
    e->G = tilesPerBlock(e, e->nUTiles, e->st.has_glacier != 0);
    e->nGroups = (e->nUTiles + e->G - 1) / e->G;
    // Optional (HBK_CLUSTER=1): one thread-block cluster per parameter-set group when all unit tiles fit in one
    // cluster (<= 16 CTAs, opt-in size): state stays in registers, the outlet is reduced through distributed shared
    // memory, DRAM traffic = algorithmic. Measured 11 % SLOWER than time-multiplexing on C2 (cluster.sync of 52 warps
    // every chunk + 13-CTA co-scheduling, profiles/r01_tuning.md), so it is not the default.
    const char* wantCluster = std::getenv("HBK_CLUSTER");
    e->clusterMode = wantCluster && wantCluster[0] == '1' && e->PB == 32 && e->nUTiles >= 2 && e->nUTiles <= 16 &&
                     !std::getenv("HBK_TILES_PER_BLOCK");
    if (e->clusterMode) {
        e->G = 1;
        e->nGroups = 1;
    }
    e->stateLaneP = e->PB == 32;
    if (e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE) {
        // hbkRunLateral: one thread per parameter set, the sets of a warp side by side in the state arrays; no unit tiles
        e->stateLaneP = true;
        e->G = e->nUTiles;
        e->nGroups = 1;
        e->clusterMode = false;
    }
    e->TC = chunkSteps(e, e->st.has_glacier != 0);
}

size_t smemBytes(const hbk_engine* e, bool tiled, bool glacier, int TC) {
    const int nq = glacier ? 3 : 1;
    const size_t fRow = tiled ? (size_t)TC * e->UB : (size_t)TC + 2;
    return 128 + 2 * e->nVars() * fRow * 8 + (size_t)nq * TC * (e->PB * e->UB + (e->clusterMode ? 2 * e->PB : 0)) * 8 +
           (HBK_PAR_SHARED ? (size_t)hbk::kParFields * e->PB * 8 : 0);
}

// Steps per chunk of a kernel variant: shared memory per block small enough for its resident blocks per SM
int chunkSteps(const hbk_engine* e, bool glacier) {
    const int minBlocks = hbk::hbkMinBlocks(e->st.n_soil_stores, glacier);
    const int nq = glacier ? 3 : 1;
    const size_t budget = (size_t)(220 * 1024) / minBlocks - 1024;
    // (measured: the shared-memory carve-out shrinks the L1 that serves the register spills; with one block
    // barrier per chunk, 16 steps per chunk is the optimum on C2: 4 -> 2.07e10, 8 -> 2.14e10, 16 -> 2.18e10,
    // profiles/r01_tuning.md)
    int tc = hbk::hbkEnsTC(glacier);
    for (;;) {
        const size_t fRow = (e->haveRaw[0] ? (size_t)tc * e->UB : (size_t)tc + 2);
        const size_t bytes = 128 + 2 * e->nVars() * fRow * 8 + (size_t)nq * tc * (e->PB * e->UB + 2 * e->PB) * 8;
        if (bytes <= budget || tc == 2) break;
        tc /= 2;
    }
    if (const char* force = std::getenv("HBK_CHUNK_STEPS")) {  // tuning knob
        const int v = std::atoi(force);
        // station series are padded for chunks of up to 16 steps (tld = ceil16(T) + 32 >= T + TC + 2)
        if (v >= 2 && v <= (e->haveRaw[0] ? 64 : 16) && (v & (v - 1)) == 0) tc = v;
    }
    return tc;
}

// Tiles per block for `nTiles` unit tiles run by a kernel variant: enough blocks to fill the machine several times,
// otherwise as many tiles per block as possible (G = all tiles: the block sums every unit of its sets itself)
int tilesPerBlock(const hbk_engine* e, int nTiles, bool glacier) {
    const long long nPGroups = (e->P + e->PB - 1) / e->PB;
    const int minBlocks = hbk::hbkMinBlocks(e->st.n_soil_stores, glacier);
    const long long wantBlocks = 148LL * minBlocks * 2;
    const long long groups = std::min<long long>(nTiles, std::max<long long>(1, (wantBlocks + nPGroups - 1) / nPGroups));
    int G = (int)((nTiles + groups - 1) / groups);
    if (const char* force = std::getenv("HBK_TILES_PER_BLOCK")) {  // testing / tuning knob
        const int g = std::atoi(force);
        if (g > 0) G = std::min(g, nTiles);
    }
    return G;
}

// The launches of a run segment. A glacier structure compiles the glacier variant of the unit kernel (three outlet
// quantities, the glacier cover's state and parameters live: 128 registers, 4 blocks per SM), and that variant costs the
// same on a unit without ice (measured on C3 / C4: 46 % / 36 % more than the plain variant on the same units). Units whose
// glacier cover is empty and can never gain area — fraction exactly 0 (ground exactly 1), in no delta-h table, never
// named by a land-cover change — compute exactly what the plain variant computes (LandCover.h:91-93: a null brick is
// skipped), so runs of unit tiles that hold only such units are launched with the plain variant; their deposits into
// the sub-basin stores are zero. Every launch writes partial outlet series per tile group, summed in group order by
// hbkReduceTiles. Not with a recorder (the plain variant has no glacier labels), clusters, or the lateral kernel.
void planRuns(hbk_engine* e) {
    e->runs.clear();
    const bool glacier = e->st.has_glacier != 0;
    auto single = [&]() {
        e->runs.push_back({0, e->nUTiles, glacier, e->TC, e->G, e->nGroups, 0});
        e->totalGroups = e->nGroups;
        e->glacierGroups = glacier ? e->nGroups : 0;
    };
    const char* knob = std::getenv("HBK_SPLIT_GLACIER");
    if (!glacier || e->clusterMode || e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE || e->recordMask != 0 ||
        (knob && knob[0] == '0') || HBK_PAR_SHARED)
        return single();
    std::vector<char> ice(e->U, 0);  // the unit may carry ice at some time of the run
    for (int u = 0; u < e->U; ++u)
        ice[u] = e->hGlacierVariant[u] && !(e->hFracGl[u] == 0.0 && e->hFracG[u] == 1.0);
    for (const auto& ev : e->events)
        if (ev.kind == 0 && ev.unit >= 0 && ev.unit < e->U) ice[ev.unit] = 1;
    for (int u : e->hDhUnits)
        if (u >= 0 && u < e->U) ice[u] = 1;
    if (e->hasTwins)
        for (int u = 0; u < e->U; ++u)
            if (e->hTwinOf[u] >= 0) ice[u] = ice[e->hTwinOf[u]] = 1;
    std::vector<char> tileIce(e->nUTiles, 0);
    for (int u = 0; u < e->U; ++u)
        if (ice[u]) tileIce[u / e->UB] = 1;
    std::vector<hbk_engine::TileRun> runs;
    for (int t = 0; t < e->nUTiles; ++t) {
        if (runs.empty() || runs.back().glacier != (tileIce[t] != 0)) runs.push_back({t, t, tileIce[t] != 0, 0, 0, 0, 0});
        runs.back().tile1 = t + 1;
    }
    bool anyPlain = false;
    for (const auto& r : runs) anyPlain = anyPlain || !r.glacier;
    // a handful of launches at most (glacier units sit together in an elevation-sorted basin). A basin — or a unit shard of
    // one — without any ice runs the plain variant alone: its deposits into the sub-basin stores are zeroed by hbk_run
    if (!anyPlain || (int)runs.size() > hbk_engine::kMaxRuns) return single();
    e->glacierGroups = 0;
    int base = 0;
    for (int pass = 0; pass < 2; ++pass)  // glacier groups first: the deposits are summed over those only
        for (auto& r : runs) {
            if (r.glacier != (pass == 0)) continue;
            r.TC = r.glacier ? e->TC : chunkSteps(e, false);
            r.G = tilesPerBlock(e, r.tile1 - r.tile0, r.glacier);
            r.nGroups = (r.tile1 - r.tile0 + r.G - 1) / r.G;
            r.grpBase = base;
            base += r.nGroups;
            if (pass == 0) e->glacierGroups = base;
        }
    e->totalGroups = base;
    e->runs = runs;
}

using KernelFn = void (*)(const KArgs);

template <int SOLVER, int NSOIL, bool GLACIER>
KernelFn pickDaily(bool daily, int shape) {
    // the sequential solvers (SOLVER 3): the generic instantiation (dt read at run time) and, for calibration ensembles with
    // daily steps — crank_nicolson is the default of the reference's Python layer —, the compile-time ensemble shape
    if constexpr (SOLVER == 3) {
        if (daily && shape == 1) return (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, true, 1>;
        return (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, false, 0>;
    } else {
        // the compile-time shapes exist for daily steps only (the case ensembles and distributed basins are run in)
        if (daily && shape == 1) return (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, true, 1>;
        if (daily && shape == 2) return (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, true, 2>;
        return daily ? (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, true, 0>
                     : (KernelFn)hbkRunUnits<SOLVER, NSOIL, GLACIER, false, 0>;
    }
}
template <int SOLVER, int NSOIL>
KernelFn pickGlacier(bool glacier, bool daily, int shape) {
    return glacier ? pickDaily<SOLVER, NSOIL, true>(daily, shape) : pickDaily<SOLVER, NSOIL, false>(daily, shape);
}
template <int SOLVER>
KernelFn pickSoil(int nSoil, bool glacier, bool daily, int shape) {
    return nSoil == 2 ? pickGlacier<SOLVER, 2>(glacier, daily, shape) : pickGlacier<SOLVER, 1>(glacier, daily, shape);
}
// "Device functions selected from a per-structure process table": the table (hbk_structure) picks the
// template instantiation, i.e. which process device functions are compiled into the time loop. ens: the launch has the
// ensemble block shape (32 sets x 4 units, 16-step chunks, station forcing, no cluster) that one instantiation per
// structure fixes at compile time.
KernelFn pickKernel(const hbk_structure& st, bool glacier, int shape) {
    const bool daily = st.time_step_days == 1.0;
    switch (st.solver) {
        case HBK_SOLVER_EULER: return pickSoil<0>(st.n_soil_stores, glacier, daily, shape);
        case HBK_SOLVER_HEUN: return pickSoil<1>(st.n_soil_stores, glacier, daily, shape);
        case HBK_SOLVER_RK4: return pickSoil<2>(st.n_soil_stores, glacier, daily, shape);
        default:  // sequential family: the compile-time ensemble shape exists for crank_nicolson only
            return pickSoil<3>(st.n_soil_stores, glacier, daily, (shape == 1 && st.solver == HBK_SOLVER_CRANK_NICOLSON) ? 1 : 0);
    }
}

template <int SOLVER, int NSOIL>
void lateralGlacier(bool glacier, int blocks, cudaStream_t stream, const KArgs& a, const LatArgs& L) {
    if (glacier) {
        hbkRunLateral<SOLVER, NSOIL, true><<<blocks, 128, 0, stream>>>(a, L);
    } else {
        hbkRunLateral<SOLVER, NSOIL, false><<<blocks, 128, 0, stream>>>(a, L);
    }
}
template <int SOLVER>
void lateralSoil(int nSoil, bool glacier, int blocks, cudaStream_t stream, const KArgs& a, const LatArgs& L) {
    if (nSoil == 2) {
        lateralGlacier<SOLVER, 2>(glacier, blocks, stream, a, L);
    } else {
        lateralGlacier<SOLVER, 1>(glacier, blocks, stream, a, L);
    }
}

// TimeMachine::GetCurrentDayOfYear (TimeMachine.cpp:87-98) of the civil date of floor(mjd).
int dayOfYear(double mjd) {
    if (mjd <= 0) return 1;
    const long long z = (long long)std::floor(mjd) - 40587LL + 719468LL;  // days since 0000-03-01
    const long long era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const unsigned doyMarch = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doyMarch + 2) / 153;
    const unsigned day = doyMarch - (153 * mp + 2) / 5 + 1;
    const unsigned month = mp < 10 ? mp + 3 : mp - 9;
    const long long year = (long long)yoe + era * 400 + (month <= 2);
    static const int cumMonthDays[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    const bool leap = (year % 4 == 0 && (year % 100 != 0 || year % 400 == 0));
    int doy = cumMonthDays[month - 1] + (int)day;
    if (leap && month > 2) doy += 1;
    return doy;
}

int ensureSetBuffers(hbk_engine* e, int P) {
    if ((size_t)P == e->allocP) return HBK_OK;
    freeDev(e->dParams);
    freeDev(e->dState);
    freeDev(e->dInitFull);
    freeDev(e->dBasinState);
    freeDev(e->dBasinInit);
    if (e->copyStream) cudaStreamSynchronize(e->copyStream);
    freeDev(e->dOutletBuf[0]);
    freeDev(e->dOutletBuf[1]);
    e->dOutlet = nullptr;
    e->outletCur = 0;
    e->copyPending[0] = e->copyPending[1] = false;
    freeDev(e->dPartQ);
    freeDev(e->dD1);
    freeDev(e->dD2);
    freeDev(e->dPartD1);
    freeDev(e->dPartD2);
    freeDev(e->dRec);
    if (e->dGradT) e->perSetDroppedVars |= 1 << HBK_VAR_TEMPERATURE;
    if (e->dGradP || e->dCorrP) e->perSetDroppedVars |= 1 << HBK_VAR_PRECIPITATION;
    if (e->perSetDroppedVars) e->perSetDropped = true;  // hbk_run refuses until they are set again
    freeDev(e->dGradT);
    freeDev(e->dGradP);
    freeDev(e->dCorrP);
    freeDev(e->dFracSet);
    freeDev(e->dDhLastRow);
    freeDev(e->dStale);
    freeDev(e->dLatThr);
    freeDev(e->dLatPend);
    e->latThrDirty = true;
    freeDev(e->dStaleSeg);
    e->staleSegN = 0;
    freeDev(e->dShardWE);
    freeDev(e->dScores);
    freeDev(e->dPerm);
    freeDev(e->dUnconvSet);
    e->perm.clear();
    e->recBytes = 0;
    e->initPerSet = false;
    e->P = P;
    e->ldp = (P + 31) / 32 * 32;
    e->ldt = (e->T + 15) / 16 * 16;
    HBK_CUDA(cudaMalloc(&e->dParams, sizeof(float) * HBK_N_PARAMS * e->ldp));
    HBK_CUDA(cudaMalloc(&e->dState, sizeof(double) * e->nStates() * (size_t)P * e->U));
    HBK_CUDA(cudaMalloc(&e->dBasinState, sizeof(double) * 2 * P));
    HBK_CUDA(cudaMalloc(&e->dBasinInit, sizeof(double) * 2 * P));
    HBK_CUDA(cudaMemsetAsync(e->dBasinInit, 0, sizeof(double) * 2 * P, e->stream));
    HBK_CUDA(cudaMalloc(&e->dOutletBuf[0], sizeof(double) * (size_t)P * e->ldt));
    e->dOutlet = e->dOutletBuf[0];
    HBK_CUDA(cudaMalloc(&e->dUnconvSet, sizeof(int) * P));
    e->allocP = P;
    return HBK_OK;
}

}  // namespace

extern "C" {

int hbk_engine_create(const hbk_structure* st, int32_t n_units, int32_t n_steps, int32_t device, hbk_engine** out) {
    if (!st || !out || n_units <= 0 || n_steps <= 0) return HBK_E_ARG;
    *out = nullptr;
    auto* e = new hbk_engine();
    e->st = *st;
    e->U = n_units;
    e->T = n_steps;
    e->device = device;
    auto bail = [&](int code, const std::string& msg) {
        std::fprintf(stderr, "hbk_engine_create: %s\n", msg.c_str());
        delete e;
        return code;
    };
    if (st->solver < 0 || st->solver > HBK_SOLVER_ANALYTIC_LINEAR) return bail(HBK_E_ARG, "unknown solver");
    if (st->n_soil_stores != 1 && st->n_soil_stores != 2) return bail(HBK_E_UNSUPPORTED, "n_soil_stores must be 1 or 2");
    if (!(st->time_step_days > 0)) return bail(HBK_E_ARG, "time_step_days must be > 0");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return bail(HBK_E_CUDA, "no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= count) return bail(HBK_E_ARG, "bad device index");
    if (cudaSetDevice(device) != cudaSuccess) return bail(HBK_E_CUDA, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(HBK_E_CUDA, "cudaStreamCreate failed");
    cudaEventCreate(&e->ev0);
    cudaEventCreate(&e->ev1);
    e->ldu = (n_units + 31) / 32 * 32;
    e->tld = (n_steps + 15) / 16 * 16 + 96;  // >= last chunk start + 64 + 2 (HBK_CHUNK_STEPS <= 64)
    bool ok = cudaMalloc(&e->dHru, sizeof(double) * kNH * e->ldu) == cudaSuccess &&
              cudaMalloc(&e->dHruFlags, sizeof(int) * e->ldu) == cudaSuccess &&
              cudaMalloc(&e->dFrac, sizeof(double) * 2 * n_units) == cudaSuccess &&
              cudaMalloc(&e->dFracInit, sizeof(double) * 2 * n_units) == cudaSuccess &&
              cudaMalloc(&e->dInitUnit, sizeof(double) * HBK_N_STATES * n_units) == cudaSuccess &&
              cudaMalloc(&e->dErr, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&e->dUnconv, sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) return bail(HBK_E_CUDA, "device allocation failed");
    if (cudaMemsetAsync(e->dInitUnit, 0, sizeof(double) * HBK_N_STATES * n_units, e->stream) != cudaSuccess ||
        cudaStreamSynchronize(e->stream) != cudaSuccess)
        return bail(HBK_E_CUDA, "device memset failed");
    if (st->snow_ice_kind < HBK_SNOW_ICE_NONE || st->snow_ice_kind > HBK_SNOW_ICE_SWAT)
        return bail(HBK_E_ARG, "unknown snow_ice_kind");
    if (st->sublimation_kind < HBK_SUBLIMATION_NONE || st->sublimation_kind > HBK_SUBLIMATION_PET)
        return bail(HBK_E_ARG, "unknown sublimation_kind");
    if (st->melt_kind < HBK_MELT_DEGREE_DAY || st->melt_kind > HBK_MELT_TEMPERATURE_INDEX)
        return bail(HBK_E_ARG, "unknown melt_kind");
    if (st->snow_slide_kind < HBK_SNOW_SLIDE_NONE || st->snow_slide_kind > HBK_SNOW_SLIDE_ALL)
        return bail(HBK_E_ARG, "unknown snow_slide_kind");
    if (HBK_PAR_SHARED && st->snow_slide_kind != HBK_SNOW_SLIDE_NONE)
        return bail(HBK_E_UNSUPPORTED, "this build (HBK_PAR_SHARED) has no transport:snow_slide");
    if (HBK_PAR_SHARED && st->melt_kind != HBK_MELT_DEGREE_DAY)
        return bail(HBK_E_UNSUPPORTED, "this build (HBK_PAR_SHARED) only has melt:degree_day");
    if (st->has_glacier && st->snow_ice_kind == HBK_SNOW_ICE_SWAT) {
        // ProcessTransformSnowToIceSwat.cpp:31-39: 1 + sin(2 pi (doy - ref) / 365) per time step, in double like
        // the reference's expression (2.0f * constants::pi promotes to double), doy as TimeMachine.cpp:87-98
        std::vector<double> factor(n_steps);
        const int daysRef = st->north_hemisphere ? 81 : 264;
        const double pi = 3.1415926535897932384626433832795;
        for (int t = 0; t < n_steps; ++t) {
            const int doy = dayOfYear(st->start_mjd + t * st->time_step_days);
            factor[t] = 1 + std::sin(2.0f * pi * (doy - daysRef) / 365.0f);
        }
        if (cudaMalloc(&e->dSwatFactor, sizeof(double) * n_steps) != cudaSuccess ||
            cudaMemcpy(e->dSwatFactor, factor.data(), sizeof(double) * n_steps, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(HBK_E_CUDA, "device allocation failed");
    }
    *out = e;
    return HBK_OK;
}

void hbk_engine_destroy(hbk_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    freeDev(e->dParams);
    freeDev(e->dHru);
    freeDev(e->dHruFlags);
    freeDev(e->dSwatFactor);
    freeDev(e->dFrac);
    freeDev(e->dFracInit);
    for (int v = 0; v < HBK_N_VARS; ++v) {
        freeDev(e->dForcingRaw[v]);
        freeDev(e->dForcingTiled[v]);
        freeDev(e->dStation[v]);
    }
    freeDev(e->dGradT);
    freeDev(e->dGradP);
    freeDev(e->dCorrP);
    freeDev(e->dState);
    freeDev(e->dInitUnit);
    freeDev(e->dInitFull);
    freeDev(e->dBasinState);
    freeDev(e->dBasinInit);
    if (e->copyStream) cudaStreamSynchronize(e->copyStream);
    freeDev(e->dOutletBuf[0]);
    freeDev(e->dOutletBuf[1]);
    if (e->copyDone[0]) cudaEventDestroy(e->copyDone[0]);
    if (e->copyDone[1]) cudaEventDestroy(e->copyDone[1]);
    if (e->copyStream) cudaStreamDestroy(e->copyStream);
    freeDev(e->dPartQ);
    freeDev(e->dD1);
    freeDev(e->dD2);
    freeDev(e->dPartD1);
    freeDev(e->dPartD2);
    freeDev(e->dRec);
    freeDev(e->dFracSet);
    freeDev(e->dDhUnits);
    freeDev(e->dDhAreas);
    freeDev(e->dDhVolumes);
    freeDev(e->dDhLastRow);
    freeDev(e->dStale);
    freeDev(e->dScores);
    freeDev(e->dPerm);
    freeDev(e->dObs);
    freeDev(e->dErr);
    freeDev(e->dUnconv);
    freeDev(e->dUnconvSet);
    for (auto& q : e->dQueue) freeDev(q);
    if (e->hParamStage) cudaFreeHost(e->hParamStage);
    for (auto& st : e->auxStream)
        if (st) cudaStreamDestroy(st);
    for (auto& ev : e->evJoin)
        if (ev) cudaEventDestroy(ev);
    if (e->evFork) cudaEventDestroy(e->evFork);
    freeDev(e->dTwinOf);
    freeDev(e->dTwinCol);
    freeDev(e->dLatOff);
    freeDev(e->dLatRecv);
    freeDev(e->dLatWeight);
    freeDev(e->dLatThr);
    freeDev(e->dLatPend);
    freeDev(e->dShardStage);
    freeDev(e->dStaleSeg);
    freeDev(e->dShardWE);
    freeDev(e->dDhRowSums);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char* hbk_last_error(const hbk_engine* e) { return e ? e->error.c_str() : "null engine"; }

int hbk_set_units(hbk_engine* e, const double* area, const double* elevation, const float* slope,
                  const int32_t* glacier_variant, const double* frac_ground, const double* frac_glacier) {
    if (!e || !area) return HBK_E_ARG;
    if (e->st.quick_kind == HBK_QUICK_SOCONT && !slope)
        return fail(e, HBK_E_ARG, "runoff:socont needs the 'slope' property of every hydro unit");
    HBK_CUDA(cudaSetDevice(e->device));
    const int U = e->U;
    double basinArea = 0;  // SubBasin::AddHydroUnit (SubBasin.cpp:173): running sum in unit order
    for (int u = 0; u < U; ++u)
        if (e->hTwinOf.empty() || e->hTwinOf[u] < 0) basinArea += area[u];  // twins carry a cover of a unit already counted
    if (e->shard) basinArea = e->shardBasinArea;  // a block of a larger basin (hbk_set_unit_shard)
    std::vector<double> hru((size_t)kNH * e->ldu, 0.0);
    std::vector<int> flags(e->ldu, 0);
    e->hArea.assign(area, area + U);
    e->hFracG.assign(U, 1.0);
    e->hFracGl.assign(U, 0.0);
    e->hGlacierVariant.assign(U, 0);
    for (int u = 0; u < U; ++u) {
        hru[0 * e->ldu + u] = area[u] / basinArea;  // ModelBuilder.cpp:468
        hru[1 * e->ldu + u] = area[u];
        // pow(_slope, 0.5) of ProcessRunoffSocont.cpp:51 with _slope float32, evaluated with the host libm
        hru[2 * e->ldu + u] = slope ? std::pow((double)slope[u], 0.5) : 0.0;
        hru[3 * e->ldu + u] = elevation ? elevation[u] : 0.0;
        e->hGlacierVariant[u] = (glacier_variant && glacier_variant[u]) ? 1 : 0;
        flags[u] = e->unitFlags(u);
        if (frac_ground) e->hFracG[u] = frac_ground[u];
        if (frac_glacier) e->hFracGl[u] = frac_glacier[u];
        // a structure without a glacier cover has the ground as its only land cover: the kernels fold its fraction (= 1)
        if (!e->st.has_glacier && e->hFracG[u] != 1.0)
            return fail(e, HBK_E_MODEL, "the ground is the only land cover of this structure: its area fraction must be 1 "
                                        "(the land-cover fractions of a hydro unit sum to 1)");
    }
    std::vector<double> frac(2 * (size_t)U);
    for (int u = 0; u < U; ++u) {
        frac[u] = e->hFracG[u];
        frac[U + u] = e->hFracGl[u];
    }
    HBK_CUDA(cudaMemcpy(e->dHru, hru.data(), sizeof(double) * hru.size(), cudaMemcpyHostToDevice));
    HBK_CUDA(cudaMemcpy(e->dHruFlags, flags.data(), sizeof(int) * flags.size(), cudaMemcpyHostToDevice));
    HBK_CUDA(cudaMemcpy(e->dFracInit, frac.data(), sizeof(double) * frac.size(), cudaMemcpyHostToDevice));
    HBK_CUDA(cudaMemcpy(e->dFrac, frac.data(), sizeof(double) * frac.size(), cudaMemcpyHostToDevice));
    e->unitsSet = true;
    return HBK_OK;
}

int hbk_set_unit_aspect(hbk_engine* e, const int32_t* aspect_class) {
    if (!e || !aspect_class) return HBK_E_ARG;
    HBK_CUDA(cudaSetDevice(e->device));
    for (int u = 0; u < e->U; ++u)
        if (aspect_class[u] < HBK_ASPECT_NORTH || aspect_class[u] > HBK_ASPECT_EAST_WEST)
            return fail(e, HBK_E_ARG, "Invalid aspect: the aspect class of a hydro unit is north, south or east/west "
                                      "(ProcessMeltDegreeDayAspect.cpp:38-58)");
    e->hAspect.assign(aspect_class, aspect_class + e->U);
    std::vector<int> flags(e->ldu, 0);
    for (int u = 0; u < e->U; ++u) flags[u] = e->unitFlags(u);
    HBK_CUDA(cudaMemcpy(e->dHruFlags, flags.data(), sizeof(int) * flags.size(), cudaMemcpyHostToDevice));
    return HBK_OK;
}

int hbk_set_unit_twins(hbk_engine* e, const int32_t* twin_of) {
    if (!e || !twin_of) return HBK_E_ARG;
    if (!e->st.has_glacier) return fail(e, HBK_E_MODEL, "twin units carry a second glacier cover: the structure has none");
    if (HBK_PAR_SHARED) return fail(e, HBK_E_UNSUPPORTED, "this build (HBK_PAR_SHARED) has no twin units");
    bool any = false;
    for (int u = 0; u < e->U; ++u) {
        if (twin_of[u] < -1 || twin_of[u] >= e->U || twin_of[u] == u || (twin_of[u] >= 0 && twin_of[twin_of[u]] >= 0))
            return fail(e, HBK_E_ARG, "twin_of: index of an ordinary unit, or -1");
        any = any || twin_of[u] >= 0;
    }
    e->hTwinOf.assign(twin_of, twin_of + e->U);
    e->hasTwins = any;
    HBK_CUDA(cudaSetDevice(e->device));
    freeDev(e->dTwinOf);
    freeDev(e->dTwinCol);
    if (any) {
        std::vector<int> col(e->U, -1);
        for (int u = 0; u < e->U; ++u) {
            if (twin_of[u] < 0) continue;
            if (col[twin_of[u]] >= 0) return fail(e, HBK_E_UNSUPPORTED, "at most one twin (two glacier covers) per unit");
            col[twin_of[u]] = u;
        }
        HBK_CUDA(cudaMalloc(&e->dTwinOf, sizeof(int) * e->U));
        HBK_CUDA(cudaMalloc(&e->dTwinCol, sizeof(int) * e->U));
        HBK_CUDA(cudaMemcpy(e->dTwinOf, twin_of, sizeof(int) * e->U, cudaMemcpyHostToDevice));
        HBK_CUDA(cudaMemcpy(e->dTwinCol, col.data(), sizeof(int) * e->U, cudaMemcpyHostToDevice));
    }
    e->unitsSet = false;  // hbk_set_units computes the basin area and the unit flags from it
    return HBK_OK;
}

int hbk_set_parameters(hbk_engine* e, int32_t n_sets, const float* values) {
    if (!e || n_sets <= 0 || !values) return HBK_E_ARG;
    HBK_CUDA(cudaSetDevice(e->device));
    for (size_t i = 0; i < (size_t)n_sets * HBK_N_PARAMS; ++i)
        if (!std::isfinite(values[i])) return fail(e, HBK_E_ARG, "parameter values must be finite");
    int rc = ensureSetBuffers(e, n_sets);
    if (rc) return rc;
    // Internal order of the parameter sets. Lanes of a warp are parameter sets; sets whose constraints bind at
    // different times make the warp run both sides of every rare branch (ncu: 23 of 32 lanes active on random
    // sets). Grouping similar sets — buckets of the slow-reservoir response factor (drives the "store would go
    // negative" limit), ordered by capacity inside a bucket (drives the overflow) — gives +14 % on the C2 ensemble.
    // Results do not depend on the order (sets are independent); everything that leaves the engine is mapped back.
    // A per-set initial state saved from a previous run keeps the order it was saved in. HBK_SORT_SETS=0 disables.
    const char* sortEnv = std::getenv("HBK_SORT_SETS");
    const bool wantSort = n_sets >= 64 && !(sortEnv && sortEnv[0] == '0');
    if (!e->initPerSet || (int)e->perm.size() != n_sets) {
        e->perm.resize(n_sets);
        for (int p = 0; p < n_sets; ++p) e->perm[p] = p;
        if (wantSort && !e->initPerSet) {
            // (the two keys gathered into contiguous arrays first: the comparators of a 100 000-set ensemble then stay in
            // cache — this runs inside the end-to-end step)
            std::vector<float> kSlow(n_sets), capacity(n_sets);
            for (int p = 0; p < n_sets; ++p) {
                kSlow[p] = values[(size_t)p * HBK_N_PARAMS + HBK_P_K_SLOW_1];
                capacity[p] = values[(size_t)p * HBK_N_PARAMS + HBK_P_CAPACITY];
            }
            std::vector<int> byK(e->perm);
            std::stable_sort(byK.begin(), byK.end(), [&](int x, int y) { return kSlow[x] < kSlow[y]; });
            const int nWarps = (n_sets + 31) / 32;
            int nBuckets = std::max(1, std::min(64, (int)std::lround(std::sqrt((double)nWarps / 3.0))));
            std::vector<int> bucket(n_sets);
            // tuning knob HBK_SORT_MODE=<second key slot>:<its buckets>[:<k_slow buckets>]: a second bucket level (e.g. the
            // snow degree-day factor or the quick-flow coefficient) between the k_slow buckets and the capacity order
            int key2 = -1, nB2 = 1;
            if (const char* mode = std::getenv("HBK_SORT_MODE")) {
                int b1 = 0;
                if (std::sscanf(mode, "%d:%d:%d", &key2, &nB2, &b1) >= 2 && key2 >= 0 && key2 < HBK_N_PARAMS && nB2 >= 1) {
                    if (b1 > 0) nBuckets = b1;
                } else {
                    key2 = -1;
                    nB2 = 1;
                }
            }
            for (int r = 0; r < n_sets; ++r) bucket[byK[r]] = (int)((long long)r * nBuckets / n_sets) * nB2;
            if (key2 >= 0 && nB2 > 1) {
                std::vector<float> k2(n_sets);
                for (int p = 0; p < n_sets; ++p) k2[p] = values[(size_t)p * HBK_N_PARAMS + key2];
                std::vector<int> by2(e->perm);
                std::stable_sort(by2.begin(), by2.end(), [&](int x, int y) {
                    if (bucket[x] != bucket[y]) return bucket[x] < bucket[y];
                    return k2[x] < k2[y];
                });
                // within every k_slow bucket: rank by the second key -> sub-bucket
                int start = 0;
                while (start < n_sets) {
                    int end = start;
                    while (end < n_sets && bucket[by2[end]] == bucket[by2[start]]) ++end;
                    for (int r = start; r < end; ++r) bucket[by2[r]] += (int)((long long)(r - start) * nB2 / (end - start));
                    start = end;
                }
            }
            std::stable_sort(e->perm.begin(), e->perm.end(), [&](int x, int y) {
                if (bucket[x] != bucket[y]) return bucket[x] < bucket[y];
                return capacity[x] < capacity[y];
            });
        }
        bool identity = true;
        for (int p = 0; p < n_sets && identity; ++p) identity = e->perm[p] == p;
        if (identity) {
            freeDev(e->dPerm);
        } else {
            if (!e->dPerm) HBK_CUDA(cudaMalloc(&e->dPerm, sizeof(int) * n_sets));
            HBK_CUDA(cudaMemcpyAsync(e->dPerm, e->perm.data(), sizeof(int) * n_sets, cudaMemcpyHostToDevice, e->stream));
        }
    }
    // [n_sets][N] -> [N][ldp]: the parameter axis is the fastest one on the device
    // staged in page-locked memory kept by the engine (no 18 MB allocation + pageable copy per end-to-end step)
    const size_t nT = (size_t)HBK_N_PARAMS * e->ldp;
    if (e->paramStageN != nT) {
        if (e->hParamStage) cudaFreeHost(e->hParamStage);
        e->hParamStage = nullptr;
        e->paramStageN = 0;
        HBK_CUDA(cudaHostAlloc((void**)&e->hParamStage, sizeof(float) * nT, cudaHostAllocDefault));
        e->paramStageN = nT;
        std::fill(e->hParamStage, e->hParamStage + nT, 0.0f);
    }
    float* t = e->hParamStage;
    for (int k = 0; k < HBK_N_PARAMS; ++k)
        for (int p = n_sets; p < e->ldp; ++p) t[(size_t)k * e->ldp + p] = 0.0f;  // padding lanes
    for (int p = 0; p < n_sets; ++p) {
        const float* src = values + (size_t)e->perm[p] * HBK_N_PARAMS;
        for (int k = 0; k < HBK_N_PARAMS; ++k) t[(size_t)k * e->ldp + p] = src[k];
    }
    HBK_CUDA(cudaMemcpyAsync(e->dParams, t, sizeof(float) * nT, cudaMemcpyHostToDevice, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    if (e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE) {
        e->hParamsT.assign(t, t + nT);
        e->latThrDirty = true;
    }
    return HBK_OK;
}

int hbk_set_lateral_connections(hbk_engine* e, int32_t n, const int32_t* giver, const int32_t* receiver,
                                const double* fraction, const float* slope_degrees) {
    if (!e || n < 0 || (n > 0 && (!giver || !receiver || !fraction)) || !slope_degrees) return HBK_E_ARG;
    if (e->st.snow_slide_kind == HBK_SNOW_SLIDE_NONE)
        return fail(e, HBK_E_STATE, "lateral connections are only used by structures with transport:snow_slide");
    HBK_CUDA(cudaSetDevice(e->device));
    const int U = e->U;
    std::vector<std::vector<std::pair<int, double>>> per(U);
    for (int i = 0; i < n; ++i) {
        if (giver[i] < 0 || giver[i] >= U || receiver[i] < 0 || receiver[i] >= U)
            return fail(e, HBK_E_ARG, "lateral connection: hydro unit index out of range");
        if (receiver[i] <= giver[i])
            return fail(e, HBK_E_UNSUPPORTED, "lateral connection: the receiving unit must come after the giving unit in "
                                              "basin order (snow slides downhill in a basin sorted by elevation)");
        per[giver[i]].push_back({receiver[i], fraction[i]});
    }
    e->hLatOff.assign(U + 1, 0);
    e->hLatRecv.clear();
    e->hLatWeight.clear();
    for (int u = 0; u < U; ++u) {
        for (auto& c : per[u]) {
            e->hLatRecv.push_back(c.first);
            e->hLatWeight.push_back(c.second);
        }
        e->hLatOff[u + 1] = (int)e->hLatRecv.size();
    }
    e->hSlopeDeg.assign(slope_degrees, slope_degrees + U);
    freeDev(e->dLatOff);
    freeDev(e->dLatRecv);
    freeDev(e->dLatWeight);
    const size_t nc = std::max<size_t>(e->hLatRecv.size(), 1);
    HBK_CUDA(cudaMalloc(&e->dLatOff, sizeof(int) * (U + 1)));
    HBK_CUDA(cudaMalloc(&e->dLatRecv, sizeof(int) * nc));
    HBK_CUDA(cudaMalloc(&e->dLatWeight, sizeof(double) * nc));
    HBK_CUDA(cudaMemcpy(e->dLatOff, e->hLatOff.data(), sizeof(int) * (U + 1), cudaMemcpyHostToDevice));
    if (!e->hLatRecv.empty()) {
        HBK_CUDA(cudaMemcpy(e->dLatRecv, e->hLatRecv.data(), sizeof(int) * e->hLatRecv.size(), cudaMemcpyHostToDevice));
        HBK_CUDA(cudaMemcpy(e->dLatWeight, e->hLatWeight.data(), sizeof(double) * e->hLatWeight.size(),
                            cudaMemcpyHostToDevice));
    }
    e->latThrDirty = true;
    return HBK_OK;
}

// Snow holding threshold [mm of snow depth] of every (unit, parameter set), ProcessLateralSnowSlide.cpp:62-78, with the
// host libm: `*_coeff * pow(slope, *_exp)` resolves to the C pow(double, double) on the promoted floats.
static int uploadSlideThresholds(hbk_engine* e) {
    const int U = e->U, ldp = e->ldp;
    std::vector<double> thr((size_t)U * ldp, 0.0);
    auto par = [&](int slot, int p) { return e->hParamsT[(size_t)slot * ldp + p]; };
    for (int u = 0; u < U; ++u) {
        const float slopeDeg = e->hSlopeDeg[u];
        for (int p = 0; p < e->P; ++p) {
            const float coeff = par(HBK_P_SLIDE_COEFF, p), expo = par(HBK_P_SLIDE_EXP, p);
            const float minSlope = par(HBK_P_SLIDE_MIN_SLOPE, p), maxSlope = par(HBK_P_SLIDE_MAX_SLOPE, p);
            const float minHolding = par(HBK_P_SLIDE_MIN_HOLDING_DEPTH, p), maxDepth = par(HBK_P_SLIDE_MAX_SNOW_DEPTH, p);
            const float slope = std::max(slopeDeg, minSlope);
            const double meters = coeff * ::pow((double)slope, (double)expo);
            double threshold = meters * 1000.0;
            if (slopeDeg > maxSlope) threshold = minHolding;
            threshold = std::max<double>(threshold, minHolding);
            if (maxDepth > 0.0f) threshold = std::min<double>(threshold, maxDepth);
            thr[(size_t)u * ldp + p] = threshold;
        }
    }
    if (!e->dLatThr) HBK_CUDA(cudaMalloc(&e->dLatThr, sizeof(double) * thr.size()));
    HBK_CUDA(cudaMemcpy(e->dLatThr, thr.data(), sizeof(double) * thr.size(), cudaMemcpyHostToDevice));
    e->latThrDirty = false;
    return HBK_OK;
}

int hbk_set_forcing(hbk_engine* e, int32_t var, const double* data) {
    if (!e || var < 0 || var >= HBK_N_VARS || !data) return HBK_E_ARG;
    HBK_CUDA(cudaSetDevice(e->device));
    const size_t n = (size_t)e->T * e->U;
    if (!e->dForcingRaw[var]) HBK_CUDA(cudaMalloc(&e->dForcingRaw[var], sizeof(double) * n));
    HBK_CUDA(cudaMemcpyAsync(e->dForcingRaw[var], data, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    e->haveRaw[var] = true;
    e->haveStation[var] = false;
    e->tiledUB = 0;  // re-tile on the next run
    return HBK_OK;
}

int hbk_set_station_forcing(hbk_engine* e, int32_t var, const double* series, double ref_elevation, double gradient,
                            const double* gradient_per_set, double correction, const double* correction_per_set) {
    if (!e || var < 0 || var >= HBK_N_VARS || !series) return HBK_E_ARG;
    HBK_CUDA(cudaSetDevice(e->device));
    if ((gradient_per_set || correction_per_set) && e->P == 0)
        return fail(e, HBK_E_STATE, "set the parameters before per-set spatialisation gradients");
    if (!e->dStation[var]) {
        HBK_CUDA(cudaMalloc(&e->dStation[var], sizeof(double) * e->tld));
    }
    HBK_CUDA(cudaMemsetAsync(e->dStation[var], 0, sizeof(double) * e->tld, e->stream));
    HBK_CUDA(cudaMemcpyAsync(e->dStation[var], series, sizeof(double) * e->T, cudaMemcpyHostToDevice, e->stream));
    auto perSet = [&](double*& dst, const double* src) -> int {
        if (!src) {
            freeDev(dst);
            return HBK_OK;
        }
        if (!dst) HBK_CUDA(cudaMalloc(&dst, sizeof(double) * e->P));
        HBK_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * e->P, cudaMemcpyHostToDevice, e->stream));
        return HBK_OK;
    };
    int rc = HBK_OK;
    if (var == HBK_VAR_TEMPERATURE) {
        e->gradTs = gradient;
        e->zrefT = ref_elevation;
        rc = perSet(e->dGradT, gradient_per_set);
    } else if (var == HBK_VAR_PRECIPITATION) {
        e->gradPs = gradient;
        e->zrefP = ref_elevation;
        e->corrPs = correction;
        rc = perSet(e->dGradP, gradient_per_set);
        if (!rc) rc = perSet(e->dCorrP, correction_per_set);
    }
    if (rc) return rc;
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    if (var == HBK_VAR_TEMPERATURE || var == HBK_VAR_PRECIPITATION) {
        // the station call of a variable replaces its per-set arrays: once both spatialised variables have been set
        // again after a change of n_sets nothing stale or missing is left
        e->perSetDroppedVars &= ~(1 << var);
        if (e->perSetDroppedVars == 0) e->perSetDropped = false;
    }
    e->haveStation[var] = true;
    e->haveRaw[var] = false;
    return HBK_OK;
}

int hbk_set_initial_state(hbk_engine* e, int32_t slot, const double* values) {
    if (!e || slot < 0 || slot >= HBK_N_STATES) return HBK_E_ARG;
    if (slot >= e->nStates()) return fail(e, HBK_E_ARG, "this state slot exists only with liquid water in the snowpacks");
    HBK_CUDA(cudaSetDevice(e->device));
    if (values) {
        HBK_CUDA(cudaMemcpyAsync(e->dInitUnit + (size_t)slot * e->U, values, sizeof(double) * e->U, cudaMemcpyHostToDevice,
                                 e->stream));
    } else {
        HBK_CUDA(cudaMemsetAsync(e->dInitUnit + (size_t)slot * e->U, 0, sizeof(double) * e->U, e->stream));
    }
    if (e->initPerSet && e->dBasinInit) {
        // leaving the per-set initial state of hbk_save_as_initial_state: the sub-basin stores start empty again
        // (their saved contents belong to the saved unit states, not to the per-unit initial contents used from now on)
        HBK_CUDA(cudaMemsetAsync(e->dBasinInit, 0, sizeof(double) * 2 * e->P, e->stream));
    }
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    e->initPerSet = false;
    return HBK_OK;
}

int hbk_save_as_initial_state(hbk_engine* e) {
    if (!e || !e->ran) return fail(e, HBK_E_STATE, "save_as_initial_state needs a completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    const size_t n = sizeof(double) * e->nStates() * (size_t)e->P * e->U;
    if (!e->dInitFull) HBK_CUDA(cudaMalloc(&e->dInitFull, n));
    // on the engine's stream (created non-blocking: the legacy default stream is NOT ordered against it), then drained,
    // so that the next hbk_run's restore copy can never overlap this save
    HBK_CUDA(cudaMemcpyAsync(e->dInitFull, e->dState, n, cudaMemcpyDeviceToDevice, e->stream));
    HBK_CUDA(cudaMemcpyAsync(e->dBasinInit, e->dBasinState, sizeof(double) * 2 * e->P, cudaMemcpyDeviceToDevice, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    e->initPerSet = true;
    return HBK_OK;
}

int hbk_set_spinup_steps(hbk_engine* e, int32_t n) {
    if (!e || n < 0) return HBK_E_ARG;
    e->spinupSteps = std::min<int>(n, e->T);  // ModelHydro.cpp:52-53
    return HBK_OK;
}

int hbk_set_record_mask(hbk_engine* e, uint64_t mask) {
    if (!e) return HBK_E_ARG;
    e->recordMask = mask & ((1ull << HBK_N_RECORDS) - 1);
    return HBK_OK;
}

int hbk_clear_actions(hbk_engine* e) {
    if (!e) return HBK_E_ARG;
    e->events.clear();
    e->dhUnits = 0;
    e->hDhUnits.clear();
    e->dhTablesSet = false;
    return HBK_OK;
}

int hbk_add_land_cover_change(hbk_engine* e, int32_t step, int32_t unit, double fraction) {
    // unit = -1 on a unit shard: the edit concerns a unit of another block; this block only cuts its run at the same step
    // (all shards must run the same segments: their per-segment sums are reduced together)
    if (!e || unit < -1 || (unit == -1 && !e->shard) || unit >= e->U || step < 1) return HBK_E_ARG;
    if (!e->st.has_glacier) return fail(e, HBK_E_MODEL, "land-cover change needs a structure with a glacier cover");
    if (!e->unitsSet || (unit >= 0 && !e->hGlacierVariant[unit]))
        return fail(e, HBK_E_MODEL, "land-cover change on a hydro unit whose structure variant has no glacier cover");
    if (fraction < 0 || fraction > 1) return fail(e, HBK_E_MODEL, "The fraction is not in the range [0 .. 1]");
    e->events.push_back({step, 0, unit, fraction});
    return HBK_OK;
}

int hbk_add_snow_to_ice_on(hbk_engine* e, int32_t step, int32_t glacier_cover) {
    if (!e || step < 1) return HBK_E_ARG;
    if (!e->st.has_glacier) return fail(e, HBK_E_MODEL, "snow-to-ice transformation needs a glacier cover");
    if (glacier_cover < 0 || glacier_cover > 1 || (glacier_cover == 1 && !e->hasTwins))
        return fail(e, HBK_E_ARG, "snow to ice: glacier cover 0, or 1 with twin units (hbk_set_unit_twins)");
    e->events.push_back({step, 1, glacier_cover, 0.0});  // Event::unit carries the glacier cover index
    return HBK_OK;
}

int hbk_add_snow_to_ice(hbk_engine* e, int32_t step) { return hbk_add_snow_to_ice_on(e, step, 0); }

int hbk_add_delta_h_update(hbk_engine* e, int32_t step) {
    if (!e || step < 1) return HBK_E_ARG;
    if (!e->dhTablesSet) return fail(e, HBK_E_STATE, "hbk_set_delta_h_tables has not been called");
    e->events.push_back({step, 2, -1, 0.0});
    return HBK_OK;
}

int hbk_add_area_scaling_update(hbk_engine* e, int32_t step) {
    if (!e || step < 1) return HBK_E_ARG;
    if (!e->dhTablesSet) return fail(e, HBK_E_STATE, "hbk_set_delta_h_tables has not been called");
    e->events.push_back({step, 3, -1, 0.0});
    return HBK_OK;
}

int hbk_set_delta_h_tables(hbk_engine* e, int32_t n, const int32_t* units, int32_t nRows, const double* areas,
                           const double* volumes) {
    // n = 0: a unit shard that holds none of the table's units (it still takes part in the reductions of every update)
    if (!e || n < 0 || (n == 0 && !e->shard) || nRows < 2 || (n > 0 && (!units || !areas || !volumes))) return HBK_E_ARG;
    if (!e->unitsSet) return fail(e, HBK_E_STATE, "hbk_set_units must precede hbk_set_delta_h_tables");
    if (!e->st.has_glacier || e->st.glacier_infinite_storage)
        return fail(e, HBK_E_MODEL, "glacier evolution needs a glacier cover with a finite ice storage");
    HBK_CUDA(cudaSetDevice(e->device));
    // ActionGlacierEvolutionDeltaH::Init (ActionGlacierEvolutionDeltaH.cpp:21-70)
    std::vector<double> ice(e->U, 0.0);
    HBK_CUDA(cudaMemcpy(ice.data(), e->dInitUnit + (size_t)HBK_S_ICE * e->U, sizeof(double) * e->U, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const int u = units[i];
        if (u < 0 || u >= e->U) return fail(e, HBK_E_ARG, "delta-h: hydro unit index out of range");
        if (!e->hGlacierVariant[u]) return fail(e, HBK_E_MODEL, "delta-h: the land cover was not found in a hydro unit");
        const double areaRef = areas[i];
        const double areaInModel = e->hFracGl[u] * e->hArea[u];
        if (std::fabs(areaInModel) <= 1e-8) {
            double fraction = areaRef / e->hArea[u];
            if (std::fabs(fraction) <= 1e-8) fraction = 0.0;
            if (std::fabs(1 - fraction) <= 1e-8) fraction = 1.0;
            if (fraction < 0 || fraction > 1) return fail(e, HBK_E_MODEL, "delta-h: initial fraction not in [0 .. 1]");
            // the generic cover that gives the area: the unit's own (a twin carries only the second glacier cover)
            const int uGen = (e->hasTwins && e->hTwinOf[u] >= 0) ? e->hTwinOf[u] : u;
            e->hFracG[uGen] -= fraction - e->hFracGl[u];
            e->hFracGl[u] = fraction;
        } else if (std::fabs(areaInModel - areaRef) > 1e-8) {
            return fail(e, HBK_E_MODEL, "The glacier area fraction of a hydro unit does not match the lookup table "
                                        "initial area.");
        }
        ice[u] = volumes[i] * 900.0 / areaRef;  // constants::iceDensity (GlobVars.h:8)
    }
    std::vector<double> frac(2 * (size_t)e->U);
    for (int u = 0; u < e->U; ++u) {
        frac[u] = e->hFracG[u];
        frac[e->U + u] = e->hFracGl[u];
    }
    HBK_CUDA(cudaMemcpy(e->dFracInit, frac.data(), sizeof(double) * frac.size(), cudaMemcpyHostToDevice));
    HBK_CUDA(cudaMemcpy(e->dInitUnit + (size_t)HBK_S_ICE * e->U, ice.data(), sizeof(double) * e->U, cudaMemcpyHostToDevice));
    freeDev(e->dDhUnits);
    freeDev(e->dDhAreas);
    freeDev(e->dDhVolumes);
    HBK_CUDA(cudaMalloc(&e->dDhUnits, sizeof(int) * std::max(n, 1)));
    HBK_CUDA(cudaMalloc(&e->dDhAreas, sizeof(double) * std::max<size_t>((size_t)n * nRows, 1)));
    HBK_CUDA(cudaMalloc(&e->dDhVolumes, sizeof(double) * std::max<size_t>((size_t)n * nRows, 1)));
    if (n > 0) {
        HBK_CUDA(cudaMemcpy(e->dDhUnits, units, sizeof(int) * n, cudaMemcpyHostToDevice));
        HBK_CUDA(cudaMemcpy(e->dDhAreas, areas, sizeof(double) * (size_t)n * nRows, cudaMemcpyHostToDevice));
        HBK_CUDA(cudaMemcpy(e->dDhVolumes, volumes, sizeof(double) * (size_t)n * nRows, cudaMemcpyHostToDevice));
    }
    e->dhUnits = n;
    e->hDhUnits.assign(units, units + n);
    e->dhRows = nRows;
    e->dhTablesSet = true;
    double initialVolume = 0;
    for (int i = 0; i < n; ++i) initialVolume += volumes[i];
    e->dhInitialWE = initialVolume * 900.0;
    e->dhShardTotals = false;
    return HBK_OK;
}

int hbk_set_delta_h_shard_totals(hbk_engine* e, int32_t n_rows, const double* row_volume_sums) {
    if (!e || !row_volume_sums) return HBK_E_ARG;
    if (e->dhRows == 0) return fail(e, HBK_E_STATE, "hbk_set_delta_h_tables has not been called");
    if (n_rows != e->dhRows) return fail(e, HBK_E_ARG, "delta-h: n_rows differs from the lookup tables");
    HBK_CUDA(cudaSetDevice(e->device));
    freeDev(e->dDhRowSums);
    HBK_CUDA(cudaMalloc(&e->dDhRowSums, sizeof(double) * n_rows));
    HBK_CUDA(cudaMemcpy(e->dDhRowSums, row_volume_sums, sizeof(double) * n_rows, cudaMemcpyHostToDevice));
    e->dhInitialWE = row_volume_sums[0] * 900.0;  // ActionGlacierEvolutionDeltaH.cpp:66-69 over the whole table
    e->dhShardTotals = true;
    return HBK_OK;
}

// Sum over all shards, through the caller's callback, of column blocks [t0, t1) of row-major [P][ldt] buffers and of
// contiguous vectors: gathered into one contiguous staging buffer (one collective), then scattered back.
struct ShardBlock {
    double* buf;
    long long ld;  // row stride in doubles; 0 = contiguous vector of n doubles
    int t0, t1;
    size_t n;
};
static int shardReduceBlocks(hbk_engine* e, const std::vector<ShardBlock>& blocks) {
    size_t total = 0;
    for (const auto& b : blocks) total += b.ld ? (size_t)e->P * (b.t1 - b.t0) : b.n;
    if (total == 0) return HBK_OK;
    if (total > e->shardStageN) {
        freeDev(e->dShardStage);
        HBK_CUDA(cudaMalloc(&e->dShardStage, sizeof(double) * total));
        e->shardStageN = total;
    }
    size_t off = 0;
    for (const auto& b : blocks) {
        if (b.ld) {
            const size_t w = (size_t)(b.t1 - b.t0);
            HBK_CUDA(cudaMemcpy2DAsync(e->dShardStage + off, w * 8, b.buf + b.t0, (size_t)b.ld * 8, w * 8, e->P,
                                       cudaMemcpyDeviceToDevice, e->stream));
            off += (size_t)e->P * w;
        } else {
            HBK_CUDA(cudaMemcpyAsync(e->dShardStage + off, b.buf, b.n * 8, cudaMemcpyDeviceToDevice, e->stream));
            off += b.n;
        }
    }
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    if (e->shardReduce(e->shardUser, e->dShardStage, (int64_t)total) != 0)
        return fail(e, HBK_E_STATE, "the shard reduction callback (hbk_set_shard_reduce) reported a failure");
    off = 0;
    for (const auto& b : blocks) {
        if (b.ld) {
            const size_t w = (size_t)(b.t1 - b.t0);
            HBK_CUDA(cudaMemcpy2DAsync(b.buf + b.t0, (size_t)b.ld * 8, e->dShardStage + off, w * 8, w * 8, e->P,
                                       cudaMemcpyDeviceToDevice, e->stream));
            off += (size_t)e->P * w;
        } else {
            HBK_CUDA(cudaMemcpyAsync(b.buf, e->dShardStage + off, b.n * 8, cudaMemcpyDeviceToDevice, e->stream));
            off += b.n;
        }
    }
    return HBK_OK;
}

static int applyEvents(hbk_engine* e, int step) {
    ActArgs a{};
    a.P = e->P;
    a.U = e->U;
    a.frac = e->dFracSet;
    a.state = e->dState;
    a.lds = (long long)e->P * e->U;
    a.sP = e->stateLaneP ? 1 : e->U;
    a.sU = e->stateLaneP ? e->P : 1;
    a.hru = e->dHru;
    a.hruFlags = e->dHruFlags;
    a.ldu = e->ldu;
    a.iceInfinite = e->st.glacier_infinite_storage;
    a.retention = e->st.snow_retention;
    a.err = e->dErr;
    a.twinOf = e->dTwinOf;
    a.twinCol = e->dTwinCol;
    const int pBlocks = (e->P + 127) / 128;
    for (const auto& ev : e->events) {
        if (ev.step != step) continue;
        if (ev.kind == 0) {
            if (ev.unit < 0) continue;  // an edit in another block of the basin (unit shard): boundary only
            hbkActLandCoverChange<<<pBlocks, 128, 0, e->stream>>>(a, ev.unit, ev.value);
        } else if (ev.kind == 1) {
            hbkActSnowToIce<<<148 * 2, 256, 0, e->stream>>>(a, ev.unit);  // ev.unit: glacier cover index
        } else if (ev.kind == 3) {
            if (e->dhUnits == 0) continue;  // a unit shard without any unit of the table: nothing to scale here
            const long long n = (long long)e->P * e->dhUnits;
            hbkActAreaScaling<<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(a, e->dhUnits, e->dDhUnits, e->dhRows,
                                                                                  e->dDhAreas, e->dDhVolumes);
        } else if (e->shard) {
            // the glacier of the whole basin: this block's water equivalent, summed over the shards by the caller
            if (!e->dShardWE) HBK_CUDA(cudaMalloc(&e->dShardWE, sizeof(double) * e->P));
            hbkActDeltaHSum<<<pBlocks, 128, 0, e->stream>>>(a, e->dhUnits, e->dDhUnits, e->dShardWE);
            int rc = shardReduceBlocks(e, {{e->dShardWE, 0, 0, 0, (size_t)e->P}});
            if (rc) return rc;
            hbkActDeltaHApply<<<pBlocks, 128, 0, e->stream>>>(a, e->dhUnits, e->dDhUnits, e->dhRows, e->dDhAreas,
                                                              e->dDhVolumes, e->dDhRowSums, e->dhInitialWE, e->dDhLastRow,
                                                              e->dShardWE);
            e->lastLaunches++;
        } else {
            hbkActDeltaH<<<pBlocks, 128, 0, e->stream>>>(a, e->dhUnits, e->dDhUnits, e->dhRows, e->dDhAreas, e->dDhVolumes,
                                                         e->dhInitialWE, e->dDhLastRow);
        }
        e->lastLaunches++;
    }
    HBK_CUDA(cudaGetLastError());
    return HBK_OK;
}

// The two sub-basin stores over [tBegin, tEnd) with the structure's solver (hbkBasinStores).
static void launchBasinStores(hbk_engine* e, int tBegin, int tEnd, const double* stale, int writeOutputs) {
    const int blocks = (e->P + 127) / 128;
    const double dt = e->st.time_step_days;
    const int seqKind = e->st.solver >= HBK_SOLVER_CRANK_NICOLSON ? e->st.solver - HBK_SOLVER_CRANK_NICOLSON : 0;
#define HBK_BASIN(S)                                                                                                   \
    hbkBasinStores<S><<<blocks, 128, 0, e->stream>>>(e->dParams, e->ldp, e->P, e->ldt, tBegin, tEnd, dt, e->dD1, e->dD2, \
                                                     stale, e->dOutlet, e->dBasinState, writeOutputs, e->dErr,         \
                                                     e->dUnconv, e->dPerm, seqKind)
    switch (e->st.solver) {
        case HBK_SOLVER_EULER: HBK_BASIN(0); break;
        case HBK_SOLVER_HEUN: HBK_BASIN(1); break;
        case HBK_SOLVER_RK4: HBK_BASIN(2); break;
        default: HBK_BASIN(3); break;
    }
#undef HBK_BASIN
}

// The tiled unit kernel (hbkRunUnits) over one segment.
static int launchUnitKernel(hbk_engine* e, KArgs& a, bool tiled, int tBegin, int tEnd, const hbk_engine::TileRun& run,
                            cudaStream_t stream, int slot) {
    a.TC = run.TC;
    a.G = run.G;
    a.nGroups = run.nGroups;
    a.tile0 = run.tile0;
    a.tile1 = run.tile1;
    a.grpBase = run.grpBase;
    const bool fixedShape = run.TC == hbk::hbkEnsTC(run.glacier) && !tiled && !e->clusterMode && a.nVars == 3 &&
                            !e->st.snow_retention && !std::getenv("HBK_NO_ENS_SHAPE");
    const int shape = !fixedShape ? 0 : ((e->PB == 32 && e->UB == 4) ? 1 : ((e->PB == 1 && e->UB == 128) ? 2 : 0));
    KernelFn fn = pickKernel(e->st, run.glacier, shape);
    const size_t smem = smemBytes(e, tiled, run.glacier, run.TC);
    HBK_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nPGroups = (e->P + e->PB - 1) / e->PB;
    const long long grid = (long long)nPGroups * (e->clusterMode ? e->nUTiles : run.nGroups);
    if (grid > 2147483647LL) return fail(e, HBK_E_UNSUPPORTED, "grid too large");
    // tuning aid: HBK_BLOCK_TIMES=<file> dumps start / end time and SM of every block of the (last) unit launch
    unsigned long long* dTimes = nullptr;
    const char* timesPath = std::getenv("HBK_BLOCK_TIMES");
    if (timesPath && cudaMalloc(&dTimes, sizeof(unsigned long long) * 3 * (size_t)grid) == cudaSuccess &&
        cudaMemset(dTimes, 0, sizeof(unsigned long long) * 3 * (size_t)grid) == cudaSuccess)
        a.blockTimes = dTimes;
    // Work queue: when there are more blocks than the device holds at once, launch one persistent block per slot and let
    // the blocks pull (time slice, virtual block) items — otherwise the launch ends with a long tail in which the slower
    // parameter-set groups finish alone (measured on C2: the last quarter of the launch ran at falling residency).
    unsigned launchGrid = (unsigned)grid;
    a.workCounter = nullptr;
    a.progress = nullptr;
    {
        int perSM = 0;
        HBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, (const void*)fn, e->PB * e->UB, smem));
        int nSM = 148;
        cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, e->device);
        long long slots = (long long)perSM * nSM;
        if (const char* forced = std::getenv("HBK_QUEUE_SLOTS")) {  // testing knob: pretend the device holds this many blocks
            const long long v = std::atoll(forced);
            if (v > 0) slots = std::min(slots, v);
        }
        const int nChunks = (tEnd - tBegin + run.TC - 1) / run.TC;
        const char* q = std::getenv("HBK_QUEUE");
        const bool want = !(q && q[0] == '0') && !e->clusterMode && slots > 0 && grid > slots && nChunks >= 8;
        if (want) {
            const int targetSlices = 16;
            a.sliceChunks = std::max(4, (nChunks + targetSlices - 1) / targetSlices);
            const int nSlices = (nChunks + a.sliceChunks - 1) / a.sliceChunks;
            a.nVBlocks = (int)grid;
            const long long nItems = grid * nSlices;
            if (nItems < 2000000000LL) {
                a.nItems = (int)nItems;
                const size_t need = 1 + (size_t)grid;
                if (need > e->queueInts[slot]) {
                    freeDev(e->dQueue[slot]);
                    HBK_CUDA(cudaMalloc(&e->dQueue[slot], sizeof(int) * need));
                    e->queueInts[slot] = need;
                }
                HBK_CUDA(cudaMemsetAsync(e->dQueue[slot], 0, sizeof(int) * need, stream));
                a.workCounter = e->dQueue[slot];
                a.progress = e->dQueue[slot] + 1;
                launchGrid = (unsigned)slots;
            }
        }
    }
    void* args[] = {(void*)&a};
    if (e->clusterMode) {
        HBK_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(e->PB * e->UB);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)e->nUTiles;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        HBK_CUDA(cudaLaunchKernelExC(&cfg, (const void*)fn, args));
    } else {
        HBK_CUDA(cudaLaunchKernel((const void*)fn, dim3(launchGrid), dim3(e->PB * e->UB), args, smem, stream));
    }
    e->lastLaunches++;
    if (dTimes) {
        std::vector<unsigned long long> h(3 * (size_t)grid);
        if (cudaStreamSynchronize(stream) == cudaSuccess &&
            cudaMemcpy(h.data(), dTimes, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (FILE* f = std::fopen(timesPath, "wb")) {
                std::fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
                std::fclose(f);
            }
        }
        cudaFree(dTimes);
    }
    return HBK_OK;
}

// transport:snow_slide: the sequential-over-units kernel (hbkRunLateral), one thread per parameter set.
static int launchLateralKernel(hbk_engine* e, KArgs& a) {
    LatArgs L{};
    L.off = e->dLatOff;
    L.recv = e->dLatRecv;
    L.weight = e->dLatWeight;
    L.thr = e->dLatThr;
    L.pend = e->dLatPend;
    L.skipGlaciers = e->st.snow_slide_kind == HBK_SNOW_SLIDE_SKIP_GLACIERS ? 1 : 0;
    for (int v = 0; v < a.nVars; ++v) L.raw[v] = e->dForcingRaw[v];
    const int blocks = (e->P + 127) / 128;
    const bool glacier = e->st.has_glacier != 0;
    switch (e->st.solver) {
        case HBK_SOLVER_EULER: lateralSoil<0>(e->st.n_soil_stores, glacier, blocks, e->stream, a, L); break;
        case HBK_SOLVER_HEUN: lateralSoil<1>(e->st.n_soil_stores, glacier, blocks, e->stream, a, L); break;
        case HBK_SOLVER_RK4: lateralSoil<2>(e->st.n_soil_stores, glacier, blocks, e->stream, a, L); break;
        default: lateralSoil<3>(e->st.n_soil_stores, glacier, blocks, e->stream, a, L); break;
    }
    e->lastLaunches++;
    HBK_CUDA(cudaGetLastError());
    return HBK_OK;
}

static int launchSegment(hbk_engine* e, int tBegin, int tEnd, bool writeOutputs) {
    const bool tiled = e->haveRaw[0];
    KArgs a{};
    a.P = e->P;
    a.U = e->U;
    a.tBegin = tBegin;
    a.tEnd = tEnd;
    a.PB = e->PB;
    a.UB = e->UB;
    a.TC = e->TC;
    a.nUTiles = e->nUTiles;
    a.G = e->G;
    a.nGroups = e->nGroups;
    a.clusterMode = e->clusterMode ? 1 : 0;
    a.flags.quickKind = e->st.quick_kind;
    a.flags.slowOrder = e->st.slow_process_order;
    a.flags.splitKind = e->st.splitter_kind;
    a.flags.iceInfinite = e->st.glacier_infinite_storage;
    a.flags.snowIceKind = e->st.has_glacier ? e->st.snow_ice_kind : 0;
    a.flags.sublimKind = e->st.sublimation_kind;
    a.flags.seqKind = e->st.solver >= HBK_SOLVER_CRANK_NICOLSON ? e->st.solver - HBK_SOLVER_CRANK_NICOLSON : 0;
    a.flags.meltKind = e->st.melt_kind;
    a.flags.hasTwins = e->hasTwins ? 1 : 0;
    a.flags.retention = e->st.snow_retention;
    a.nVars = e->nVars();
    a.swatFactor = (a.flags.snowIceKind == HBK_SNOW_ICE_SWAT) ? e->dSwatFactor : nullptr;
    a.flags.noMeltWhenSnow = e->st.glacier_no_melt_when_snow_cover;
    a.dt = e->st.time_step_days;
    a.params = e->dParams;
    a.ldp = e->ldp;
    a.hru = e->dHru;
    a.hruFlags = e->dHruFlags;
    a.ldu = e->ldu;
    a.frac = e->fracPerSet ? e->dFracSet : e->dFrac;
    a.fracPerSet = e->fracPerSet;
    a.fracLd = e->fracPerSet ? (long long)e->P * e->U : e->U;
    a.forcingMode = tiled ? 0 : 1;
    for (int v = 0; v < a.nVars; ++v) a.forcing[v] = tiled ? e->dForcingTiled[v] : e->dStation[v];
    a.tld = e->tld;
    a.gradT = e->dGradT;
    a.gradP = e->dGradP;
    a.corrP = e->dCorrP;
    a.gradTs = e->gradTs;
    a.gradPs = e->gradPs;
    a.corrPs = e->corrPs;
    a.zrefT = e->zrefT;
    a.zrefP = e->zrefP;
    a.state = e->dState;
    a.lds = (long long)e->P * e->U;
    a.sP = e->stateLaneP ? 1 : e->U;
    a.sU = e->stateLaneP ? e->P : 1;
    const bool multi = e->totalGroups > 1;
    a.outQ = multi ? e->dPartQ : e->dOutlet;
    a.outD1 = multi ? e->dPartD1 : e->dD1;
    a.outD2 = multi ? e->dPartD2 : e->dD2;
    a.ldt = e->ldt;
    a.writeOutputs = writeOutputs ? 1 : 0;
    a.recordMask = e->recordMask;
    a.rec = e->dRec;
    a.T = e->T;
    a.err = e->dErr;
    a.unconverged = e->dUnconv;
    a.unconvergedPerSet = e->dUnconvSet;
    a.perm = e->dPerm;
    a.outQFinal = multi ? 0 : 1;

    if (e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE) {
        int rc = launchLateralKernel(e, a);
        if (rc) return rc;
    } else {
        // the launches of a split glacier structure are independent (their own units, partial buffers and work queues):
        // they run side by side, the second and later ones on auxiliary streams forked from / joined to the engine's
        if (e->st.has_glacier && e->glacierGroups == 0) {
            // no unit with ice: nothing deposits into the sub-basin stores in this segment (the series may hold the sums
            // of an earlier reduction over the shards, e.g. of the spin-up over the same steps)
            const size_t width = sizeof(double) * (size_t)(tEnd - tBegin), pitch = sizeof(double) * (size_t)e->ldt;
            HBK_CUDA(cudaMemset2DAsync(e->dD1 + tBegin, pitch, 0, width, e->P, e->stream));
            HBK_CUDA(cudaMemset2DAsync(e->dD2 + tBegin, pitch, 0, width, e->P, e->stream));
        }
        const char* serial = std::getenv("HBK_SPLIT_SERIAL");
        const bool fork = e->runs.size() > 1 && !(serial && serial[0] == '1');
        if (fork) {
            if (!e->evFork) HBK_CUDA(cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming));
            HBK_CUDA(cudaEventRecord(e->evFork, e->stream));
        }
        for (size_t i = 0; i < e->runs.size(); ++i) {
            cudaStream_t stream = e->stream;
            if (fork && i > 0) {
                if (!e->auxStream[i]) HBK_CUDA(cudaStreamCreateWithFlags(&e->auxStream[i], cudaStreamNonBlocking));
                if (!e->evJoin[i]) HBK_CUDA(cudaEventCreateWithFlags(&e->evJoin[i], cudaEventDisableTiming));
                stream = e->auxStream[i];
                HBK_CUDA(cudaStreamWaitEvent(stream, e->evFork, 0));
            }
            int rc = launchUnitKernel(e, a, tiled, tBegin, tEnd, e->runs[i], stream, (int)i);
            if (rc) return rc;
            if (fork && i > 0) {
                HBK_CUDA(cudaEventRecord(e->evJoin[i], stream));
                HBK_CUDA(cudaStreamWaitEvent(e->stream, e->evJoin[i], 0));
            }
        }
    }

    if (writeOutputs && multi) {
        const long long n = (long long)e->P * (tEnd - tBegin);
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
        hbkReduceTiles<<<blocks, 256, 0, e->stream>>>(e->dPartQ, e->dOutlet, e->totalGroups, e->P, e->ldt, tBegin, tEnd, e->dPerm);
        e->lastLaunches++;
        if (e->st.has_glacier && e->glacierGroups > 0) {
            hbkReduceTiles<<<blocks, 256, 0, e->stream>>>(e->dPartD1, e->dD1, e->glacierGroups, e->P, e->ldt, tBegin, tEnd, nullptr);
            hbkReduceTiles<<<blocks, 256, 0, e->stream>>>(e->dPartD2, e->dD2, e->glacierGroups, e->P, e->ldt, tBegin, tEnd, nullptr);
            e->lastLaunches += 2;
        }
    } else if (!writeOutputs && multi && e->st.has_glacier && e->glacierGroups > 0) {
        // spin-up: the deposits are still needed by the sub-basin stores
        const long long n = (long long)e->P * (tEnd - tBegin);
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
        hbkReduceTiles<<<blocks, 256, 0, e->stream>>>(e->dPartD1, e->dD1, e->glacierGroups, e->P, e->ldt, tBegin, tEnd, nullptr);
        hbkReduceTiles<<<blocks, 256, 0, e->stream>>>(e->dPartD2, e->dD2, e->glacierGroups, e->P, e->ldt, tBegin, tEnd, nullptr);
        e->lastLaunches += 2;
    }
    if (e->st.has_glacier && e->shard && e->shardReduce && writeOutputs) {
        // the sub-basin stores wait for the sums over the shards (end of hbk_run); what has to be kept per segment is the
        // stale deposit amounts of this block's vanished glaciers
        const size_t seg = e->shardSegments.size();
        e->shardSegments.push_back({tBegin, tEnd});
        if (e->fracPerSet) {
            const size_t need = (seg + 1) * 2 * (size_t)e->P;
            if (need > e->staleSegN) {
                const size_t cap = std::max(need, 2 * e->staleSegN + 64 * (size_t)e->P);
                double* grown = nullptr;
                HBK_CUDA(cudaMalloc(&grown, sizeof(double) * cap));
                if (e->dStaleSeg)
                    HBK_CUDA(cudaMemcpyAsync(grown, e->dStaleSeg, sizeof(double) * e->staleSegN, cudaMemcpyDeviceToDevice,
                                             e->stream));
                HBK_CUDA(cudaStreamSynchronize(e->stream));
                freeDev(e->dStaleSeg);
                e->dStaleSeg = grown;
                e->staleSegN = cap;
            }
            ActArgs aa{};
            aa.P = e->P;
            aa.U = e->U;
            aa.frac = e->dFracSet;
            aa.state = e->dState;
            aa.lds = (long long)e->P * e->U;
            aa.sP = e->stateLaneP ? 1 : e->U;
            aa.sU = e->stateLaneP ? e->P : 1;
            aa.hruFlags = e->dHruFlags;
            aa.ldu = e->ldu;
            hbkStaleDeposits<<<e->P, 128, 0, e->stream>>>(aa, e->dStaleSeg + seg * 2 * (size_t)e->P);
            e->lastLaunches++;
        }
    }
    if (e->st.has_glacier && !e->shard) {
        const double* stale = nullptr;
        if (e->fracPerSet) {  // a glacier may have vanished through an action: its stale flux amounts
            if (!e->dStale) HBK_CUDA(cudaMalloc(&e->dStale, sizeof(double) * 2 * e->P));
            ActArgs aa{};
            aa.P = e->P;
            aa.U = e->U;
            aa.frac = e->dFracSet;
            aa.state = e->dState;  // the amounts of null bricks do not change within a segment
            aa.lds = (long long)e->P * e->U;
            aa.sP = e->stateLaneP ? 1 : e->U;
            aa.sU = e->stateLaneP ? e->P : 1;
            aa.hruFlags = e->dHruFlags;
            aa.ldu = e->ldu;
            hbkStaleDeposits<<<e->P, 128, 0, e->stream>>>(aa, e->dStale);
            e->lastLaunches++;
            stale = e->dStale;
        }
        launchBasinStores(e, tBegin, tEnd, stale, writeOutputs ? 1 : 0);
        e->lastLaunches++;
    }
    HBK_CUDA(cudaGetLastError());
    return HBK_OK;
}

static int runImpl(hbk_engine* e, bool& enqueued);

int hbk_run(hbk_engine* e) {
    if (!e) return HBK_E_ARG;
    bool enqueued = false;  // work has been put on the engine's streams and the final synchronisation was not reached
    const int rc = runImpl(e, enqueued);
    if (rc != HBK_OK && enqueued) {
        // a launch / allocation error in the middle of a run: drain what is already on the streams (the getters must not
        // read a half-written run, the next call must not race with it) and forget the run
        const std::string msg = e->error;
        cudaStreamSynchronize(e->stream);
        for (auto& st : e->auxStream)
            if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
        e->error = msg;
        e->ran = false;
    }
    return rc;
}

static int runImpl(hbk_engine* e, bool& enqueued) {
    if (!e->unitsSet) return fail(e, HBK_E_STATE, "hbk_set_units has not been called");
    if (e->P <= 0) return fail(e, HBK_E_STATE, "hbk_set_parameters has not been called");
    bool tiled = true, station = true;
    for (int v = 0; v < e->nVars(); ++v) {
        tiled = tiled && e->haveRaw[v];
        station = station && e->haveStation[v];
    }
    if (!tiled && !station)
        return fail(e, HBK_E_STATE, e->nVars() == 4
                        ? "forcing incomplete: precipitation, temperature, pet and solar_radiation (melt:temperature_index) "
                          "are all required, either all per unit or all as station series"
                        : "forcing incomplete: precipitation, temperature and pet are all required, "
                          "either all per unit or all as station series");
    if (e->st.melt_kind == HBK_MELT_DEGREE_DAY_ASPECT && e->hAspect.empty())
        return fail(e, HBK_E_STATE, "melt:degree_day_aspect needs the aspect class of every hydro unit (hbk_set_unit_aspect)");
    if (station && e->perSetDropped)
        return fail(e, HBK_E_STATE, "per-set gradients / correction factors must be set again (hbk_set_station_forcing) "
                                    "after the number of parameter sets changed");
    const bool shardCb = e->shard && e->shardReduce != nullptr;
    if (e->shard && !shardCb && (e->spinupSteps > 0 || !e->events.empty()))
        return fail(e, HBK_E_UNSUPPORTED, "a unit shard runs spin-up and dated actions only with a reduction callback "
                                          "(hbk_set_shard_reduce): the sub-basin stores and the delta-h water equivalent "
                                          "span all shards");
    if (shardCb && e->dhTablesSet && !e->dhShardTotals) {
        for (const auto& ev : e->events)
            if (ev.kind == 2)
                return fail(e, HBK_E_STATE, "delta-h on a unit shard needs the row sums of the whole lookup table "
                                            "(hbk_set_delta_h_shard_totals)");
    }
    e->shardSegments.clear();
    if (e->st.snow_retention && (e->hasTwins || e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE))
        return fail(e, HBK_E_UNSUPPORTED, "liquid water in the snowpacks (outflow:snow_holding / refreeze:degree_day) runs "
                                          "neither with a second glacier cover nor with lateral connections");
    if (e->hasTwins && (e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE || e->shard))
        return fail(e, HBK_E_UNSUPPORTED, "units with a second glacier cover (hbk_set_unit_twins) run neither lateral "
                                          "connections nor unit shards");
    const bool lateral = e->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE;
    if (lateral) {
        if (!e->dLatOff)
            return fail(e, HBK_E_STATE, "transport:snow_slide needs the lateral connections (hbk_set_lateral_connections)");
        if (e->shard) return fail(e, HBK_E_UNSUPPORTED, "transport:snow_slide couples the units: no unit shards");
    }
    HBK_CUDA(cudaSetDevice(e->device));
    chooseShape(e);
    e->lastLaunches = 0;
    if (e->copyPending[e->outletCur]) {
        // an asynchronous fetch (hbk_get_outlet_async) is still reading the outlet of the last run: this run writes the
        // other buffer; should that one still be in flight too, the unit kernel waits for its copy on the device
        e->outletCur ^= 1;
        if (!e->dOutletBuf[e->outletCur])
            HBK_CUDA(cudaMalloc(&e->dOutletBuf[e->outletCur], sizeof(double) * (size_t)e->P * e->ldt));
        if (e->copyPending[e->outletCur]) {
            HBK_CUDA(cudaStreamWaitEvent(e->stream, e->copyDone[e->outletCur], 0));
            e->copyPending[e->outletCur] = false;
        }
        e->dOutlet = e->dOutletBuf[e->outletCur];
    }

    // (re)allocate shape-dependent buffers
    planRuns(e);
    const bool multi = e->totalGroups > 1;
    const size_t outN = (size_t)e->P * e->ldt;
    if (e->partGroups != e->totalGroups) {
        freeDev(e->dPartQ);
        freeDev(e->dPartD1);
        freeDev(e->dPartD2);
        e->partGroups = e->totalGroups;
    }
    if (multi && !e->dPartQ) HBK_CUDA(cudaMalloc(&e->dPartQ, sizeof(double) * outN * e->totalGroups));
    if (e->st.has_glacier) {
        if (!e->dD1) HBK_CUDA(cudaMalloc(&e->dD1, sizeof(double) * outN));
        if (!e->dD2) HBK_CUDA(cudaMalloc(&e->dD2, sizeof(double) * outN));
        if (multi && !e->dPartD1) HBK_CUDA(cudaMalloc(&e->dPartD1, sizeof(double) * outN * e->totalGroups));
        if (multi && !e->dPartD2) HBK_CUDA(cudaMalloc(&e->dPartD2, sizeof(double) * outN * e->totalGroups));

    }
    const int nRec = __builtin_popcountll(e->recordMask);
    const size_t recBytes = sizeof(double) * (size_t)nRec * e->P * e->T * e->U;
    if (recBytes != e->recBytes) {
        freeDev(e->dRec);
        if (recBytes) HBK_CUDA(cudaMalloc(&e->dRec, recBytes));
        e->recBytes = recBytes;
    }
    if (lateral) {
        if (e->latThrDirty || !e->dLatThr) {
            int rcT = uploadSlideThresholds(e);
            if (rcT) return rcT;
        }
        if (!e->dLatPend) HBK_CUDA(cudaMalloc(&e->dLatPend, sizeof(double) * 2 * (size_t)e->P * e->U));
        HBK_CUDA(cudaMemsetAsync(e->dLatPend, 0, sizeof(double) * 2 * (size_t)e->P * e->U, e->stream));
    }
    if (tiled && !lateral && e->tiledUB != e->UB) {
        for (int v = 0; v < e->nVars(); ++v) {
            freeDev(e->dForcingTiled[v]);
            const size_t n = (size_t)e->nUTiles * e->tld * e->UB;
            HBK_CUDA(cudaMalloc(&e->dForcingTiled[v], sizeof(double) * n));
            HBK_CUDA(cudaMemsetAsync(e->dForcingTiled[v], 0, sizeof(double) * n, e->stream));
            hbkTileForcing<<<148 * 4, 256, 0, e->stream>>>(e->dForcingRaw[v], e->dForcingTiled[v], e->T, e->U, e->UB, e->tld,
                                                           e->nUTiles);
        }
        HBK_CUDA(cudaGetLastError());
        e->tiledUB = e->UB;
    }

    enqueued = true;
    HBK_CUDA(cudaEventRecord(e->ev0, e->stream));
    // ModelHydro::Reset (ModelHydro.cpp:183-193): containers back to their initial state, fluxes to 0.
    const size_t nPU = (size_t)e->P * e->U;
    if (e->initPerSet) {
        HBK_CUDA(cudaMemcpyAsync(e->dState, e->dInitFull, sizeof(double) * e->nStates() * nPU, cudaMemcpyDeviceToDevice,
                                 e->stream));
        // flux amounts back to 0 (the slots between the storages and the two snowpack water storages at the end)
        HBK_CUDA(cudaMemsetAsync(e->dState + (size_t)HBK_S_AMT_INFILTRATION * nPU, 0,
                                 sizeof(double) * (std::min<int>(e->nStates(), HBK_S_GROUND_SNOWPACK_WATER) - HBK_S_AMT_INFILTRATION) * nPU,
                                 e->stream));
    } else {
        for (int s = 0; s < e->nStates(); ++s) {
            const bool storage = s < HBK_S_AMT_INFILTRATION || s >= HBK_S_GROUND_SNOWPACK_WATER;
            hbkBroadcastState<<<148 * 2, 256, 0, e->stream>>>(e->dState + (size_t)s * nPU,
                                                              storage ? e->dInitUnit + (size_t)s * e->U : nullptr, 0.0,
                                                              e->P, e->U, e->stateLaneP ? 1 : 0);
        }
        e->lastLaunches += e->nStates();
    }
    HBK_CUDA(cudaMemcpyAsync(e->dBasinState, e->dBasinInit, sizeof(double) * 2 * e->P, cudaMemcpyDeviceToDevice, e->stream));
    HBK_CUDA(cudaMemcpyAsync(e->dFrac, e->dFracInit, sizeof(double) * 2 * e->U, cudaMemcpyDeviceToDevice, e->stream));
    e->fracPerSet = e->events.empty() ? 0 : 1;
    if (e->fracPerSet) {  // LandCover::Reset + ActionsManager::Reset: initial extents, lookup-table row 0
        if (!e->dFracSet) HBK_CUDA(cudaMalloc(&e->dFracSet, sizeof(double) * 2 * nPU));
        if (!e->dDhLastRow) HBK_CUDA(cudaMalloc(&e->dDhLastRow, sizeof(int) * e->P));
        hbkBroadcastFractions<<<148 * 2, 256, 0, e->stream>>>(e->dFracSet, e->dFracInit, e->P, e->U);
        HBK_CUDA(cudaMemsetAsync(e->dDhLastRow, 0, sizeof(int) * e->P, e->stream));
        e->lastLaunches++;
    }
    HBK_CUDA(cudaMemsetAsync(e->dErr, 0, sizeof(int), e->stream));
    HBK_CUDA(cudaMemsetAsync(e->dUnconv, 0, sizeof(unsigned long long), e->stream));
    HBK_CUDA(cudaMemsetAsync(e->dUnconvSet, 0, sizeof(int) * e->P, e->stream));

    int rc;
    if (e->spinupSteps > 0) {  // ModelHydro::RunSpinup (ModelHydro.cpp:135-181)
        rc = launchSegment(e, 0, e->spinupSteps, false);
        if (rc) return rc;
        if (shardCb && e->st.has_glacier) {
            // the sub-basin stores are spun up on the deposits of the whole basin
            rc = shardReduceBlocks(e, {{e->dD1, e->ldt, 0, e->spinupSteps, 0}, {e->dD2, e->ldt, 0, e->spinupSteps, 0}});
            if (rc) return rc;
            launchBasinStores(e, 0, e->spinupSteps, nullptr, 0);
            e->lastLaunches++;
        }
    }
    if (e->events.empty()) {
        rc = launchSegment(e, 0, e->T, true);
        if (rc) return rc;
    } else {
        // kernel boundaries at the action dates (ActionsManager::DateUpdate between two time steps)
        std::vector<int> steps;
        for (const auto& ev : e->events)
            if (ev.step >= 1 && ev.step < e->T) steps.push_back(ev.step);
        std::sort(steps.begin(), steps.end());
        steps.erase(std::unique(steps.begin(), steps.end()), steps.end());
        int t = 0;
        for (int k : steps) {
            rc = launchSegment(e, t, k, true);
            if (rc) return rc;
            rc = applyEvents(e, k);
            if (rc) return rc;
            t = k;
        }
        rc = launchSegment(e, t, e->T, true);
        if (rc) return rc;
    }
    if (shardCb) {
        // end of the unit pass: partial outlet, deposits and stale amounts summed over the shards in one collective, then
        // the sub-basin stores segment by segment on the sums (exact: no superposition across shards)
        std::vector<ShardBlock> blocks = {{e->dOutlet, e->ldt, 0, e->T, 0}};
        if (e->st.has_glacier) {
            blocks.push_back({e->dD1, e->ldt, 0, e->T, 0});
            blocks.push_back({e->dD2, e->ldt, 0, e->T, 0});
            if (e->fracPerSet) blocks.push_back({e->dStaleSeg, 0, 0, 0, e->shardSegments.size() * 2 * (size_t)e->P});
        }
        rc = shardReduceBlocks(e, blocks);
        if (rc) return rc;
        if (e->st.has_glacier) {
            for (size_t seg = 0; seg < e->shardSegments.size(); ++seg) {
                launchBasinStores(e, e->shardSegments[seg].tBegin, e->shardSegments[seg].tEnd,
                                  e->fracPerSet ? e->dStaleSeg + seg * 2 * (size_t)e->P : nullptr, 1);
                e->lastLaunches++;
            }
        }
        HBK_CUDA(cudaGetLastError());
    }
    HBK_CUDA(cudaEventRecord(e->ev1, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    enqueued = false;  // the run is complete: errors from here on are findings about its results
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->lastMs = ms;
    int err = 0;
    unsigned long long unc = 0;
    HBK_CUDA(cudaMemcpy(&err, e->dErr, sizeof(int), cudaMemcpyDeviceToHost));
    HBK_CUDA(cudaMemcpy(&unc, e->dUnconv, sizeof(unc), cudaMemcpyDeviceToHost));
    e->lastUnconverged = (long long)unc;
    e->ran = true;
    e->shardPending = e->shard && !shardCb;
    if (err & kErrRateTooHigh)  // WaterContainer.cpp:83-86 throws; ModelHydro::Run fails
        return fail(e, HBK_E_MODEL, "a change rate exceeded 10000 (WaterContainer::ApplyConstraints)");
    if (err & kErrQueueStall) return fail(e, HBK_E_CUDA, "internal error: the unit kernel's work queue stalled");
    if (err & kErrActionFraction)
        return fail(e, HBK_E_MODEL, "Application of an action failed: land-cover fraction outside [0 .. 1] or the "
                                    "generic land cover cannot absorb the area change");
    if (err & kErrActionInfiniteIce)
        return fail(e, HBK_E_MODEL, "Trying to set the content of an infinite storage (snow-to-ice on infinite glacier ice)");
    return HBK_OK;
}

int hbk_set_unit_shard(hbk_engine* e, double basin_area_m2) {
    if (!e) return HBK_E_ARG;
    if (!e->unitsSet) return fail(e, HBK_E_STATE, "hbk_set_units has not been called");
    HBK_CUDA(cudaSetDevice(e->device));
    e->shard = basin_area_m2 > 0;
    double basinArea = basin_area_m2;
    if (!e->shard) {
        basinArea = 0;
        for (double a : e->hArea) basinArea += a;
    }
    e->shardBasinArea = basinArea;
    e->shardPending = false;
    std::vector<double> w(e->U);
    for (int u = 0; u < e->U; ++u) w[u] = e->hArea[u] / basinArea;  // ModelBuilder.cpp:468, whole-basin area
    HBK_CUDA(cudaMemcpy(e->dHru, w.data(), sizeof(double) * e->U, cudaMemcpyHostToDevice));
    return HBK_OK;
}

int hbk_set_shard_reduce(hbk_engine* e, hbk_shard_reduce_fn reduce, void* user) {
    if (!e) return HBK_E_ARG;
    e->shardReduce = reduce;
    e->shardUser = user;
    return HBK_OK;
}

int hbk_shard_partials_dev(hbk_engine* e, double** outlet, double** deposits_rain_snowmelt, double** deposits_icemelt,
                           int64_t* n_rows, int64_t* ld) {
    if (!e) return HBK_E_ARG;
    if (!e->shard) return fail(e, HBK_E_STATE, "not a unit shard (hbk_set_unit_shard)");
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    if (outlet) *outlet = e->dOutlet;
    if (deposits_rain_snowmelt) *deposits_rain_snowmelt = e->st.has_glacier ? e->dD1 : nullptr;
    if (deposits_icemelt) *deposits_icemelt = e->st.has_glacier ? e->dD2 : nullptr;
    if (n_rows) *n_rows = e->P;
    if (ld) *ld = e->ldt;
    return HBK_OK;
}

int hbk_finish_sharded_run(hbk_engine* e) {
    if (!e) return HBK_E_ARG;
    if (!e->shard) return fail(e, HBK_E_STATE, "not a unit shard (hbk_set_unit_shard)");
    if (!e->ran || !e->shardPending) return fail(e, HBK_E_STATE, "hbk_run has not been called since the last finish");
    HBK_CUDA(cudaSetDevice(e->device));
    e->shardPending = false;
    if (!e->st.has_glacier) return HBK_OK;  // no sub-basin stores: the reduced unit sums are the outlet
    HBK_CUDA(cudaEventRecord(e->ev0, e->stream));
    launchBasinStores(e, 0, e->T, nullptr, 1);
    e->lastLaunches++;
    HBK_CUDA(cudaGetLastError());
    HBK_CUDA(cudaEventRecord(e->ev1, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->lastMs += ms;
    int err = 0;
    HBK_CUDA(cudaMemcpy(&err, e->dErr, sizeof(int), cudaMemcpyDeviceToHost));
    if (err & kErrRateTooHigh)
        return fail(e, HBK_E_MODEL, "a change rate exceeded 10000 (WaterContainer::ApplyConstraints)");
    return HBK_OK;
}

int hbk_get_outlet(hbk_engine* e, int32_t first, int32_t n, double* out) {
    if (!e || !out || first < 0 || n <= 0 || first + n > e->P) return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    HBK_CUDA(cudaMemcpy2D(out, sizeof(double) * e->T, e->dOutlet + (size_t)first * e->ldt, sizeof(double) * e->ldt,
                          sizeof(double) * e->T, n, cudaMemcpyDeviceToHost));
    return HBK_OK;
}

int hbk_get_outlet_async(hbk_engine* e, int32_t first, int32_t n, double* out) {
    if (!e || !out || first < 0 || n <= 0 || first + n > e->P) return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    if (!e->copyStream) {
        HBK_CUDA(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
        HBK_CUDA(cudaEventCreateWithFlags(&e->copyDone[0], cudaEventDisableTiming));
        HBK_CUDA(cudaEventCreateWithFlags(&e->copyDone[1], cudaEventDisableTiming));
    }
    // hbk_run has drained the compute stream, so the outlet is complete; one strided copy on the copy stream
    HBK_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * e->T, e->dOutlet + (size_t)first * e->ldt, sizeof(double) * e->ldt,
                               sizeof(double) * e->T, n, cudaMemcpyDeviceToHost, e->copyStream));
    HBK_CUDA(cudaEventRecord(e->copyDone[e->outletCur], e->copyStream));
    e->copyPending[e->outletCur] = true;
    return HBK_OK;
}

int hbk_wait_outlet(hbk_engine* e) {
    if (!e) return HBK_E_ARG;
    if (!e->copyStream) return HBK_OK;
    HBK_CUDA(cudaSetDevice(e->device));
    HBK_CUDA(cudaStreamSynchronize(e->copyStream));
    e->copyPending[0] = e->copyPending[1] = false;
    return HBK_OK;
}

int hbk_get_recorded(hbk_engine* e, int32_t slot, int32_t set, double* out) {
    if (!e || !out || slot < 0 || slot >= HBK_N_RECORDS || set < 0 || set >= e->P) return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    if (!(e->recordMask & (1ull << slot))) return fail(e, HBK_E_ARG, "this item was not recorded (record mask)");
    HBK_CUDA(cudaSetDevice(e->device));
    const int k = __builtin_popcountll(e->recordMask & ((1ull << slot) - 1));
    const size_t n = (size_t)e->T * e->U;
    HBK_CUDA(cudaMemcpy(out, e->dRec + ((size_t)k * e->P + set) * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return HBK_OK;
}

int hbk_get_state(hbk_engine* e, int32_t slot, double* out) {
    if (!e || !out || slot < 0 || slot >= HBK_N_STATES) return HBK_E_ARG;
    if (slot >= e->nStates()) return fail(e, HBK_E_ARG, "this state slot exists only with liquid water in the snowpacks");
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    const size_t n = (size_t)e->P * e->U;
    if (!e->stateLaneP && !e->dPerm) {
        HBK_CUDA(cudaMemcpy(out, e->dState + (size_t)slot * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    } else {  // device layout [u][p] or [p][u] in the internal set order -> out[caller's p][u]
        std::vector<double> tmp(n);
        HBK_CUDA(cudaMemcpy(tmp.data(), e->dState + (size_t)slot * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
        const size_t sP = e->stateLaneP ? 1 : e->U, sU = e->stateLaneP ? e->P : 1;
        for (int p = 0; p < e->P; ++p) {
            const size_t row = e->dPerm ? e->perm[p] : p;
            for (int u = 0; u < e->U; ++u) out[row * e->U + u] = tmp[p * sP + u * sU];
        }
    }
    return HBK_OK;
}

int hbk_get_basin_state(hbk_engine* e, double* out) {
    if (!e || !out) return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    if (!e->dPerm) {
        HBK_CUDA(cudaMemcpy(out, e->dBasinState, sizeof(double) * 2 * e->P, cudaMemcpyDeviceToHost));
    } else {
        std::vector<double> tmp(2 * (size_t)e->P);
        HBK_CUDA(cudaMemcpy(tmp.data(), e->dBasinState, sizeof(double) * 2 * e->P, cudaMemcpyDeviceToHost));
        for (int p = 0; p < e->P; ++p) {
            out[2 * (size_t)e->perm[p]] = tmp[2 * (size_t)p];
            out[2 * (size_t)e->perm[p] + 1] = tmp[2 * (size_t)p + 1];
        }
    }
    return HBK_OK;
}

int hbk_outlet_dev(hbk_engine* e, const double** ptr, int64_t* ld) {
    if (!e || !ptr || !ld) return HBK_E_ARG;
    *ptr = e->dOutlet;
    *ld = e->ldt;
    return HBK_OK;
}

int hbk_last_run_stats(hbk_engine* e, double* ms, int32_t* launches, int64_t* unconverged) {
    if (!e) return HBK_E_ARG;
    if (ms) *ms = e->lastMs;
    if (launches) *launches = e->lastLaunches;
    if (unconverged) *unconverged = e->lastUnconverged;
    return HBK_OK;
}

int hbk_last_run_shape(hbk_engine* e, int32_t* out) {
    if (!e || !out) return HBK_E_ARG;
    if (e->runs.empty()) return fail(e, HBK_E_STATE, "hbk_last_run_shape needs a run");
    int plainTiles = 0;
    for (const auto& r : e->runs)
        if (e->st.has_glacier && !r.glacier) plainTiles += r.tile1 - r.tile0;
    const int32_t v[8] = {e->PB, e->UB, e->nUTiles, (int32_t)e->runs.size(), e->runs[0].TC, e->runs[0].G, e->totalGroups,
                          plainTiles};
    std::copy(v, v + 8, out);
    return HBK_OK;
}

int hbk_get_unconverged(hbk_engine* e, int32_t* out) {
    if (!e || !out) return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    std::vector<int> tmp(e->P);
    HBK_CUDA(cudaMemcpy(tmp.data(), e->dUnconvSet, sizeof(int) * e->P, cudaMemcpyDeviceToHost));
    for (int p = 0; p < e->P; ++p) out[e->dPerm ? e->perm[p] : p] = tmp[p];
    return HBK_OK;
}

void* hbk_stream(hbk_engine* e) { return e ? (void*)e->stream : nullptr; }

int hbk_compute_scores(hbk_engine* e, int32_t metric, const double* observed, int32_t warmup, double* out) {
    if (!e || !observed || warmup < 0 || warmup >= e->T || metric < 0 || metric >= HBK_N_SCORES)
        return HBK_E_ARG;
    if (!e->ran) return fail(e, HBK_E_STATE, "no completed run");
    HBK_CUDA(cudaSetDevice(e->device));
    if (!e->dObs) HBK_CUDA(cudaMalloc(&e->dObs, sizeof(double) * (e->T + 4)));
    if (!e->dScores) HBK_CUDA(cudaMalloc(&e->dScores, sizeof(double) * e->P));
    HBK_CUDA(cudaMemcpyAsync(e->dObs, observed, sizeof(double) * e->T, cudaMemcpyHostToDevice, e->stream));
    hbkObsStats<<<1, 256, 0, e->stream>>>(e->dObs, warmup, e->T, e->dObs + e->T);
    hbkScores<<<e->P, 256, 0, e->stream>>>(e->dOutlet, e->ldt, e->dObs, warmup, e->T, e->dObs + e->T, metric, e->dScores);
    HBK_CUDA(cudaGetLastError());
    if (out) HBK_CUDA(cudaMemcpyAsync(out, e->dScores, sizeof(double) * e->P, cudaMemcpyDeviceToHost, e->stream));
    HBK_CUDA(cudaStreamSynchronize(e->stream));
    return HBK_OK;
}

int hbk_scores_dev(hbk_engine* e, const double** ptr) {
    if (!e || !ptr) return HBK_E_ARG;
    *ptr = e->dScores;
    return HBK_OK;
}

int hbk_selftest_sqrt(int32_t device, int64_t n, int64_t* mismatches) {
    if (!mismatches || n <= 0) return HBK_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return HBK_E_CUDA;
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, sizeof(*d)) != cudaSuccess || cudaMemset(d, 0, sizeof(*d)) != cudaSuccess) return HBK_E_CUDA;
    hbkSqrtCheck<<<148 * 8, 256>>>((long long)n, d);
    unsigned long long h = 0;
    const bool ok = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d);
    if (!ok) return HBK_E_CUDA;
    *mismatches = (int64_t)h;
    return HBK_OK;
}

int hbk_measure_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return HBK_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return HBK_E_CUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return HBK_E_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int iters = 1 << 16, blocks = 148 * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        hbkFp64Peak<<<blocks, threads>>>(d, iters, 1.0 + rep);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) {
            cudaFree(d);
            return HBK_E_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return HBK_OK;
}

}  // extern "C"
