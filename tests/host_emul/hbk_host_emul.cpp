// tests/host_emul/hbk_host_emul.cpp — TEST INFRASTRUCTURE ONLY.
//
// Runs the per-(parameter set, hydro unit) arithmetic of the CUDA kernels — hydrobricks_b200/csrc/hbk_device.cuh, the
// same source, compiled by g++ with -DHBK_HOST_EMULATION — over one parameter set on the CPU, with the thread-level
// glue of hbkRunUnits / hbkBasinStores (hbk_engine.cu) restated in plain loops: parameter promotion, per-unit
// constants, the time loop (directPhase -> solveUnit | solveUnitSequential), the outlet sum over the units in unit
// order and the two sub-basin stores. It lets `pytest -m "not gpu"` check the device arithmetic against the reference's
// golden vectors at the 1e-10 bar without a GPU; what it does NOT cover is the kernel plumbing (tiles, TMA staging,
// state streaming, reductions), which only the `-m gpu` tests exercise. Never linked into the product.
#include "../../include/hbk.h"
static long long hbkSeqEvaluations = 0;  // counted by seqBrickRates (rate-law evaluations inside the bisection)
#include "../../hydrobricks_b200/csrc/hbk_device.cuh"

using namespace hbk;

namespace {

template <int SOLVER, int NSOIL, bool GLACIER>
int runSet(const hbk_structure& st, int U, int T, const double* area, const float* slope, const int32_t* glacierVariant,
           const double* fracG, const double* fracGl, const float* pp, const double* precip, const double* temp,
           const double* pet, const double* swat, double* outlet, double* recSlow, double* recSnowG, double* recIce,
           long long* unconvergedOut, const int32_t* aspect, const double* radiation) {
    Flags f{};
    f.quickKind = st.quick_kind;
    f.slowOrder = st.slow_process_order;
    f.splitKind = st.splitter_kind;
    f.iceInfinite = st.glacier_infinite_storage;
    f.noMeltWhenSnow = st.glacier_no_melt_when_snow_cover;
    f.snowIceKind = st.has_glacier ? st.snow_ice_kind : 0;
    f.sublimKind = st.sublimation_kind;
    f.seqKind = st.solver >= HBK_SOLVER_CRANK_NICOLSON ? st.solver - HBK_SOLVER_CRANK_NICOLSON : 0;
    f.retention = st.snow_retention;
    const double dt = st.time_step_days, invDt = 1.0 / st.time_step_days;

    // parameter promotion as in hbkRunUnits (float32 -> double where the reference's expressions promote)
    Par par{};
    const float t0f = pp[HBK_P_TRANSITION_START], t1f = pp[HBK_P_TRANSITION_END];
    par.v[0] = t0f;
    par.v[1] = t1f;
    par.v[2] = 1.0 / (double)(t1f - t0f);
    par.v[3] = pp[HBK_P_RAIN_CORRECTION];
    par.v[4] = pp[HBK_P_SNOW_CORRECTION];
    par.v[5] = pp[HBK_P_GROUND_SNOW_DDF];
    par.v[6] = pp[HBK_P_GROUND_SNOW_TMELT];
    par.v[7] = pp[HBK_P_GLACIER_SNOW_DDF];
    par.v[8] = pp[HBK_P_GLACIER_SNOW_TMELT];
    par.v[9] = pp[HBK_P_ICE_DDF];
    par.v[10] = pp[HBK_P_ICE_TMELT];
    const double capacity = pp[HBK_P_CAPACITY];
    par.v[11] = capacity;
    par.v[12] = (capacity == 0.0) ? kInvZeroCapacity : 1.0 / capacity;
    par.v[13] = pp[HBK_P_K_SLOW_1];
    par.v[14] = pp[HBK_P_PERCOLATION];
    par.v[15] = pp[HBK_P_K_SLOW_2];
    par.v[16] = pp[HBK_P_SNOW_ICE];
    const double quickBase = pp[HBK_P_QUICK];

    double basinArea = 0;  // SubBasin::AddHydroUnit (SubBasin.cpp:173), as hbk_set_units
    for (int u = 0; u < U; ++u) basinArea += area[u];

    std::vector<Hru> state(U, Hru{});
    std::vector<double> d1(T, 0.0), d2(T, 0.0);
    int err = 0, unconverged = 0;
    for (int t = 0; t < T; ++t) outlet[t] = 0.0;
    for (int u = 0; u < U; ++u) {
        Hru& h = state[u];
        const double fa = area[u] / basinArea;
        const double faW = (fa < 1.0) ? fa : 1.0;  // weight of the outlet fluxes (Flux.cpp:20-26)
        const double slopeTerm = slope ? std::pow((double)slope[u], 0.5) : 0.0;
        const double socontUnit = 3.174802103936398e-05 * 86400.0 * 1000.0;
        par.quickV = (f.quickKind == 0) ? quickBase * slopeTerm / area[u] * socontUnit : quickBase;
        const bool gv = GLACIER && glacierVariant && glacierVariant[u];
        const double fG = fracG ? fracG[u] : 1.0, fGl = fracGl ? fracGl[u] : 0.0;
        // Melt-process variants: the glue of hbkRunUnits (hbk_engine.cu loadUnit / time loop) around the unchanged step
        // functions.
        //  melt:degree_day_aspect  -> the degree-day factors of the unit's aspect class (slots DDF | MELT_B | MELT_C)
        //                             replace the three per-set factors when the unit is loaded;
        //  melt:temperature_index  -> per step, factor = melt_factor + radiation_coefficient * solar radiation
        //                             (ProcessMeltTemperatureIndex.cpp:68-71), radiation[T][U] the fourth forcing series.
        if (st.melt_kind == HBK_MELT_DEGREE_DAY_ASPECT) {
            const int a = aspect[u];
            par.v[5] = pp[a == 0 ? HBK_P_GROUND_SNOW_DDF : (a == 1 ? HBK_P_GROUND_SNOW_MELT_B : HBK_P_GROUND_SNOW_MELT_C)];
            par.v[7] = pp[a == 0 ? HBK_P_GLACIER_SNOW_DDF : (a == 1 ? HBK_P_GLACIER_SNOW_MELT_B : HBK_P_GLACIER_SNOW_MELT_C)];
            par.v[9] = pp[a == 0 ? HBK_P_ICE_DDF : (a == 1 ? HBK_P_ICE_MELT_B : HBK_P_ICE_MELT_C)];
        }
        // liquid water in the snowpacks (HBK_RETENTION_*): the unit's six extra state values, as retentionCtx of
        // hbk_engine.cu points them into the state arrays
        double extra[6] = {0, 0, 0, 0, 0, 0};
        RetentionCtx rc{};
        rc.g = {&extra[0], &extra[1], &extra[2], (double)pp[HBK_P_GROUND_SNOW_WHC], (double)pp[HBK_P_GROUND_SNOW_CFR]};
        rc.gl = {&extra[3], &extra[4], &extra[5], (double)pp[HBK_P_GLACIER_SNOW_WHC], (double)pp[HBK_P_GLACIER_SNOW_CFR]};
        rc.refreeze = (st.snow_retention & HBK_RETENTION_REFREEZE) != 0;
        rc.rainToSnowpack = (st.snow_retention & HBK_RETENTION_RAIN_TO_SNOWPACK) != 0;
        const RetentionCtx* ret = st.snow_retention ? &rc : nullptr;
        for (int t = 0; t < T; ++t) {
            StepOut o{};
            const double p = precip[(size_t)t * U + u], tt = temp[(size_t)t * U + u], e = pet[(size_t)t * U + u];
            if (st.melt_kind == HBK_MELT_TEMPERATURE_INDEX) {
                const double rad = radiation[(size_t)t * U + u];
                par.v[5] = (double)pp[HBK_P_GROUND_SNOW_DDF] + (double)pp[HBK_P_GROUND_SNOW_MELT_B] * rad;
                par.v[7] = (double)pp[HBK_P_GLACIER_SNOW_DDF] + (double)pp[HBK_P_GLACIER_SNOW_MELT_B] * rad;
                par.v[9] = (double)pp[HBK_P_ICE_DDF] + (double)pp[HBK_P_ICE_MELT_B] * rad;
            }
            double sublimG = 0.0, sublimGl = 0.0;
            if (f.sublimKind != 0) {
                const double scale = (f.sublimKind == HBK_SUBLIMATION_PET) ? e : 1.0;
                sublimG = scale * (double)pp[HBK_P_GROUND_SUBLIMATION];
                if (GLACIER) sublimGl = scale * (double)pp[HBK_P_GLACIER_SUBLIMATION];
            }
            directPhase<GLACIER>(h, par, p, tt, dt, invDt, fa, fG, fGl, gv, f, o, err, (GLACIER && swat) ? swat[t] : 0.0,
                                 sublimG, sublimGl, 0.0, 0.0, NoSlide(), NoSlide(), ret);
            if (SOLVER == 3) {
                solveUnitSequential<NSOIL>(h, par, e, faW, dt, invDt, f, o, err);
            } else {
                solveUnit<(SOLVER == 3 ? 0 : SOLVER), NSOIL>(h, par, e, faW, dt, invDt, f, o, err, unconverged);
            }
            outlet[t] += o.q;  // units in unit order (the kernel sums per tile, then the tiles in order)
            d1[t] += o.dep1;
            d2[t] += o.dep2;
            if (recSlow) recSlow[(size_t)t * U + u] = h.slow;
            if (recSnowG) recSnowG[(size_t)t * U + u] = h.snowG;
            if (recIce) recIce[(size_t)t * U + u] = gv ? h.ice : NAN;
        }
    }
    if (GLACIER) {  // hbkBasinStores
        const double kSnow = pp[HBK_P_K_SNOW], kIce = pp[HBK_P_K_ICE];
        double c1 = 0.0, c2 = 0.0;
        for (int t = 0; t < T; ++t) {
            const double o1 = basinStoreStep<SOLVER>(c1, kSnow, d1[t], d1[t], dt, err, unconverged, f.seqKind);
            const double o2 = basinStoreStep<SOLVER>(c2, kIce, d2[t], d2[t], dt, err, unconverged, f.seqKind);
            outlet[t] = (o1 + o2) + outlet[t];
        }
    }
    if (unconvergedOut) *unconvergedOut = unconverged;
    return err;
}

}  // namespace

#define HBK_EMUL_ARGS                                                                                               \
    *st, n_units, n_steps, area, slope, glacier_variant, frac_ground, frac_glacier, params, precip, temp, pet, swat, \
        outlet, rec_slow, rec_snow_ground, rec_ice, unconverged, g_aspect, g_radiation

// melt-variant inputs of the next call (hbk_emul_set_melt_inputs): HBK_ASPECT_* per unit (melt:degree_day_aspect) and
// the solar radiation [n_steps][n_units] (melt:temperature_index); reset after every run
static const int32_t* g_aspect = nullptr;
static const double* g_radiation = nullptr;

extern "C" long long hbk_emul_seq_evaluations(int reset) {
    const long long n = hbkSeqEvaluations;
    if (reset) hbkSeqEvaluations = 0;
    return n;
}

extern "C" void hbk_emul_set_melt_inputs(const int32_t* aspect, const double* radiation) {
    g_aspect = aspect;
    g_radiation = radiation;
}

template <int SOLVER, int NSOIL>
static int pickGlacier(const hbk_structure* st, int n_units, int n_steps, const double* area, const float* slope,
                       const int32_t* glacier_variant, const double* frac_ground, const double* frac_glacier,
                       const float* params, const double* precip, const double* temp, const double* pet,
                       const double* swat, double* outlet, double* rec_slow, double* rec_snow_ground, double* rec_ice,
                       long long* unconverged) {
    return st->has_glacier ? runSet<SOLVER, NSOIL, true>(HBK_EMUL_ARGS) : runSet<SOLVER, NSOIL, false>(HBK_EMUL_ARGS);
}

// One parameter set (params[HBK_N_PARAMS]) over the whole basin; forcing [n_steps][n_units]; swat: the [n_steps]
// day-of-year factor of transform:snow_ice_swat or NULL. Returns the OR of the kernel's error bits.
extern "C" int hbk_emul_run(const hbk_structure* st, int n_units, int n_steps, const double* area, const float* slope,
                            const int32_t* glacier_variant, const double* frac_ground, const double* frac_glacier,
                            const float* params, const double* precip, const double* temp, const double* pet,
                            const double* swat, double* outlet, double* rec_slow, double* rec_snow_ground,
                            double* rec_ice, long long* unconverged) {
#define HBK_EMUL_SOIL(S)                                                            \
    (st->n_soil_stores == 2 ? pickGlacier<S, 2>(st, n_units, n_steps, area, slope, glacier_variant, frac_ground,   \
                                                frac_glacier, params, precip, temp, pet, swat, outlet, rec_slow,   \
                                                rec_snow_ground, rec_ice, unconverged)                              \
                            : pickGlacier<S, 1>(st, n_units, n_steps, area, slope, glacier_variant, frac_ground,   \
                                                frac_glacier, params, precip, temp, pet, swat, outlet, rec_slow,   \
                                                rec_snow_ground, rec_ice, unconverged))
    int rc;
    switch (st->solver) {
        case HBK_SOLVER_EULER: rc = HBK_EMUL_SOIL(0); break;
        case HBK_SOLVER_HEUN: rc = HBK_EMUL_SOIL(1); break;
        case HBK_SOLVER_RK4: rc = HBK_EMUL_SOIL(2); break;
        default: rc = HBK_EMUL_SOIL(3); break;
    }
    hbk_emul_set_melt_inputs(nullptr, nullptr);
    return rc;
}
