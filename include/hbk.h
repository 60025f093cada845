/* include/hbk.h — C-ABI of the B200-native batched simulation engine for hydrobricks.
 *
 * Drop-in scope: the per-time-step solve behind ModelHydro::Run() (reference
 * core/src/base/ModelHydro.cpp:100-133 and callees) for Socont-family structures
 * with the explicit solvers, batched over (parameter set x hydro unit). The reference has no FFI of
 * its own for this path (its boundary is the pybind11 module core/bindings/Binding.cpp:308-360);
 * each entry point below names the ModelHydro / SettingsModel member it replaces. The pybind11
 * module `_hydrobricks` of this repo (hydrobricks_b200/csrc/hb_binding.cpp) re-implements those
 * classes on top of this ABI; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: every call returns 0 on success or a negative HBK_E_* code; hbk_last_error() gives the
 * message. Handles are opaque. All buffers are caller-owned HOST memory unless the name says `_dev`.
 * One engine per CUDA device; an engine is not thread-safe, distinct engines are independent.
 * There is NO CPU fallback: every call fails with HBK_E_CUDA when no CUDA device is usable.
 */
#ifndef HBK_H
#define HBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hbk_engine hbk_engine;

enum { HBK_OK = 0, HBK_E_ARG = -1, HBK_E_CUDA = -2, HBK_E_STATE = -3, HBK_E_MODEL = -4, HBK_E_UNSUPPORTED = -5 };

/* Solver::Factory names (core/src/base/Solver.cpp:42-64): euler_explicit, heun_explicit, rk4|runge_kutta, and the
 * sequential family of base/SolverSequential.cpp (one brick after the other): crank_nicolson|
 * trapezoidal (the default of the reference's Python layer, models/model.py:80), implicit_euler|euler_implicit,
 * exponential_euler, analytic_linear|analytic */
enum { HBK_SOLVER_EULER = 0, HBK_SOLVER_HEUN = 1, HBK_SOLVER_RK4 = 2, HBK_SOLVER_CRANK_NICOLSON = 3,
       HBK_SOLVER_IMPLICIT_EULER = 4, HBK_SOLVER_EXPONENTIAL_EULER = 5, HBK_SOLVER_ANALYTIC_LINEAR = 6 };
/* surface_runoff brick process: runoff:socont (ProcessRunoffSocont.cpp:41-56) | outflow:linear */
enum { HBK_QUICK_SOCONT = 0, HBK_QUICK_LINEAR = 1 };
/* snow_rain_transition splitter: snow_rain:linear | snow_rain:threshold */
enum { HBK_SPLIT_LINEAR = 0, HBK_SPLIT_THRESHOLD = 1 };
/* order of the slow_reservoir processes = rate-vector layout (Processor.cpp:146-172) */
enum { HBK_SLOW_ORDER_ET_OUT_OVF_PERC = 0 /* models/socont.py:171-190 */,
       HBK_SLOW_ORDER_ET_OUT_PERC_OVF = 1 /* core/tests/src/helpers.cpp:74-86 */ };
/* forcing variables (TimeSeries.cpp:134-146); solar radiation only feeds melt:temperature_index */
enum { HBK_VAR_PRECIPITATION = 0, HBK_VAR_TEMPERATURE = 1, HBK_VAR_PET = 2, HBK_VAR_SOLAR_RADIATION = 3, HBK_N_VARS = 4 };
/* glacier snowpack process after the melt: none | transform:snow_ice_constant | transform:snow_ice_swat */
enum { HBK_SNOW_ICE_NONE = 0, HBK_SNOW_ICE_CONSTANT = 1, HBK_SNOW_ICE_SWAT = 2 };
/* snowpack -> atmosphere: none | sublimation:constant | sublimation:pet (ProcessSublimation{Constant,PET}.cpp) */
enum { HBK_SUBLIMATION_NONE = 0, HBK_SUBLIMATION_CONSTANT = 1, HBK_SUBLIMATION_PET = 2 };
/* melt process of every snowpack and of the glacier ice (the reference's Socont option snow_melt_process,
 * models/socont.py:68,141 -> modules/glacier.py:87-131): melt:degree_day (ProcessMeltDegreeDay.cpp:49-60) |
 * melt:degree_day_aspect (ProcessMeltDegreeDayAspect.cpp:38-80: the degree-day factor of the unit's aspect class) |
 * melt:temperature_index (ProcessMeltTemperatureIndex.cpp:62-74: factor = melt_factor + radiation_coefficient x solar
 * radiation, a fourth forcing series) */
enum { HBK_MELT_DEGREE_DAY = 0, HBK_MELT_DEGREE_DAY_ASPECT = 1, HBK_MELT_TEMPERATURE_INDEX = 2 };
/* aspect class of a hydro unit (HydroUnit::GetPropertyString("aspect_class")): north|N, south|S, east|west|E|W */
enum { HBK_ASPECT_NORTH = 0, HBK_ASPECT_SOUTH = 1, HBK_ASPECT_EAST_WEST = 2 };
/* snow redistribution (SettingsModel::AddSnowRedistribution, SettingsModel.cpp:560-572): none | transport:snow_slide on
 * the snowpacks of the non-glacier covers (skipGlaciers, the default of the reference's Socont) | on every snowpack.
 * ProcessLateralSnowSlide.cpp:46-134: snow above a slope-dependent holding depth slides, within the step, onto the
 * snowpacks of the units the giving unit is laterally connected to (hbk_set_lateral_connections). */
enum { HBK_SNOW_SLIDE_NONE = 0, HBK_SNOW_SLIDE_SKIP_GLACIERS = 1, HBK_SNOW_SLIDE_ALL = 2 };
/* snowpack liquid water: the melt water stays in the snowpack's water container and outflow:snow_holding releases what
 * exceeds water_holding_capacity x snow water equivalent (ProcessOutflowSnowHolding.cpp:37-57) | + refreeze:degree_day
 * moves it back to the snow (ProcessRefreezeDegreeDay.cpp:72-87) | + the rain goes through the snowpack
 * (ChangeSplitterOutputTargetIfFound, SettingsModel.cpp:626-632) */
enum { HBK_RETENTION_HOLDING = 1, HBK_RETENTION_REFREEZE = 2, HBK_RETENTION_RAIN_TO_SNOWPACK = 4 };

/* The per-structure process table for the Socont family, as the structure compiler
 * (hydrobricks_b200/csrc/hb_host.hpp, hb::Compiled) derives it from a SettingsModel. */
typedef struct {
    int32_t solver;
    int32_t n_soil_stores;            /* 1 | 2 (slow_reservoir [+ percolation -> slow_reservoir_2]) */
    int32_t quick_kind;               /* HBK_QUICK_* */
    int32_t slow_process_order;       /* HBK_SLOW_ORDER_* */
    int32_t splitter_kind;            /* HBK_SPLIT_* */
    int32_t has_glacier;              /* a structure variant with a glacier land cover exists */
    int32_t glacier_infinite_storage; /* Glacier.cpp:21-33 */
    int32_t glacier_no_melt_when_snow_cover;
    double time_step_days;            /* TimeMachine time step in days */
    /* continuous snow -> ice transformation on the glacier snowpack (SettingsModel::AddSnowIceTransformation,
     * SettingsModel.cpp:547-559): HBK_SNOW_ICE_* ; north_hemisphere and the date of the first step feed the
     * day-of-year term of transform:snow_ice_swat (ProcessTransformSnowToIceSwat.cpp:31-39). */
    int32_t snow_ice_kind;
    int32_t north_hemisphere;
    double start_mjd;                 /* modified Julian date of time step 0 (TimeMachine) */
    /* snow sublimation, last process of every snowpack (SettingsModel::AddSnowpackSublimation,
     * SettingsModel.cpp:586-598): HBK_SUBLIMATION_* */
    int32_t sublimation_kind;
    int32_t melt_kind;                /* HBK_MELT_*, the same process on every snowpack and on the glacier ice */
    int32_t snow_slide_kind;          /* HBK_SNOW_SLIDE_* */
    /* liquid water in the snowpacks (SettingsModel::GenerateSnowpacksWithWaterRetention, SettingsModel.cpp:608-634;
     * AddSnowpackRefreezing :574-590), bits HBK_RETENTION_*: 0 = melt water goes straight to the land cover */
    int32_t snow_retention;
} hbk_structure;

/* Parameter slots of one parameter set: float32 like the reference (Parameter.h:121). */
enum {
    HBK_P_TRANSITION_START = 0, /* snow_rain_transition:transition_start (or :threshold) */
    HBK_P_TRANSITION_END,
    HBK_P_RAIN_CORRECTION,
    HBK_P_SNOW_CORRECTION,
    HBK_P_GROUND_SNOW_DDF, /* <cover>_snowpack:degree_day_factor | :degree_day_factor_n | :melt_factor (HBK_MELT_*) */
    HBK_P_GROUND_SNOW_TMELT,
    HBK_P_GLACIER_SNOW_DDF,
    HBK_P_GLACIER_SNOW_TMELT,
    HBK_P_ICE_DDF, /* glacier:degree_day_factor */
    HBK_P_ICE_TMELT,
    HBK_P_CAPACITY,   /* slow_reservoir:capacity (A) */
    HBK_P_K_SLOW_1,   /* slow_reservoir:response_factor */
    HBK_P_PERCOLATION,/* slow_reservoir:percolation_rate */
    HBK_P_K_SLOW_2,   /* slow_reservoir_2:response_factor */
    HBK_P_QUICK,      /* surface_runoff:beta or :response_factor */
    HBK_P_K_SNOW,     /* glacier_area_rain_snowmelt_storage:response_factor */
    HBK_P_K_ICE,      /* glacier_area_icemelt_storage:response_factor */
    HBK_P_SNOW_ICE,   /* glacier_snowpack:snow_ice_transformation_rate or :snow_ice_transformation_basal_acc_coeff */
    HBK_P_GROUND_SUBLIMATION,  /* <cover>_snowpack:sublimation_rate or :sublimation_pet_factor */
    HBK_P_GLACIER_SUBLIMATION,
    /* second and third parameter of the three melt processes: degree_day_factor_s and degree_day_factor_ew|_we
     * (melt:degree_day_aspect) or radiation_coefficient and nothing (melt:temperature_index); unused by melt:degree_day */
    HBK_P_GROUND_SNOW_MELT_B, HBK_P_GROUND_SNOW_MELT_C,
    HBK_P_GLACIER_SNOW_MELT_B, HBK_P_GLACIER_SNOW_MELT_C,
    HBK_P_ICE_MELT_B, HBK_P_ICE_MELT_C,
    /* transport:snow_slide (ProcessLateralSnowSlide.cpp:24-31), the same values on every snowpack that slides */
    HBK_P_SLIDE_COEFF, HBK_P_SLIDE_EXP, HBK_P_SLIDE_MIN_SLOPE, HBK_P_SLIDE_MAX_SLOPE, HBK_P_SLIDE_MIN_HOLDING_DEPTH,
    HBK_P_SLIDE_MAX_SNOW_DEPTH,
    /* second glacier cover of a unit (hbk_set_unit_twins; glacier_ice + glacier_debris of the reference's Socont,
     * parameters.py:1766-1835): its snowpack's and its ice's melt parameters, snow -> ice transformation, sublimation */
    HBK_P_GLACIER2_SNOW_DDF, HBK_P_GLACIER2_SNOW_TMELT, HBK_P_ICE2_DDF, HBK_P_ICE2_TMELT, HBK_P_SNOW_ICE2,
    HBK_P_GLACIER2_SUBLIMATION, HBK_P_GLACIER2_SNOW_MELT_B, HBK_P_GLACIER2_SNOW_MELT_C, HBK_P_ICE2_MELT_B, HBK_P_ICE2_MELT_C,
    /* <cover>_snowpack:water_holding_capacity and :refreezing_factor (HBK_RETENTION_*) */
    HBK_P_GROUND_SNOW_WHC, HBK_P_GLACIER_SNOW_WHC, HBK_P_GROUND_SNOW_CFR, HBK_P_GLACIER_SNOW_CFR,
    HBK_N_PARAMS
};

/* Recorder slots = hydro-unit log labels (SettingsModel.cpp:808-841, Logger.cpp:93-120). */
enum {
    HBK_R_GROUND_WATER = 0, HBK_R_GROUND_INFILTRATION, HBK_R_GROUND_RUNOFF,
    HBK_R_GLACIER_WATER, HBK_R_GLACIER_ICE, HBK_R_GLACIER_OUTFLOW, HBK_R_GLACIER_MELT,
    HBK_R_GROUND_SNOWPACK_WATER, HBK_R_GROUND_SNOWPACK_SNOW, HBK_R_GROUND_SNOWPACK_MELT,
    HBK_R_GLACIER_SNOWPACK_WATER, HBK_R_GLACIER_SNOWPACK_SNOW, HBK_R_GLACIER_SNOWPACK_MELT,
    HBK_R_SLOW_WATER, HBK_R_SLOW_ET, HBK_R_SLOW_OUTFLOW, HBK_R_SLOW_OVERFLOW, HBK_R_SLOW_PERCOLATION,
    HBK_R_SLOW2_WATER, HBK_R_SLOW2_OUTFLOW,
    HBK_R_QUICK_WATER, HBK_R_QUICK_RUNOFF,
    HBK_R_FRACTION_GROUND, HBK_R_FRACTION_GLACIER, /* land-cover fraction series (Logger::RecordFractions) */
    HBK_R_GLACIER_SNOWPACK_TRANSFO,                /* glacier_snowpack:snow_ice_transfo:output */
    HBK_R_GROUND_SNOWPACK_SUBLIMATION, HBK_R_GLACIER_SNOWPACK_SUBLIMATION, /* <cover>_snowpack:sublimation:output;
                                                      0 while the cover is null (the reference keeps the last amount) */
    /* <cover>_snowpack:meltwater:output and :refreeze:output (HBK_RETENTION_*) */
    HBK_R_GROUND_SNOWPACK_MELTWATER, HBK_R_GLACIER_SNOWPACK_MELTWATER,
    HBK_R_GROUND_SNOWPACK_REFREEZE, HBK_R_GLACIER_SNOWPACK_REFREEZE,
    HBK_N_RECORDS
};

/* Storages that can be given an initial content / read back (WaterContainer::SetInitialState). */
enum {
    HBK_S_GROUND_SNOW = 0, HBK_S_GLACIER_SNOW, HBK_S_GROUND_WATER, HBK_S_GLACIER_WATER, HBK_S_ICE,
    HBK_S_SLOW, HBK_S_SLOW2, HBK_S_QUICK,
    HBK_S_AMT_INFILTRATION, HBK_S_AMT_REST, HBK_S_AMT_MELT_GROUND, HBK_S_AMT_MELT_GLACIER,
    HBK_S_AMT_DEPOSIT_RAIN_SNOWMELT, HBK_S_AMT_DEPOSIT_ICEMELT, HBK_S_AMT_SNOW_ICE,
    /* HBK_RETENTION_*: the refreeze fluxes' last amounts (they count as static inputs of the snow containers until the
     * next refreeze, WaterContainer.cpp:98-101) and the melt amounts (with retention HBK_S_AMT_MELT_* are the melt WATER
     * released to the covers); then two more storages, the snowpacks' liquid water */
    HBK_S_AMT_REFREEZE_GROUND, HBK_S_AMT_REFREEZE_GLACIER, HBK_S_AMT_SNOWMELT_GROUND, HBK_S_AMT_SNOWMELT_GLACIER,
    HBK_S_GROUND_SNOWPACK_WATER, HBK_S_GLACIER_SNOWPACK_WATER,
    HBK_N_STATES
};

/* ModelHydro::InitializeWithBasin (ModelHydro.cpp:20-74): structure + time axis. */
int hbk_engine_create(const hbk_structure* structure, int32_t n_units, int32_t n_steps, int32_t device,
                      hbk_engine** out);
void hbk_engine_destroy(hbk_engine* e);
const char* hbk_last_error(const hbk_engine* e);

/* SubBasin::BuildBasin + AssignFractions (SubBasin.cpp:28-101), units in basin order.
 * slope: float(tan(deg*pi/180)) as HydroUnit::GetPropertyFloat returns it (HydroUnit.cpp:146-148);
 * may be NULL when quick_kind is linear. glacier_variant[u] != 0: the unit was assigned the structure
 * variant that carries the glacier bricks (ModelBuilder.cpp:36-97). fractions: land-cover area fractions. */
int hbk_set_units(hbk_engine* e, const double* area_m2, const double* elevation_m, const float* slope_m_per_m,
                  const int32_t* glacier_variant, const double* fraction_ground, const double* fraction_glacier);

/* Aspect class of every unit, HBK_ASPECT_* (HydroUnit property "aspect_class", read by
 * ProcessMeltDegreeDayAspect::SetHydroUnitProperties, ProcessMeltDegreeDayAspect.cpp:35-37); required before hbk_run
 * when melt_kind is HBK_MELT_DEGREE_DAY_ASPECT. */
int hbk_set_unit_aspect(hbk_engine* e, const int32_t* aspect_class);

/* Several glacier covers per hydro unit (models/socont.py:131-140: every land cover of type glacier gets its own brick,
 * snowpack, melt parameters and area fraction; all of them deposit into the same two sub-basin stores). The columns of a
 * unit's land covers are independent of one another, so the engine carries the SECOND glacier cover of a unit as a twin
 * unit: an extra entry of hbk_set_units with the same area, elevation, slope and forcing, ground fraction 0 (a null brick:
 * its soil and surface stores stay empty) and the second cover's fraction as its glacier fraction, whose glacier
 * parameters come from the HBK_P_GLACIER2_* / HBK_P_ICE2_* slots. twin_of[n_units]: index of the unit a twin belongs to,
 * -1 for an ordinary unit. Twins do not count towards the basin area (outlet weights area / basinArea,
 * ModelBuilder.cpp:468). Call before hbk_set_units. Dated actions and lateral connections are refused with twins. */
int hbk_set_unit_twins(hbk_engine* e, const int32_t* twin_of);

/* SettingsBasin::AddLateralConnection (hydro_units.py set_connectivity; ModelBuilder.cpp:476-508): n connections
 * giver -> receiver (unit indices in basin order) with the share of the giver's sliding snow that the receiver gets, in
 * the order they were added (= the order of the process's outputs); slope_degrees[n_units]: the units' "slope" property in
 * degrees as ProcessLateralSnowSlide::SetHydroUnitProperties reads it (float). Required when snow_slide_kind is set. The
 * units of a parameter set are then integrated one after the other inside a time step, in basin order, as the reference
 * does (a giver processed before its receiver feeds it in the same step); every receiver must come AFTER its giver in
 * basin order (downhill in a basin sorted by elevation, HBK_E_UNSUPPORTED otherwise). */
int hbk_set_lateral_connections(hbk_engine* e, int32_t n, const int32_t* giver, const int32_t* receiver,
                                const double* fraction, const float* slope_degrees);

/* SettingsModel::SetParameterValue + ModelHydro::UpdateParameters (ModelHydro.cpp:76-84), batched:
 * values[n_sets][HBK_N_PARAMS] float32. */
int hbk_set_parameters(hbk_engine* e, int32_t n_sets, const float* values);

/* ModelHydro::CreateTimeSeries + AttachTimeSeriesToHydroUnits (ModelHydro.cpp:306-346):
 * data[n_steps][n_units] float64, columns in basin order, shared by all parameter sets. variable = HBK_VAR_*;
 * solar radiation is required iff melt_kind is HBK_MELT_TEMPERATURE_INDEX. */
int hbk_set_forcing(hbk_engine* e, int32_t variable, const double* data);

/* Fused forcing spatialisation (python forcing.py:1061-1113): station series[n_steps]; per unit
 *   temperature: x + g*(z - z_ref)/100;  precipitation: max(0, (x*corr) * (1 + g*(z - z_ref)/100));
 *   pet, solar radiation: x.   gradient / correction may be per parameter set ([n_sets], NULL -> scalar). */
int hbk_set_station_forcing(hbk_engine* e, int32_t variable, const double* series, double ref_elevation,
                            double gradient, const double* gradient_per_set, double correction,
                            const double* correction_per_set);

/* WaterContainer::SetInitialState for every (set, unit): values[n_units] broadcast over sets, NULL = 0. */
int hbk_set_initial_state(hbk_engine* e, int32_t state_slot, const double* values);
/* ModelHydro::SaveAsInitialState (ModelHydro.cpp:195-197): end-of-run contents become the initial state. */
int hbk_save_as_initial_state(hbk_engine* e);
/* SettingsModel::SetSpinupDays (ModelHydro.cpp:48-53, 135-181). */
int hbk_set_spinup_steps(hbk_engine* e, int32_t n_steps);
/* Selective recorder: bit i of mask = HBK_R_* slot i (Logger.cpp:93-120). Memory = popcount*P*T*U*8 B. */
int hbk_set_record_mask(hbk_engine* e, uint64_t mask);

/* Dated actions = kernel boundaries (TimeMachine::IncrementTime -> ActionsManager::DateUpdate,
 * TimeMachine.cpp:49-59, ActionsManager.cpp:80-104). An event fires between time steps, BEFORE step `step`
 * (1 <= step < n_steps); events of the same step fire in the order they were added (the host adds recursive
 * actions in registration order, then sporadic items in date order). Not applied during spin-up.
 *  - land-cover change: ActionLandCoverChange::Apply -> HydroUnit::ChangeLandCoverAreaFraction
 *    (ActionLandCoverChange.cpp:25-39, HydroUnit.cpp:332-490): new glacier fraction of one unit, the generic
 *    cover absorbs the difference, the water/snow column of the strip changes hands;
 *  - snow to ice: ActionGlacierSnowToIceTransformation::Apply (…SnowToIceTransformation.cpp:41-74);
 *  - delta-h: ActionGlacierEvolutionDeltaH::Apply (…DeltaH.cpp:79-142) with the lookup tables
 *    areas/volumes[n_rows][n_units] (row-major) of hbk_set_delta_h_tables, which also performs Init
 *    (…DeltaH.cpp:21-70): checks the initial glacier areas and sets the initial ice content. */
int hbk_clear_actions(hbk_engine* e);
/* unit = -1 (unit shards only): the edit concerns a unit of another block; this block just cuts its run at `step`. */
int hbk_add_land_cover_change(hbk_engine* e, int32_t step, int32_t unit, double glacier_fraction);
int hbk_add_snow_to_ice(hbk_engine* e, int32_t step);
/* ... of the units' second glacier cover (glacier_cover = 1, twin units) or the first (0 = hbk_add_snow_to_ice). With twin
 * units, hbk_add_land_cover_change / hbk_set_delta_h_tables take ENGINE unit indices: an ordinary unit stands for its first
 * glacier cover, its twin for the second; the generic cover that absorbs an area change is the unit's
 * (HydroUnit::ChangeLandCoverAreaFraction, HydroUnit.cpp:332-490, with all three covers in FixLandCoverFractionsTotal). */
int hbk_add_snow_to_ice_on(hbk_engine* e, int32_t step, int32_t glacier_cover);
int hbk_set_delta_h_tables(hbk_engine* e, int32_t n_units, const int32_t* units, int32_t n_rows, const double* areas,
                           const double* volumes);
int hbk_add_delta_h_update(hbk_engine* e, int32_t step);
/*  - area scaling: ActionGlacierEvolutionAreaScaling::Apply (…AreaScaling.cpp:75-123) on the SAME lookup tables and
 *    Init as delta-h (hbk_set_delta_h_tables): every unit of the table follows its own column — retreat of its own
 *    ice volume -> row -> new glacier area, ice w.e. rescaled to the new area. */
int hbk_add_area_scaling_update(hbk_engine* e, int32_t step);

/* ModelHydro::Reset + Run (ModelHydro.cpp:100-133, 183-193) for all parameter sets. */
int hbk_run(hbk_engine* e);

/* Unit sharding (large distributed basins): this engine holds a BLOCK of the basin's hydro units,
 * the other blocks live in other engines (other GPUs). basin_area_m2 = area of the whole basin summed in unit order
 * as SubBasin::AddHydroUnit does (SubBasin.cpp:173), so that the outlet weights area/basinArea (ModelBuilder.cpp:468)
 * are those of the full model; <= 0 turns sharding off. Call after hbk_set_units. In shard mode hbk_run stops after
 * the unit pass: hbk_shard_partials_dev then gives the device buffers [n_rows][ld] float64 holding this block's
 * area-weighted outlet contribution and its per-step deposits into the two sub-basin stores
 * (glacier_area_rain_snowmelt_storage, glacier_area_icemelt_storage; NULL without glacier) — the caller sums them
 * IN PLACE over all shards (one ncclAllReduce each; rows: outlet in the caller's set order, deposits in the engine's
 * internal set order, identical on all shards given identical parameter sets) and calls hbk_finish_sharded_run,
 * which integrates the two sub-basin stores on the summed deposits (SubBasin bricks, modules/glacier.py:108-131)
 * and adds their outflow: hbk_get_outlet then returns the outlet of the whole basin on every shard.
 * Without a reduction callback (below) spin-up and dated actions are refused in shard mode (HBK_E_UNSUPPORTED). */
int hbk_set_unit_shard(hbk_engine* e, double basin_area_m2);
int hbk_shard_partials_dev(hbk_engine* e, double** outlet, double** deposits_rain_snowmelt, double** deposits_icemelt,
                           int64_t* n_rows, int64_t* ld);
int hbk_finish_sharded_run(hbk_engine* e);
/* Unit shards with spin-up (ModelHydro::RunSpinup, ModelHydro.cpp:135-181) and dated actions: the engine then needs sums
 * over all shards INSIDE hbk_run — the spin-up deposits before the sub-basin stores are spun up, the glacier water
 * equivalent of every parameter set at each delta-h date (ActionGlacierEvolutionDeltaH.cpp:90-104 sums it over all
 * glacier units of the basin), the partial outlet / deposits / stale amounts at the end — and asks the caller for them:
 * reduce(user, buf, n) must replace the n float64 at the DEVICE address buf by their sum over all shards (one
 * ncclAllReduce / MPI_Allreduce on the caller's communicator; every shard calls it the same number of times with the same
 * n, in the same order) and return 0 once the result is in place. The engine's stream is idle during the call. With a
 * callback set, hbk_run completes the run itself (hbk_finish_sharded_run is not called) and hbk_get_outlet returns the
 * outlet of the whole basin on every shard. NULL removes the callback. */
typedef int (*hbk_shard_reduce_fn)(void* user, double* device_buffer, int64_t n);
int hbk_set_shard_reduce(hbk_engine* e, hbk_shard_reduce_fn reduce, void* user);
/* Delta-h lookup tables of a unit shard: hbk_set_delta_h_tables is given the shard's own columns; the quantities
 * ActionGlacierEvolutionDeltaH::Apply takes over the WHOLE table — the volume sum of every row (summed in table order,
 * ...DeltaH.cpp:123-126; row 0 x ice density is the initial water equivalent, :66-69) — come from the caller:
 * row_volume_sums[n_rows]. Call after hbk_set_delta_h_tables. */
int hbk_set_delta_h_shard_totals(hbk_engine* e, int32_t n_rows, const double* row_volume_sums);

/* ModelHydro::GetOutletDischarge (ModelHydro.cpp:203-205): out[n_sets][n_steps]. */
int hbk_get_outlet(hbk_engine* e, int32_t first_set, int32_t n_sets, double* out);
/* The same without blocking the caller: the copy is queued on a stream of its own (one cudaMemcpy2DAsync; `out` should be
 * page-locked) and hbk_wait_outlet returns once it has landed. A following hbk_run computes into a second outlet buffer,
 * so the outlet of run i travels to the host while run i+1 integrates (calibration loops that keep every series). */
int hbk_get_outlet_async(hbk_engine* e, int32_t first_set, int32_t n_sets, double* out);
int hbk_wait_outlet(hbk_engine* e);
/* ModelHydro::GetHydroUnitValues(label) (ModelHydro.cpp:227-236) for one set: out[n_steps][n_units];
 * NaN where the unit's structure variant has no such brick. */
int hbk_get_recorded(hbk_engine* e, int32_t record_slot, int32_t set, double* out);
/* End-of-run storage contents: out[n_sets][n_units]. */
int hbk_get_state(hbk_engine* e, int32_t state_slot, double* out);
/* Sub-basin stores (glacier_area_*_storage) end-of-run contents: out[n_sets][2]. */
int hbk_get_basin_state(hbk_engine* e, double* out);

/* Device-resident access for callers that keep data on the GPU (bench `value`, multi-GPU gather):
 * pointer to outlet[n_sets][ld] float64 and its leading dimension. */
int hbk_outlet_dev(hbk_engine* e, const double** ptr, int64_t* ld);
/* Device time of the last hbk_run (CUDA events on the engine's stream), and kernel launch count. */
int hbk_last_run_stats(hbk_engine* e, double* kernel_ms, int32_t* n_launches, int64_t* unconverged_sweeps);
/* Launch shape of the last hbk_run (what a bench line reports as its block shape), out[8]: parameter sets per block (lanes),
 * hydro units per block, unit tiles, unit-kernel launches per run segment (2+: the tiles without glacier units run the
 * plain kernel variant), time steps per chunk and tiles per block of the first launch, tile groups in total (> 1: partial
 * outlet series + hbkReduceTiles), unit tiles run by the plain variant of a glacier structure. */
int hbk_last_run_shape(hbk_engine* e, int32_t* out);
/* Per parameter set: number of constraint passes of its hydro units in the last run that ended with the reference's
 * "The storage constraints did not stabilize after 10 sweeps" (Processor.cpp:191-209) — the ill-conditioned small-capacity
 * regime in which the reference's own result hangs on the last bit. out[n_sets], caller's set order. */
int hbk_get_unconverged(hbk_engine* e, int32_t* out);
/* Raw CUDA stream handle (cudaStream_t) the engine launches on, for callers timing with events. */
void* hbk_stream(hbk_engine* e);

/* Batched objective functions on the device (python evaluation/metrics.py:40-62 via HydroErr):
 * one score per parameter set of the last run against observed[n_steps] (NaN = missing, dropped), skipping the first
 * `warmup` steps (trainer.py:1091-1141). HBK_SCORE_NSE = 1 - sum((s-o)^2)/sum((o-mean o)^2);
 * HBK_SCORE_KGE_2012 = 1 - sqrt((r-1)^2 + (CVs/CVo-1)^2 + (mean s/mean o-1)^2), std with ddof=1;
 * HBK_SCORE_KGE_2009 the same with the ratio of the standard deviations in place of the CV ratio; the error metrics
 * me = mean(s-o), mae = mean|s-o|, mse, rmse = sqrt(mse), nrmse_mean = rmse/mean(o); r_squared = Pearson r^2
 * (HydroErr's names; published formulas: parity against HydroErr itself is unpinned, the package is absent).
 * out[n_sets] on the host; hbk_scores_dev gives the device copy (for NCCL gathers). */
enum {
    HBK_SCORE_NSE = 0, HBK_SCORE_KGE_2012 = 1, HBK_SCORE_KGE_2009, HBK_SCORE_ME, HBK_SCORE_MAE, HBK_SCORE_MSE, HBK_SCORE_RMSE,
    HBK_SCORE_NRMSE_MEAN, HBK_SCORE_R_SQUARED, HBK_N_SCORES
};
int hbk_compute_scores(hbk_engine* e, int32_t metric, const double* observed, int32_t warmup, double* out);
int hbk_scores_dev(hbk_engine* e, const double** ptr);

/* Measured FP64 (non-tensor) DFMA throughput of the device in TFLOP/s: the roofline denominator of the
 * compute-bound ensemble configurations (MEASURED_PEAKS.json has no FP64 figure). */
int hbk_measure_fp64_peak(int32_t device, double* tflops);

/* Device self-test of the kernels' square root of a filling ratio (csrc/hbk_device.cuh sqrtUnit: CUDA's sqrt sequence
 * without its special-case path) against sqrt() on n pseudo-random values of (2^-41, 1]: number of values whose bits
 * differ (must be 0). Diagnostic entry point, not part of the ModelHydro surface. */
int hbk_selftest_sqrt(int32_t device, int64_t n, int64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* HBK_H */
