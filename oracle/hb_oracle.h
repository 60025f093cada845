/* oracle/hb_oracle.h — C API of the CPU restatement of hydrobricks' ModelHydro::Run() hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under hydrobricks_b200/ may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the
 * checker. See hb_oracle.cpp for the reference file:line each function follows.
 *
 * Parity status: PINNED — checked bit-for-bit against the compiled reference core (oracle/_ref) and
 * against the reference's own known-answer tests (tests/test_oracle_*.py, tests/golden/).
 *
 * The model is a flat graph assembled by the caller (oracle/hb_oracle.py mirrors
 * core/src/base/ModelBuilder.cpp for that): containers, bricks, processes, fluxes, splitters.
 * All ids are indices returned by the hbo_add_* calls.
 */
#ifndef HB_ORACLE_H
#define HB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hbo_model hbo_model;

enum { HBO_WATER = 0, HBO_SNOW = 1, HBO_ICE = 2 };
enum { HBO_BRICK_STORAGE = 0, HBO_BRICK_GENERIC_LAND_COVER = 1, HBO_BRICK_GLACIER = 2, HBO_BRICK_SNOWPACK = 3 };
enum {
    HBO_P_MELT_DEGREE_DAY = 0, /* params: degree_day_factor, melting_temperature; forcing temperature */
    HBO_P_ET_SOCONT = 1,       /* forcing pet */
    HBO_P_INFILTRATION_SOCONT = 2,
    HBO_P_RUNOFF_SOCONT = 3,   /* params: beta */
    HBO_P_OUTFLOW_LINEAR = 4,  /* params: response_factor */
    HBO_P_OUTFLOW_DIRECT = 5,
    HBO_P_OUTFLOW_REST = 6,
    HBO_P_OVERFLOW = 7,
    HBO_P_PERCOLATION_CONSTANT = 8, /* params: percolation_rate */
    HBO_P_TRANSFORM_SNOW_ICE_CONSTANT = 9, /* params: snow_ice_transformation_rate; on the snowpack's snow container */
    HBO_P_TRANSFORM_SNOW_ICE_SWAT = 10, /* params: snow_ice_transformation_basal_acc_coeff, north_hemisphere */
    HBO_P_SUBLIMATION_CONSTANT = 11,    /* params: sublimation_rate; snow container -> atmosphere */
    HBO_P_SUBLIMATION_PET = 12,         /* params: sublimation_pet_factor; forcing pet */
    /* melt:temperature_index (ProcessMeltTemperatureIndex.cpp): params melt_factor, melting_temperature,
     * radiation_coefficient; forcing temperature + solar_radiation. melt:degree_day_aspect
     * (ProcessMeltDegreeDayAspect.cpp) needs no type of its own: the builder hands HBO_P_MELT_DEGREE_DAY the
     * degree-day factor of the unit's aspect class, exactly what SetParameters binds (…Aspect.cpp:38-58). */
    HBO_P_MELT_TEMPERATURE_INDEX = 13,
    /* transport:snow_slide (ProcessLateralSnowSlide.cpp): params coeff, exp, min_slope, max_slope,
     * min_snow_holding_depth, max_snow_depth; one instantaneous snow flux per (lateral connection, receiving snowpack) */
    HBO_P_LATERAL_SNOW_SLIDE = 14,
    /* outflow:snow_holding (ProcessOutflowSnowHolding.cpp): params water_holding_capacity; on the snowpack's water
     * container, drains what exceeds the capacity x snow water equivalent */
    HBO_P_OUTFLOW_SNOW_HOLDING = 15,
    /* refreeze:degree_day (ProcessRefreezeDegreeDay.cpp): params refreezing_factor; on the snowpack's water container,
     * instantaneous snow flux back to the same snowpack; degree-day factor and melting temperature of the sibling
     * melt:degree_day process; forcing temperature */
    HBO_P_REFREEZE_DEGREE_DAY = 16,
    /* HBV-96 (models/hbv.py, hbv_96.py; ORACLE ONLY — the CUDA engine does not run this family):
     * infiltration:hbv (ProcessInfiltrationHBV.cpp) params beta; et:hbv (ProcessETHBV.cpp) params lp, et_correction_factor,
     * forcing pet; capillary:hbv (ProcessCapillaryHBV.cpp) params max_capillary_flux, one output per target soil store,
     * weighted by the area fractions of the land covers that feed it; runoff:hbv (ProcessRunoffHBV.cpp) params
     * response_factor, alpha; routing:hbv (ProcessRoutingHBV.cpp) params maxbas, a triangular unit hydrograph with its
     * delivery schedule as process state */
    HBO_P_INFILTRATION_HBV = 17,
    HBO_P_ET_HBV = 18,
    HBO_P_CAPILLARY_HBV = 19,
    HBO_P_RUNOFF_HBV = 20,
    HBO_P_ROUTING_HBV = 21,
    /* GR4J (models/gr4j.py; ORACLE ONLY): interception:gr4j (publishes the net evaporative demand through the unit's PET
     * forcing, ProcessInterceptionGR4J.cpp), production:gr4j and et:gr4j on the production store (helpers/GR4JFormulas.h),
     * routing:gr4j params exchange_factor, routing_capacity, uh_base_time (two unit hydrographs, the routing store and the
     * groundwater exchange as process state, ProcessRoutingGR4J.cpp) */
    HBO_P_INTERCEPTION_GR4J = 22,
    HBO_P_PRODUCTION_GR4J = 23,
    HBO_P_ET_GR4J = 24,
    HBO_P_ROUTING_GR4J = 25,
    /* GR6J (models/gr6j.py): routing:gr6j params exchange_factor, routing_capacity, uh_base_time, exchange_threshold,
     * exp_store_coeff — GR4J's routing with a 60/40 split into the routing store and a bottomless exponential store
     * (ProcessRoutingGR6J.cpp, helpers/GR6JFormulas.h); its container may go negative */
    HBO_P_ROUTING_GR6J = 26
};
enum {
    HBO_F_TO_BRICK = 0, HBO_F_TO_BRICK_INSTANT = 1, HBO_F_TO_OUTLET = 2, HBO_F_TO_ATMOSPHERE = 3,
    HBO_F_SIMPLE = 4, HBO_F_FORCING = 5
};
enum { HBO_S_SNOW_RAIN_LINEAR = 0, HBO_S_MULTI_FLUXES = 1, HBO_S_RAIN = 2, HBO_S_SNOW_RAIN_THRESHOLD = 3 };
enum { HBO_V_PRECIPITATION = 0, HBO_V_TEMPERATURE = 1, HBO_V_PET = 2, HBO_V_SOLAR_RADIATION = 3 };
enum { HBO_SOLVER_EULER = 0, HBO_SOLVER_HEUN = 1, HBO_SOLVER_RK4 = 2,
       /* sequential family (base/SolverSequential.cpp): Solver::Factory names crank_nicolson|trapezoidal,
        * implicit_euler|euler_implicit, exponential_euler, analytic_linear|analytic (Solver.cpp:42-64) */
       HBO_SOLVER_CRANK_NICOLSON = 3, HBO_SOLVER_IMPLICIT_EULER = 4, HBO_SOLVER_EXPONENTIAL_EULER = 5,
       HBO_SOLVER_ANALYTIC_LINEAR = 6 };
enum { HBO_REC_CONTAINER = 0, HBO_REC_FLUX = 1, HBO_REC_OUTLET = 2, HBO_REC_FRACTION = 3 };

hbo_model* hbo_create(int n_units, int n_steps, double dt_days, int solver, int spinup_steps);
void hbo_destroy(hbo_model* m);

int hbo_add_brick(hbo_model* m, int kind, int unit /* -1 = sub-basin */, int needs_solver, int parent_brick);
void hbo_set_area_fraction(hbo_model* m, int brick, double fraction);
void hbo_set_parent(hbo_model* m, int brick, int parent_brick);
void hbo_set_target_brick(hbo_model* m, int process, int target_brick);
/* capillary:hbv: one more target soil store of the process; weight_bricks = the land covers feeding that store */
void hbo_add_capillary_target(hbo_model* m, int process, int target_brick, int n_weights, const int* weight_bricks);
/* capacity: NaN = no capacity. related_snow_container: -1 = none. */
int hbo_add_container(hbo_model* m, int brick, int content_kind, float capacity, int infinite_storage,
                      int no_melt_when_snow_cover, double initial_state);
void hbo_set_related_snowpack(hbo_model* m, int ice_container, int snowpack_brick);
int hbo_add_process(hbo_model* m, int brick, int type, int container, const float* params, int n_params,
                    int target_brick, double unit_area, float slope);
int hbo_add_flux(hbo_model* m, int kind, int is_static, int needs_weighting, double fraction_unit_area,
                 int target_container /* -1 */, int forcing_var /* -1 */, int unit /* -1 */);
void hbo_process_add_output(hbo_model* m, int process, int flux);
void hbo_container_add_input(hbo_model* m, int container, int flux);
int hbo_add_splitter(hbo_model* m, int unit, int type, const float* params, int n_params);
void hbo_splitter_add_output(hbo_model* m, int splitter, int flux);
void hbo_splitter_add_input(hbo_model* m, int splitter, int flux);
void hbo_add_outlet_flux(hbo_model* m, int flux);
/* data[T][U], row-major */
void hbo_set_forcing(hbo_model* m, int var, const double* data);
void hbo_set_lateral_slope(hbo_model* m, int process, float slope_deg);
void hbo_add_lateral_output(hbo_model* m, int process, int target_snowpack_brick, double weight);
void hbo_set_unit_area(hbo_model* m, int unit, double area);
int hbo_add_record(hbo_model* m, int kind, int id);
/* Date of the first time step (modified Julian date), for the processes that read the day of the year
 * (TimeMachine::GetCurrentDayOfYear, TimeMachine.cpp:87-98). Default 44605 = 1981-01-01. */
void hbo_set_start_mjd(hbo_model* m, double mjd);

/* Dated actions (ActionsManager::DateUpdate, ActionsManager.cpp:80-104), applied between time steps, before
 * step `step` (>= 1), in the order added for equal steps. Not applied during spin-up (ModelHydro.cpp:141-147). */
enum { HBO_ACT_LAND_COVER_CHANGE = 0, HBO_ACT_SNOW_TO_ICE = 1, HBO_ACT_DELTA_H = 2, HBO_ACT_AREA_SCALING = 3 };
/* land-cover change: `brick` = the changed land-cover brick, value = new area fraction (already checked/snapped).
 * snow-to-ice: brick = -1 (all glacier units registered with hbo_add_snow_to_ice_unit). delta-h: brick = -1. */
void hbo_add_event(hbo_model* m, int step, int kind, int brick, double value);
void hbo_add_snow_to_ice_unit(hbo_model* m, int glacier_brick, int snowpack_brick);
/* delta-h lookup tables (ActionGlacierEvolutionDeltaH::AddLookupTables): glacier bricks of the table's units,
 * unit areas, areas/volumes [n_rows][n_units] row-major. */
void hbo_set_delta_h(hbo_model* m, int n_units, const int* glacier_bricks, const double* unit_areas, int n_rows,
                     const double* areas, const double* volumes);
/* generic (soil) land cover and snowpack bricks of a land cover, for the content transfer of a fraction change */
void hbo_set_cover_links(hbo_model* m, int cover_brick, int generic_cover_brick, int snowpack_brick);

/* Re-applies the land-cover fractions to the weighted fluxes (LandCover::SetAreaFraction). */
void hbo_finalize_graph(hbo_model* m);

/* Runs the model; outlet[T]; records[n_records][T]. Returns 0, or <0 with hbo_last_error(). */
int hbo_run(hbo_model* m, double* outlet, double* records);
const char* hbo_last_error(hbo_model* m);
/* Number of constraint sweeps that hit the 10-sweep limit (Processor.cpp:208 warning). */
long hbo_unconverged_sweeps(hbo_model* m);

#ifdef __cplusplus
}
#endif
#endif
