"""ctypes binding of the C-ABI in include/hbk.h (libhbk.so, CUDA sm_100a).

This is product code: it stays independent of the test-only checker directory and raises when the CUDA library or a
CUDA device is missing — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
import os  # noqa: E402

LIB_PATH = Path(os.environ.get("HBK_LIB", HERE / "libhbk.so"))  # HBK_LIB: tuning builds of the same sources

# enums of include/hbk.h
SOLVER = {"euler_explicit": 0, "heun_explicit": 1, "rk4": 2, "runge_kutta": 2,
          # sequential family (base/SolverSequential.cpp); crank_nicolson is the reference Python layer's default
          "crank_nicolson": 3, "trapezoidal": 3, "implicit_euler": 4, "euler_implicit": 4, "exponential_euler": 5,
          "analytic_linear": 6, "analytic": 6}
QUICK_SOCONT, QUICK_LINEAR = 0, 1
SPLIT_LINEAR, SPLIT_THRESHOLD = 0, 1
SNOW_ICE_NONE, SNOW_ICE_CONSTANT, SNOW_ICE_SWAT = 0, 1, 2
SUBLIMATION_NONE, SUBLIMATION_CONSTANT, SUBLIMATION_PET = 0, 1, 2
SNOW_SLIDE_NONE, SNOW_SLIDE_SKIP_GLACIERS, SNOW_SLIDE_ALL = 0, 1, 2
RETENTION_HOLDING, RETENTION_REFREEZE, RETENTION_RAIN_TO_SNOWPACK = 1, 2, 4
# HBK_SCORE_*: HydroErr's names of the batched objective functions of hbk_compute_scores
SCORES = ["nse", "kge_2012", "kge_2009", "me", "mae", "mse", "rmse", "nrmse_mean", "r_squared"]
MELT_DEGREE_DAY, MELT_DEGREE_DAY_ASPECT, MELT_TEMPERATURE_INDEX = 0, 1, 2
MELT_KIND = {"melt:degree_day": 0, "melt:degree_day_aspect": 1, "melt:temperature_index": 2}
# HydroUnit property "aspect_class" (ProcessMeltDegreeDayAspect.cpp:38-58) -> HBK_ASPECT_*
ASPECT = {"north": 0, "N": 0, "south": 1, "S": 1, "east": 2, "west": 2, "E": 2, "W": 2}
VAR = {"precipitation": 0, "p": 0, "temperature": 1, "t": 1, "pet": 2,
       "solar_radiation": 3, "r_solar": 3, "r": 3}  # TimeSeries.cpp:134-146
PARAM_SLOTS = [
    "transition_start", "transition_end", "rain_correction", "snow_correction", "ground_snow_ddf",
    "ground_snow_tmelt", "glacier_snow_ddf", "glacier_snow_tmelt", "ice_ddf", "ice_tmelt", "capacity", "k_slow_1",
    "percolation", "k_slow_2", "quick", "k_snow", "k_ice", "snow_ice",
    "ground_sublimation", "glacier_sublimation",
    "ground_snow_melt_b", "ground_snow_melt_c", "glacier_snow_melt_b", "glacier_snow_melt_c", "ice_melt_b", "ice_melt_c",
    "slide_coeff", "slide_exp", "slide_min_slope", "slide_max_slope", "slide_min_holding_depth", "slide_max_snow_depth",
    "glacier2_snow_ddf", "glacier2_snow_tmelt", "ice2_ddf", "ice2_tmelt", "snow_ice2", "glacier2_sublimation",
    "glacier2_snow_melt_b", "glacier2_snow_melt_c", "ice2_melt_b", "ice2_melt_c",
    "ground_snow_whc", "glacier_snow_whc", "ground_snow_cfr", "glacier_snow_cfr",
]
N_PARAMS = len(PARAM_SLOTS)
RECORD_SLOTS = [
    "ground_water", "ground_infiltration", "ground_runoff", "glacier_water", "glacier_ice", "glacier_outflow",
    "glacier_melt", "ground_snowpack_water", "ground_snowpack_snow", "ground_snowpack_melt",
    "glacier_snowpack_water", "glacier_snowpack_snow", "glacier_snowpack_melt", "slow_water", "slow_et",
    "slow_outflow", "slow_overflow", "slow_percolation", "slow2_water", "slow2_outflow", "quick_water",
    "quick_runoff", "fraction_ground", "fraction_glacier", "glacier_snowpack_transfo",
    "ground_snowpack_sublimation", "glacier_snowpack_sublimation",
    "ground_snowpack_meltwater", "glacier_snowpack_meltwater", "ground_snowpack_refreeze", "glacier_snowpack_refreeze",
]
STATE_SLOTS = [
    "ground_snow", "glacier_snow", "ground_water", "glacier_water", "ice", "slow", "slow2", "quick",
    "amt_infiltration", "amt_rest", "amt_melt_ground", "amt_melt_glacier", "amt_deposit_rain_snowmelt",
    "amt_deposit_icemelt", "amt_snow_ice",
    "amt_refreeze_ground", "amt_refreeze_glacier", "amt_snowmelt_ground", "amt_snowmelt_glacier",
    "ground_snowpack_water", "glacier_snowpack_water",
]


class HbkStructure(ctypes.Structure):
    _fields_ = [
        ("solver", ctypes.c_int32),
        ("n_soil_stores", ctypes.c_int32),
        ("quick_kind", ctypes.c_int32),
        ("slow_process_order", ctypes.c_int32),
        ("splitter_kind", ctypes.c_int32),
        ("has_glacier", ctypes.c_int32),
        ("glacier_infinite_storage", ctypes.c_int32),
        ("glacier_no_melt_when_snow_cover", ctypes.c_int32),
        ("time_step_days", ctypes.c_double),
        ("snow_ice_kind", ctypes.c_int32),
        ("north_hemisphere", ctypes.c_int32),
        ("start_mjd", ctypes.c_double),
        ("sublimation_kind", ctypes.c_int32),
        ("melt_kind", ctypes.c_int32),
        ("snow_slide_kind", ctypes.c_int32),
        ("snow_retention", ctypes.c_int32),
    ]


# hbk_shard_reduce_fn: int (*)(void* user, double* device_buffer, int64_t n)
SHARD_REDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)


class HbkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"hbk error {code}: {message}")
        self.code = code


_lib = None


def load_library():
    """Load libhbk.so; fail loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                          "(nvcc, sm_100a). hydrobricks_b200 has no CPU fallback.")
    L = ctypes.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    dp, fp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    L.hbk_engine_create.argtypes = [ctypes.POINTER(HbkStructure), i32, i32, i32, ctypes.POINTER(vp)]
    L.hbk_engine_destroy.argtypes = [vp]
    L.hbk_engine_destroy.restype = None
    L.hbk_last_error.argtypes = [vp]
    L.hbk_last_error.restype = ctypes.c_char_p
    L.hbk_set_units.argtypes = [vp, dp, dp, fp, ip, dp, dp]
    L.hbk_set_unit_aspect.argtypes = [vp, ip]
    L.hbk_set_unit_twins.argtypes = [vp, ip]
    L.hbk_set_lateral_connections.argtypes = [vp, i32, ip, ip, dp, fp]
    L.hbk_set_parameters.argtypes = [vp, i32, fp]
    L.hbk_set_forcing.argtypes = [vp, i32, dp]
    L.hbk_set_station_forcing.argtypes = [vp, i32, dp, dbl, dbl, dp, dbl, dp]
    L.hbk_set_initial_state.argtypes = [vp, i32, dp]
    L.hbk_save_as_initial_state.argtypes = [vp]
    L.hbk_set_spinup_steps.argtypes = [vp, i32]
    L.hbk_set_record_mask.argtypes = [vp, ctypes.c_uint64]
    L.hbk_clear_actions.argtypes = [vp]
    L.hbk_add_land_cover_change.argtypes = [vp, i32, i32, dbl]
    L.hbk_add_snow_to_ice.argtypes = [vp, i32]
    L.hbk_add_snow_to_ice_on.argtypes = [vp, i32, i32]
    L.hbk_set_delta_h_tables.argtypes = [vp, i32, ip, i32, dp, dp]
    L.hbk_add_delta_h_update.argtypes = [vp, i32]
    L.hbk_add_area_scaling_update.argtypes = [vp, i32]
    L.hbk_run.argtypes = [vp]
    L.hbk_get_outlet.argtypes = [vp, i32, i32, dp]
    L.hbk_get_outlet_async.argtypes = [vp, i32, i32, dp]
    L.hbk_wait_outlet.argtypes = [vp]
    L.hbk_get_recorded.argtypes = [vp, i32, i32, dp]
    L.hbk_get_state.argtypes = [vp, i32, dp]
    L.hbk_get_basin_state.argtypes = [vp, dp]
    L.hbk_outlet_dev.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(i64)]
    L.hbk_last_run_stats.argtypes = [vp, dp, ip, ctypes.POINTER(i64)]
    L.hbk_last_run_shape.argtypes = [vp, ip]
    L.hbk_get_unconverged.argtypes = [vp, ip]
    L.hbk_stream.argtypes = [vp]
    L.hbk_stream.restype = ctypes.c_void_p
    L.hbk_measure_fp64_peak.argtypes = [i32, dp]
    L.hbk_selftest_sqrt.argtypes = [i32, i64, ctypes.POINTER(i64)]
    L.hbk_compute_scores.argtypes = [vp, i32, dp, i32, dp]
    L.hbk_scores_dev.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p)]
    L.hbk_set_unit_shard.argtypes = [vp, dbl]
    L.hbk_shard_partials_dev.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                         ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.hbk_finish_sharded_run.argtypes = [vp]
    L.hbk_set_shard_reduce.argtypes = [vp, SHARD_REDUCE_FN, vp]
    L.hbk_set_delta_h_shard_totals.argtypes = [vp, i32, dp]
    _lib = L
    return L


def selftest_sqrt(n=1 << 26, device=0):
    """Number of values on which the kernels' square root of a filling ratio differs from sqrt() (must be 0)."""
    out = ctypes.c_int64()
    rc = load_library().hbk_selftest_sqrt(int(device), int(n), ctypes.byref(out))
    if rc != 0:
        raise HbkError(rc, "hbk_selftest_sqrt failed")
    return out.value


def measure_fp64_peak(device=0):
    """Measured DFMA throughput [TFLOP/s] (roofline denominator for the compute-bound configurations)."""
    out = ctypes.c_double()
    rc = load_library().hbk_measure_fp64_peak(int(device), ctypes.byref(out))
    if rc != 0:
        raise HbkError(rc, "hbk_measure_fp64_peak failed")
    return out.value


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


class Engine:
    """One engine = one structure + basin + time axis on one CUDA device."""

    def __init__(self, structure: HbkStructure, n_units: int, n_steps: int, device: int = 0):
        self.L = load_library()
        self.h = ctypes.c_void_p()
        self.n_units, self.n_steps, self.n_sets = int(n_units), int(n_steps), 0
        rc = self.L.hbk_engine_create(ctypes.byref(structure), self.n_units, self.n_steps, int(device),
                                      ctypes.byref(self.h))
        if rc != 0:
            raise HbkError(rc, "hbk_engine_create failed (is a CUDA device present? there is no CPU fallback)")
        self.structure = structure

    @classmethod
    def from_handle(cls, handle, n_units, n_steps, n_sets):
        """Non-owning view of an engine created elsewhere (the compiled host module's ModelHydro.engine_handle())."""
        self = cls.__new__(cls)
        self.L = load_library()
        self.h = ctypes.c_void_p(handle)
        self.n_units, self.n_steps, self.n_sets = int(n_units), int(n_steps), int(n_sets)
        self._borrowed = True
        return self

    def close(self):
        if getattr(self, "h", None) and not getattr(self, "_borrowed", False):
            self.L.hbk_engine_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HbkError(rc, self.L.hbk_last_error(self.h).decode())

    def set_units(self, area, elevation=None, slope=None, glacier_variant=None, fraction_ground=None,
                  fraction_glacier=None):
        f64 = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)  # noqa: E731
        area, elevation, fg, fgl = f64(area), f64(elevation), f64(fraction_ground), f64(fraction_glacier)
        slope = None if slope is None else np.ascontiguousarray(slope, dtype=np.float32)
        gv = None if glacier_variant is None else np.ascontiguousarray(glacier_variant, dtype=np.int32)
        assert area.shape == (self.n_units,)
        self._check(self.L.hbk_set_units(
            self.h, _dptr(area), _dptr(elevation),
            slope.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if slope is not None else None,
            gv.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if gv is not None else None, _dptr(fg), _dptr(fgl)))

    def set_unit_twins(self, twin_of):
        """twin_of[n_units]: index of the unit whose second glacier cover this entry carries, -1 for an ordinary unit
        (hbk_set_unit_twins); before set_units."""
        t = np.ascontiguousarray(twin_of, dtype=np.int32)
        assert t.shape == (self.n_units,)
        self._check(self.L.hbk_set_unit_twins(self.h, t.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))

    def set_unit_aspect(self, aspect_classes):
        """Aspect class per unit: HBK_ASPECT_* integers or the reference's strings (north|N|south|S|east|west|E|W)."""
        try:
            a = [ASPECT[x] if isinstance(x, str) else int(x) for x in aspect_classes]
        except KeyError as err:
            raise ValueError(f"Invalid aspect: {err.args[0]}") from None
        a = np.ascontiguousarray(a, dtype=np.int32)
        assert a.shape == (self.n_units,)
        self._check(self.L.hbk_set_unit_aspect(self.h, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))

    def set_lateral_connections(self, givers, receivers, fractions, slope_degrees):
        """SettingsBasin::AddLateralConnection, in the order added: unit positions (basin order) and the share of the
        giver's sliding snow; slope_degrees[n_units] float32 (transport:snow_slide)."""
        i32p = ctypes.POINTER(ctypes.c_int32)
        g = np.ascontiguousarray(givers, dtype=np.int32)
        r = np.ascontiguousarray(receivers, dtype=np.int32)
        f = np.ascontiguousarray(fractions, dtype=np.float64)
        s = np.ascontiguousarray(slope_degrees, dtype=np.float32)
        assert g.shape == r.shape == f.shape and s.shape == (self.n_units,)
        self._check(self.L.hbk_set_lateral_connections(self.h, len(g), g.ctypes.data_as(i32p), r.ctypes.data_as(i32p),
                                                       _dptr(f), s.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))

    def set_parameters(self, values):
        v = np.ascontiguousarray(values, dtype=np.float32)
        if v.ndim == 1:
            v = v[None, :]
        assert v.shape[1] == N_PARAMS, v.shape
        self.n_sets = v.shape[0]
        self._check(self.L.hbk_set_parameters(self.h, self.n_sets, v.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))

    def _per_engine_unit(self, a):
        """[... x basin units] -> [... x engine units] when the engine carries twin units (structure.UnitLayout)."""
        layout = getattr(self, "layout", None)
        if layout is not None and layout.has_twins and a.shape[-1] == layout.n_basin:
            return np.ascontiguousarray(layout.expand(a))
        return a

    def set_forcing(self, variable, data):
        a = self._per_engine_unit(np.ascontiguousarray(data, dtype=np.float64))
        assert a.shape == (self.n_steps, self.n_units), a.shape
        self._check(self.L.hbk_set_forcing(self.h, VAR[variable] if isinstance(variable, str) else variable, _dptr(a)))

    def set_station_forcing(self, variable, series, ref_elevation=0.0, gradient=0.0, correction=1.0):
        s = np.ascontiguousarray(series, dtype=np.float64)
        assert s.shape == (self.n_steps,)
        g_set = c_set = None
        if np.ndim(gradient) > 0:
            g_set = np.ascontiguousarray(gradient, dtype=np.float64)
            assert g_set.shape == (self.n_sets,)
            gradient = 0.0
        if np.ndim(correction) > 0:
            c_set = np.ascontiguousarray(correction, dtype=np.float64)
            assert c_set.shape == (self.n_sets,)
            correction = 1.0
        self._check(self.L.hbk_set_station_forcing(self.h, VAR[variable] if isinstance(variable, str) else variable,
                                                   _dptr(s), float(ref_elevation), float(gradient), _dptr(g_set),
                                                   float(correction), _dptr(c_set)))

    def set_initial_state(self, slot, values=None):
        v = None if values is None else self._per_engine_unit(np.ascontiguousarray(values, dtype=np.float64))
        self._check(self.L.hbk_set_initial_state(self.h, STATE_SLOTS.index(slot) if isinstance(slot, str) else slot,
                                                 _dptr(v)))

    def save_as_initial_state(self):
        self._check(self.L.hbk_save_as_initial_state(self.h))

    def set_spinup_steps(self, n):
        self._check(self.L.hbk_set_spinup_steps(self.h, int(n)))

    def set_record_mask(self, slots):
        mask = 0
        for s in slots:
            mask |= 1 << (RECORD_SLOTS.index(s) if isinstance(s, str) else int(s))
        self._check(self.L.hbk_set_record_mask(self.h, mask))

    # dated actions: events fire before time step `step` (>= 1), same-step events in the order added
    def clear_actions(self):
        self._check(self.L.hbk_clear_actions(self.h))

    def add_land_cover_change(self, step, unit, glacier_fraction):
        self._check(self.L.hbk_add_land_cover_change(self.h, int(step), int(unit), float(glacier_fraction)))

    def add_snow_to_ice(self, step, glacier_cover=0):
        """glacier_cover 1 = the units' second glacier cover (twin units)."""
        self._check(self.L.hbk_add_snow_to_ice_on(self.h, int(step), int(glacier_cover)))

    def set_delta_h_tables(self, units, areas, volumes):
        u = np.ascontiguousarray(units, dtype=np.int32)
        a = np.ascontiguousarray(areas, dtype=np.float64)
        v = np.ascontiguousarray(volumes, dtype=np.float64)
        assert a.shape == v.shape == (a.shape[0], len(u))
        self._check(self.L.hbk_set_delta_h_tables(self.h, len(u), u.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                  a.shape[0], _dptr(a), _dptr(v)))

    def add_delta_h_update(self, step):
        self._check(self.L.hbk_add_delta_h_update(self.h, int(step)))

    def add_area_scaling_update(self, step):
        self._check(self.L.hbk_add_area_scaling_update(self.h, int(step)))

    def run(self):
        self._check(self.L.hbk_run(self.h))

    def get_outlet(self, first=0, n=None):
        n = self.n_sets - first if n is None else n
        out = np.empty((n, self.n_steps), dtype=np.float64)
        self._check(self.L.hbk_get_outlet(self.h, int(first), int(n), _dptr(out)))
        return out

    def get_outlet_async(self, out, first=0, n=None):
        """Queue the copy of the outlet rows into `out` (C-contiguous float64 [n][n_steps], ideally page-locked) on the
        engine's copy stream; `wait_outlet()` returns when it has landed. The next run() may start right away."""
        n = self.n_sets - first if n is None else n
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (n, self.n_steps)
        self._check(self.L.hbk_get_outlet_async(self.h, int(first), int(n), _dptr(out)))

    def wait_outlet(self):
        self._check(self.L.hbk_wait_outlet(self.h))

    def get_recorded(self, slot, set_index=0):
        out = np.empty((self.n_steps, self.n_units), dtype=np.float64)
        self._check(self.L.hbk_get_recorded(self.h, RECORD_SLOTS.index(slot) if isinstance(slot, str) else slot,
                                            int(set_index), _dptr(out)))
        return out

    def get_state(self, slot):
        out = np.empty((self.n_sets, self.n_units), dtype=np.float64)
        self._check(self.L.hbk_get_state(self.h, STATE_SLOTS.index(slot) if isinstance(slot, str) else slot, _dptr(out)))
        return out

    def get_basin_state(self):
        out = np.empty((self.n_sets, 2), dtype=np.float64)
        self._check(self.L.hbk_get_basin_state(self.h, _dptr(out)))
        return out

    def compute_scores(self, metric, observed, warmup=0, out=None, to_host=True):
        """Batched objective function of the last run on the device (SCORES: "nse", "kge_2012", "kge_2009", "me", "mae",
        "mse", "rmse", "nrmse_mean", "r_squared"); returns float64[n_sets]
        (to_host=False leaves them on the device only: scores_device_pointer())."""
        obs = np.ascontiguousarray(observed, dtype=np.float64)
        assert obs.shape == (self.n_steps,)
        kind = SCORES.index(metric.lower()) if isinstance(metric, str) else int(metric)
        if not to_host:
            self._check(self.L.hbk_compute_scores(self.h, kind, _dptr(obs), int(warmup), None))
            return None
        if out is None:
            out = np.empty(self.n_sets, dtype=np.float64)
        self._check(self.L.hbk_compute_scores(self.h, kind, _dptr(obs), int(warmup), _dptr(out)))
        return out

    def scores_device_pointer(self):
        ptr = ctypes.c_void_p()
        self._check(self.L.hbk_scores_dev(self.h, ctypes.byref(ptr)))
        return ptr.value

    # unit sharding (large basins split across GPUs): see include/hbk.h
    def set_unit_shard(self, basin_area_m2):
        self._check(self.L.hbk_set_unit_shard(self.h, float(basin_area_m2)))

    def shard_partials_device_pointers(self):
        """(outlet_ptr, deposits_rain_snowmelt_ptr | None, deposits_icemelt_ptr | None, n_rows, ld) of the last run."""
        q, d1, d2 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n, ld = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.L.hbk_shard_partials_dev(self.h, ctypes.byref(q), ctypes.byref(d1), ctypes.byref(d2),
                                                  ctypes.byref(n), ctypes.byref(ld)))
        return q.value, d1.value, d2.value, n.value, ld.value

    def finish_sharded_run(self):
        self._check(self.L.hbk_finish_sharded_run(self.h))

    def set_shard_reduce(self, reduce):
        """reduce(device_pointer: int, n: int) must sum the n float64 at that device address over all shards in place
        (hbk_set_shard_reduce); with it a unit shard runs spin-up and dated actions and finishes inside run(). None
        removes it."""
        if reduce is None:
            self._reduce_cb = None
            self._check(self.L.hbk_set_shard_reduce(self.h, SHARD_REDUCE_FN(0), None))
            return

        def trampoline(_user, ptr, n):
            try:
                reduce(int(ptr), int(n))
                return 0
            except Exception:  # an exception must not unwind through the C frames
                import traceback

                traceback.print_exc()
                return 1

        self._reduce_cb = SHARD_REDUCE_FN(trampoline)  # kept alive as long as the engine may call it
        self._check(self.L.hbk_set_shard_reduce(self.h, self._reduce_cb, None))

    def set_delta_h_shard_totals(self, row_volume_sums):
        v = np.ascontiguousarray(row_volume_sums, dtype=np.float64)
        self._check(self.L.hbk_set_delta_h_shard_totals(self.h, len(v), _dptr(v)))

    def outlet_device_pointer(self):
        ptr, ld = ctypes.c_void_p(), ctypes.c_int64()
        self._check(self.L.hbk_outlet_dev(self.h, ctypes.byref(ptr), ctypes.byref(ld)))
        return ptr.value, ld.value

    def last_run_stats(self):
        ms, n, unc = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int64()
        self._check(self.L.hbk_last_run_stats(self.h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(unc)))
        return {"kernel_ms": ms.value, "launches": n.value, "unconverged_sweeps": unc.value}

    def last_run_shape(self):
        """Launch shape of the last run (hbk_last_run_shape)."""
        v = (ctypes.c_int32 * 8)()
        self._check(self.L.hbk_last_run_shape(self.h, v))
        keys = ("sets_per_block", "units_per_block", "unit_tiles", "launches_per_segment", "steps_per_chunk",
                "tiles_per_block", "tile_groups", "plain_variant_tiles")
        return dict(zip(keys, (int(x) for x in v)))

    def unconverged_per_set(self):
        """Constraint passes per parameter set that did not stabilise within 10 sweeps (hbk_get_unconverged)."""
        out = np.zeros(self.n_sets, dtype=np.int32)
        self._check(self.L.hbk_get_unconverged(self.h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return out

    def stream(self):
        return self.L.hbk_stream(self.h)
