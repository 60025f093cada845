"""The arithmetic of the CUDA kernels on the CPU: hydrobricks_b200/csrc/hbk_device.cuh — the same source the kernels
are built from — compiled by g++ with -DHBK_HOST_EMULATION (tests/host_emul) and run over every golden case of the
reference (tests/golden, one parameter set, all solvers incl. the sequential family, glacier / snow-ice / sublimation
options), at the GPU parity bar: 1e-10 relative + 1e-12 mm on the outlet discharge and on recorded storages.

What this covers is the per-(set, unit) step functions (directPhase, solveUnit, solveUnitSequential, basinStoreStep,
the constraint sweep, the rate laws); the kernel plumbing around them is only exercised by the `-m gpu` tests. It makes
a change to the device arithmetic checkable without a GPU box."""
import ctypes
import math
import subprocess
from pathlib import Path

import numpy as np
import pytest

import golden_utils as gu

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
EMUL = HERE / "host_emul"
RTOL, ATOL = 1e-10, 1e-12
CASES = [b for b in gu.bundles() if not b.startswith(("c1_", "kat_linear", "actions_", "actions2_", "melt_", "snow_slide_", "two_glaciers_"))]


def build_emul(name="libhbkemul.so", defines=()):
    """Compile hbk_device.cuh + the harness for the host (rebuilt when a source is newer) and bind it."""
    lib = EMUL / name
    srcs = [EMUL / "hbk_host_emul.cpp", EMUL / "hbk_host_shims.h", ROOT / "hydrobricks_b200" / "csrc" / "hbk_device.cuh",
            ROOT / "include" / "hbk.h"]
    if not lib.exists() or lib.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-DHBK_HOST_EMULATION",
                               *defines, f"-I{EMUL}", "-o", str(lib), str(EMUL / "hbk_host_emul.cpp")])
    L = ctypes.CDLL(str(lib))
    dp, fp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    L.hbk_emul_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp, fp, ip, dp, dp, fp, dp, dp, dp, dp, dp, dp,
                               dp, dp, ctypes.POINTER(ctypes.c_longlong)]
    return L


@pytest.fixture(scope="module")
def emul():
    return build_emul()


@pytest.fixture(scope="module")
def emul_faithful():
    """The same arithmetic built with -DHBK_FAITHFUL=1: every division of the reference written as a division (the
    shipped default computes the same correctly rounded quotients from reciprocals, hbk_device.cuh divExact)."""
    return build_emul("libhbkemul_faithful.so", ("-DHBK_FAITHFUL=1",))


@pytest.fixture(scope="module")
def emul_fast():
    """-DHBK_FAST_ARITH=1: the round-1 kernels' arithmetic (bare reciprocal multiplications, verification sweep skipped),
    kept as a measurement variant only."""
    return build_emul("libhbkemul_fast.so", ("-DHBK_FAST_ARITH=1",))


def _day_of_year(mjd):
    d = np.datetime64("1858-11-17") + np.timedelta64(int(math.floor(mjd)), "D")
    return int((d - d.astype("datetime64[Y]")).astype(int)) + 1


def _close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    ok = np.isnan(b) | (np.abs(a - b) <= RTOL * np.abs(b) + ATOL)
    if not ok.all():
        i = int(np.argmax(~ok))
        raise AssertionError(f"{(~ok).sum()} values differ; first at {np.unravel_index(i, a.shape)}: "
                             f"{a.flat[i]!r} vs {b.flat[i]!r}")


def run_emul(L, meta, forcing, solver=None, dt=1.0):
    from hydrobricks_b200 import engine as _e
    from hydrobricks_b200 import structure

    cs = structure.CompiledStructure(meta["structures"], solver or meta["solver"], dt)
    units = meta["units"]
    U, T = len(units), len(next(iter(forcing.values())))
    area = np.array([u["area"] for u in units], dtype=np.float64)
    slope = None
    if cs.hbk.quick_kind == _e.QUICK_SOCONT:
        slope = np.array([structure.slope_m_per_m(*u["slope"]) for u in units], dtype=np.float32)
    covers = [u["land_covers"] for u in units]
    gv = np.array(cs.assign_structure_variants(covers), dtype=np.int32)
    fg = np.array([dict(c).get(cs.ground_name, 1.0) for c in covers], dtype=np.float64)
    fgl = np.array([dict(c).get(cs.glacier_name, 0.0) if cs.glacier_name else 0.0 for c in covers], dtype=np.float64)
    params = np.ascontiguousarray(cs.parameter_vector(), dtype=np.float32)
    f = {("precipitation" if k in ("p", "precipitation") else "temperature" if k in ("t", "temperature") else "pet"):
         np.ascontiguousarray(v, dtype=np.float64) for k, v in forcing.items()}
    swat = None
    if cs.hbk.has_glacier and cs.hbk.snow_ice_kind == _e.SNOW_ICE_SWAT:
        # hbk_engine_create: 1 + sin(2 pi (doy - ref) / 365), ProcessTransformSnowToIceSwat.cpp:31-39
        ref_day = 81 if cs.hbk.north_hemisphere else 264
        pi = 3.1415926535897932384626433832795
        swat = np.array([1 + math.sin(2.0 * pi * (_day_of_year(cs.hbk.start_mjd + t * dt) - ref_day) / 365.0)
                         for t in range(T)], dtype=np.float64)
    dp, fp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    ptr = lambda a, t: None if a is None else a.ctypes.data_as(t)  # noqa: E731
    outlet = np.empty(T)
    rec = {k: np.empty((T, U)) for k in ("slow", "snow", "ice")}
    unconv = ctypes.c_longlong(0)
    err = L.hbk_emul_run(ctypes.addressof(cs.hbk), U, T, ptr(area, dp), ptr(slope, fp), ptr(gv, ip), ptr(fg, dp),
                         ptr(fgl, dp), ptr(params, fp), ptr(f["precipitation"], dp), ptr(f["temperature"], dp),
                         ptr(f["pet"], dp), ptr(swat, dp), ptr(outlet, dp), ptr(rec["slow"], dp), ptr(rec["snow"], dp),
                         ptr(rec["ice"], dp), ctypes.byref(unconv))
    assert err & 1 == 0, "a change rate exceeded 10000"
    return cs, outlet, rec, unconv.value


@pytest.mark.parametrize("name", CASES)
def test_device_arithmetic_on_host_matches_reference(name, emul):
    meta, forcing, outlet, records, _ = gu.load(name)
    cs, q, rec, _ = run_emul(emul, meta, forcing)
    _close(q, outlet)
    _close(rec["slow"], records["slow_reservoir:water_content"])
    snow_label = f"{cs.ground_name}_snowpack:snow_content"
    if snow_label in records:
        _close(rec["snow"], records[snow_label])
    if cs.glacier_name and f"{cs.glacier_name}:ice_content" in records and not cs.hbk.glacier_infinite_storage:
        _close(rec["ice"], records[f"{cs.glacier_name}:ice_content"])


@pytest.mark.parametrize("name", CASES)
def test_faithful_build_matches_reference(name, emul_faithful):
    meta, forcing, outlet, records, _ = gu.load(name)
    _, q, rec, _ = run_emul(emul_faithful, meta, forcing)
    _close(q, outlet)
    _close(rec["slow"], records["slow_reservoir:water_content"])


@pytest.mark.parametrize("solver", ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "implicit_euler",
                                    "exponential_euler", "analytic_linear"])
def test_device_arithmetic_on_host_two_day_step(solver, emul):
    """A time step other than one day, every solver, against the oracle restatement built with the same dt."""
    from oracle import hb_oracle

    meta, forcing, _, _, _ = gu.load("gsm_socont_rk4_glacier_2store")
    n = 500
    forcing = {k: np.ascontiguousarray(v[:n]) for k, v in forcing.items()}
    _, q, rec, _ = run_emul(emul, meta, forcing, solver=solver, dt=2.0)
    om = hb_oracle.OracleModel(meta["structures"], meta["units"], solver, n, dt=2.0)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    om.record("slow_reservoir:water_content")
    _close(q, om.run())
    _close(rec["slow"], om.records["slow_reservoir:water_content"])


MELT_CASES = sorted(p.stem for p in gu.GOLDEN.glob("melt_*.npz"))


@pytest.mark.parametrize("name", MELT_CASES)
def test_melt_variants_on_host(name, emul):
    """melt:degree_day_aspect and melt:temperature_index: the structure compiler's parameter slots (degree_day_factor_n /
    _s / _ew resp. melt_factor / radiation_coefficient) + the kernel glue restated in tests/host_emul — the factor of the
    unit's aspect class picked when the unit is loaded, resp. melt_factor + radiation_coefficient x solar radiation per
    step — around the unchanged device step functions, against the reference's goldens."""
    from hydrobricks_b200 import engine as _e

    meta, forcing, outlet, records, _ = gu.load(name)
    forcing = dict(forcing)
    radiation = forcing.pop("solar_radiation", None)
    aspect = np.array([_e.ASPECT[u["aspect_class"]] for u in meta["units"]], dtype=np.int32)
    if radiation is not None:
        radiation = np.ascontiguousarray(radiation, dtype=np.float64)
    emul.hbk_emul_set_melt_inputs.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]
    emul.hbk_emul_set_melt_inputs.restype = None
    emul.hbk_emul_set_melt_inputs(aspect.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                  radiation.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
                                  if radiation is not None else None)
    cs, q, rec, _ = run_emul(emul, meta, forcing)
    assert cs.hbk.melt_kind == (_e.MELT_TEMPERATURE_INDEX if radiation is not None else _e.MELT_DEGREE_DAY_ASPECT)
    _close(q, outlet)
    _close(rec["slow"], records["slow_reservoir:water_content"])
    _close(rec["snow"], records[f"{cs.ground_name}_snowpack:snow_content"])
    if cs.glacier_name:
        _close(rec["ice"], records[f"{cs.glacier_name}:ice_content"])


def test_instruction_diet_options_are_bit_identical(emul):
    """-DHBK_OPT (the instruction-diet options of hbk_device.cuh, default 7 = all on) must not change a single bit of
    any result: the build with all of them off gives the same outlet and storages on every golden."""
    plain = build_emul("libhbkemul_opt0.so", ("-DHBK_OPT=0",))
    for name in CASES:
        meta, forcing, _, _, _ = gu.load(name)
        _, qa, ra, _ = run_emul(emul, meta, forcing)
        _, qb, rb, _ = run_emul(plain, meta, forcing)
        assert np.array_equal(qa, qb), name
        for k in ra:
            assert np.array_equal(ra[k], rb[k], equal_nan=True), (name, k)


def test_default_arithmetic_is_bit_identical_to_the_faithful_build(emul, emul_faithful):
    """The shipped arithmetic rounds every quotient like the reference's division (reciprocal + two fused multiply-adds,
    divExact) and keeps the verification sweep: not one bit may differ from the build that writes the divisions as
    divisions, on any golden."""
    for name in CASES:
        meta, forcing, _, _, _ = gu.load(name)
        _, qa, ra, ua = run_emul(emul, meta, forcing)
        _, qb, rb, ub = run_emul(emul_faithful, meta, forcing)
        assert np.array_equal(qa, qb), name
        assert ua == ub
        for k in ra:
            assert np.array_equal(ra[k], rb[k], equal_nan=True), (name, k)


PENDING_GPU = sorted(p.stem for pat in ("splitter_threshold_*.npz", "slow_order_*.npz") for p in gu.GOLDEN.glob(pat))


@pytest.mark.parametrize("name", PENDING_GPU)
def test_threshold_splitter_and_helper_process_order_on_host(name, emul):
    """snow_rain:threshold (SplitterSnowRainThreshold.cpp:45-54, hbk_structure.splitter_kind = 1) and the slow-reservoir
    process order of core/tests/src/helpers.cpp (slow_process_order = 1): the device step functions against vectors of
    the reference."""
    meta, forcing, outlet, records, _ = gu.load(name)
    cs, q, rec, _ = run_emul(emul, meta, forcing)
    assert (cs.hbk.splitter_kind == 1) if "threshold" in name else (cs.hbk.slow_process_order == 1)
    _close(q, outlet)
    _close(rec["slow"], records["slow_reservoir:water_content"])
    _close(rec["snow"], records[f"{cs.ground_name}_snowpack:snow_content"])


def test_c1_full_period_on_host(emul):
    """BASELINE config C1 (ch_sitter_appenzell, 35 elevation bands, 1981-2020 = 14 610 daily steps, RK4): the device
    arithmetic over the whole 40 years against the reference's outlet, forcing rebuilt from the station series with the
    reference's formulas (forcing.py:1061-1113) as in tests/test_oracle_golden.py."""
    meta, _, outlet, _, extra = gu.load("c1_sitter_rk4_full")
    elev = np.array([u["elevation"] for u in meta["units"]])
    sp = meta["spatialisation"]
    t = extra["station_t"][:, None] + sp["temperature"]["gradient"] * (elev - sp["temperature"]["ref_elevation"]) / 100
    p = extra["station_p"][:, None] * sp["precipitation"]["correction_factor"]
    p = p * (1 + sp["precipitation"]["gradient"] * (elev - sp["precipitation"]["ref_elevation"]) / 100)
    p[p < 0] = 0
    pet = np.repeat(extra["station_pet"][:, None], len(elev), axis=1)
    _, q, _, unconverged = run_emul(emul, meta, {"p": p, "t": t, "pet": pet})
    assert len(q) == 14610 and unconverged == 0
    _close(q, outlet)


@pytest.fixture(scope="module")
def emul_plain_bisection():
    return build_emul("libhbkemul_plainbisect.so", ["-DHBK_SEQ_FAST_BISECTION=0"])


SEQ_CASES = [b for b in CASES if any(s in b for s in ("crank_nicolson", "implicit_euler"))]


@pytest.mark.parametrize("name", SEQ_CASES)
def test_fast_bisection_replays_the_reference_halving(name, emul, emul_plain_bisection):
    """crank_nicolson / implicit_euler (SolverCrankNicolson.cpp:12-63, SolverImplicitEuler.cpp:12-61): the shipped
    bisection decides midpoints outside a certified root bracket without evaluating the rate laws; against the plain loop
    (every midpoint evaluated, -DHBK_SEQ_FAST_BISECTION=0) the results must agree to the last bits (a decision can only
    flip within round-off of the root: < 1e-12 mm) at a fraction of the evaluations."""
    meta, forcing, outlet, records, _ = gu.load(name)
    for L in (emul, emul_plain_bisection):
        L.hbk_emul_seq_evaluations.restype = ctypes.c_longlong
        L.hbk_emul_seq_evaluations.argtypes = [ctypes.c_int]
        L.hbk_emul_seq_evaluations(1)
    _, q_fast, rec_fast, _ = run_emul(emul, meta, forcing)
    _, q_plain, rec_plain, _ = run_emul(emul_plain_bisection, meta, forcing)
    n_fast, n_plain = emul.hbk_emul_seq_evaluations(1), emul_plain_bisection.hbk_emul_seq_evaluations(1)
    assert np.max(np.abs(q_fast - q_plain)) <= 1e-11 * max(1.0, float(np.max(np.abs(q_plain))))
    assert np.nanmax(np.abs(rec_fast["slow"] - rec_plain["slow"])) <= 1e-11
    assert (q_fast == q_plain).mean() > 0.99  # bit-identical but for the rare midpoint within round-off of the root
    assert n_plain > 2.5 * n_fast, (n_plain, n_fast)
    _close(q_fast, outlet)


@pytest.mark.parametrize("kind", ["melt:degree_day_aspect", "melt:temperature_index"])
@pytest.mark.parametrize("solver", ["rk4", "crank_nicolson"])
def test_snow_retention_with_the_melt_variants_on_host(kind, solver, emul):
    """Liquid water kept in the snowpacks (outflow:snow_holding, no refreezing: that one needs melt:degree_day) together
    with melt:degree_day_aspect / melt:temperature_index, a combination no golden covers: the device's retention step
    (snowpackStepRetention) takes its degree-day factor from the same per-unit / per-step slots as the plain step; against
    the oracle restatement (pinned on each of the two features separately)."""
    import sys
    sys.path.insert(0, str(gu.GOLDEN.parent.parent / "tools"))
    import ref_melt_cases as rm
    from hydrobricks_b200 import engine as _e
    from hydrobricks_b200 import models
    from oracle import hb_oracle

    m0, units, forcing = rm.build("python", kind, solver, glacier=True, nsoil=1, n_units=6, n_steps=500, seed=23)
    covers = ["ground", "glacier"]
    m = models.Socont(solver=solver, soil_storage_nb=1, surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, snow_melt_process=kind, record_all=True, glacier_infinite_storage=False,
                      snow_water_retention_process="outflow:snow_holding", rain_to_snowpack=True, backend="python")
    s = m.settings
    for st in m0.settings.get_structure():  # the melt-case tool's parameter values
        for el in st["hydro_unit_bricks"] + st["sub_basin_bricks"] + st["hydro_unit_splitters"]:
            for p in el.get("parameters", []):
                if p["name"] not in ("infinite_storage", "no_melt_when_snow_cover"):
                    assert s.set_parameter_value(el["name"], p["name"], p["value"])
            for proc in el.get("processes", []):
                for p in proc["parameters"]:
                    assert s.set_parameter_value(el["name"], p["name"], p["value"])
    assert s.set_parameter_value("type:snowpack", "water_holding_capacity", 0.18)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(499, "D")), 1, "day")
    meta = {"structures": s.get_structure(), "units": units, "solver": solver}
    om = hb_oracle.OracleModel(meta["structures"], units, solver, 500)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    for label in ("slow_reservoir:water_content", "ground_snowpack:snow_content", "glacier:ice_content"):
        om.record(label)
    q_ref = om.run()
    forcing = dict(forcing)
    radiation = forcing.pop("solar_radiation", None)
    aspect = np.array([_e.ASPECT[u["aspect_class"]] for u in units], dtype=np.int32)
    if radiation is not None:
        radiation = np.ascontiguousarray(radiation, dtype=np.float64)
    emul.hbk_emul_set_melt_inputs.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]
    emul.hbk_emul_set_melt_inputs.restype = None
    emul.hbk_emul_set_melt_inputs(aspect.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                  radiation.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if radiation is not None else None)
    cs, q, rec, _ = run_emul(emul, meta, forcing)
    assert cs.hbk.snow_retention == (_e.RETENTION_HOLDING | _e.RETENTION_RAIN_TO_SNOWPACK)
    _close(q, q_ref)
    _close(rec["slow"], om.records["slow_reservoir:water_content"])
    _close(rec["snow"], om.records["ground_snowpack:snow_content"])
    _close(rec["ice"], om.records["glacier:ice_content"])
    assert q_ref.max() > 0
