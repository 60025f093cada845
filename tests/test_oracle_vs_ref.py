"""Live pinning of the oracle restatement against the compiled, UNMODIFIED reference core (oracle/_ref), in the build
container only (needs /root/reference for the reference's Python layer and its bundled meteo files): the same cases
tools/ref_cases.py and tools/ref_actions_case.py sweep in full, here a subset that runs in seconds. Bit-exact: outlet
discharge and every recorded series."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.skipif(not Path("/root/reference/python/src/hydrobricks").exists() or
                                not list((ROOT / "oracle" / "_ref").glob("_hydrobricks*.so")),
                                reason="reference sources / compiled reference not available")


@pytest.mark.parametrize("solver, kw", [
    ("rk4", dict(soil_storage_nb=2, with_glacier=True)),
    ("crank_nicolson", dict(soil_storage_nb=2, with_glacier=True, infinite_storage=False)),
    ("implicit_euler", dict(soil_storage_nb=1, with_glacier=False)),
    ("exponential_euler", dict(soil_storage_nb=1, with_glacier=True, surface_runoff="linear_storage")),
    ("analytic_linear", dict(soil_storage_nb=2, with_glacier=True, snow_sublimation_process="sublimation:pet")),
])
def test_oracle_bit_exact_against_compiled_reference(solver, kw):
    import ref_cases
    import refpy
    from oracle import hb_oracle

    hb = refpy.load_reference_python("ref")
    model, hu, forcing, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-06-30", n_units=8, **kw)
    structures, units, fdata, q_ref, rec_ref = ref_cases.collect(model, hu, forcing)
    om = hb_oracle.OracleModel(structures, units, solver, len(q_ref))
    for k, v in fdata.items():
        om.set_forcing(k, v)
    for label in rec_ref:
        om.record(label)
    q = om.run()
    assert np.array_equal(q, q_ref)
    for label, ref in rec_ref.items():
        assert np.array_equal(om.records[label], ref, equal_nan=True), label


@pytest.mark.parametrize("solver, flags, area_scaling", [
    ("rk4", (True, True, True), True),
    ("crank_nicolson", (True, True, False), False),
])
def test_oracle_actions_bit_exact_against_compiled_reference(solver, flags, area_scaling):
    import ref_actions_case as rc
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    meta, forcing, q, rec, _, areas, volumes = rc.run_reference(ref, solver, *flags, area_scaling=area_scaling)
    om = rc.run_oracle(meta, forcing, areas, volumes, list(rec))
    assert np.array_equal(om.outlet, q)
    for label in rec:
        assert np.array_equal(om.records[label], rec[label], equal_nan=True), label


@pytest.mark.parametrize("kind, solver, kw", [
    ("melt:degree_day_aspect", "heun_explicit", dict(glacier=True, nsoil=1)),
    ("melt:temperature_index", "implicit_euler", dict(glacier=True, nsoil=2)),
])
def test_oracle_melt_variants_bit_exact_against_compiled_reference(kind, solver, kw):
    import ref_melt_cases as rm
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    meta, forcing, q, rec = rm.run_reference(ref, kind, solver, n_steps=500, **kw)
    om = rm.run_oracle(meta, forcing, list(rec))
    assert np.array_equal(om.outlet, q)
    for label in rec:
        assert np.array_equal(om.records[label], rec[label], equal_nan=True), label


def test_oracle_snow_slide_bit_exact_against_compiled_reference():
    import ref_melt_cases as rm
    import ref_snow_slide_case as rs
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    meta, forcing, q, rec = rs.run_reference(ref, "heun_explicit", glacier=True, nsoil=2, n_steps=500, skip_glaciers=False)
    om = rm.run_oracle(meta, forcing, list(rec))
    assert np.array_equal(om.outlet, q)
    for label in rec:
        assert np.array_equal(om.records[label], rec[label], equal_nan=True), label
    snow = rec["ground_snowpack:snow_content"]
    assert snow[-1, -1] > 4 * snow[-1, 0]  # the snow did slide to the lowest unit


@pytest.mark.parametrize("solver", ["rk4", "crank_nicolson"])
def test_oracle_spinup_bit_exact_against_compiled_reference(solver):
    """ModelHydro::RunSpinup (ModelHydro.cpp:135-181): the first steps replayed unlogged, clock rewound, states kept."""
    import ref_cases
    import refpy
    from oracle import hb_oracle

    hb = refpy.load_reference_python("ref")
    model, hu, forcing, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-03-31", n_units=8, soil_storage_nb=2,
                                                 with_glacier=True, infinite_storage=False, spinup_days=200)
    structures, units, fdata, q_ref, rec_ref = ref_cases.collect(model, hu, forcing)
    om = hb_oracle.OracleModel(structures, units, solver, len(q_ref), spinup_steps=200)
    for k, v in fdata.items():
        om.set_forcing(k, v)
    for label in rec_ref:
        om.record(label)
    q = om.run()
    assert q_ref[0] > 0  # the run does not start from empty storages
    assert np.array_equal(q, q_ref)
    for label, ref in rec_ref.items():
        assert np.array_equal(om.records[label], ref, equal_nan=True), label


@pytest.mark.parametrize("solver", ["rk4", "crank_nicolson"])
def test_oracle_two_day_step_bit_exact_against_compiled_reference(solver):
    """A time step of two days (TimeMachine step 2 'day'): the oracle's dt handling against the reference, including the
    reference's own artefacts at that step length (land-cover water driven negative by a direct outflow of content/day
    over two days, logged and reset to 0, WaterContainer.cpp:212-215)."""
    import ref_threshold_case as rt
    import refpy
    from oracle import hb_oracle

    ref = refpy.load_extension("ref")
    ref.init()
    n = 200
    m, units, forcing = rt.build(ref, solver, glacier=True, nsoil=2, n_steps=n)
    s = m.settings
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(2 * (n - 1), "D")), 2, "day")
    basin = ref.SettingsBasin()
    for u in units:
        basin.add_hydro_unit(u["id"], u["area"], u["elevation"])
        basin.add_hydro_unit_property_double("slope", u["slope"][0], u["slope"][1])
        for name, frac in u["land_covers"]:
            basin.add_land_cover(name, "", frac)
    model = ref.ModelHydro()
    model.init_with_basin(s, basin)
    time = 44605.0 + 2.0 * np.arange(n)
    ids = np.array([u["id"] for u in units], dtype=np.int32)
    for k, v in forcing.items():
        assert model.create_time_series(k, time, ids, np.ascontiguousarray(v))
    assert model.attach_time_series_to_hydro_units()
    model.run()
    q_ref = np.asarray(model.get_outlet_discharge())
    rec = {label: model.get_hydro_unit_values(label) for label in model.get_recorded_hydro_unit_labels()}
    om = hb_oracle.OracleModel(s.get_structure(), units, solver, n, dt=2.0)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    for label in rec:
        om.record(label)
    assert np.array_equal(om.run(), q_ref)
    for label, r in rec.items():
        assert np.array_equal(om.records[label], r, equal_nan=True), label


@pytest.mark.parametrize("solver", ["rk4", "crank_nicolson"])
def test_oracle_snow_retention_fuzz_bit_exact_against_compiled_reference(solver):
    """The randomised snowpack-retention cases of tests/test_device_math_fuzz.py (holding / refreezing / rain through the
    snowpack, whc and cfr over their ranges, 1-2 soil stores, both quick-flow laws, soil capacities down to 0) through the
    compiled reference core: the oracle those device-arithmetic checks lean on is bit-exact there too."""
    import fuzz_cases
    import make_goldens as mg
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    for seed in range(12):
        m, units, forcing, vals = fuzz_cases.case(seed, solver, backend=ref, retention=True)
        m.settings.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(fuzz_cases.T - 1, "D")), 1, "day")
        q_ref, _ = mg.run_ref_module(ref, m.settings, units, forcing)
        _, _, om, _ = fuzz_cases.run_oracle(seed, solver, retention=True)
        assert np.array_equal(om.outlet, np.asarray(q_ref)), f"seed {seed} ({vals})"
