"""Pins the CPU restatement (oracle/hb_oracle.cpp) against vectors produced by the unmodified reference.

* tests/golden/*.npz (tools/make_goldens.py): bit-exact outlet and every recorded series.
* The reference's own known-answer tests: SolverTest.cpp:168-170,205-207,242-244 (tol 1e-6 there) and
  ModelSocontTest.cpp:604-613 (tol 1e-6 there).
"""
import numpy as np
import pytest

import golden_utils as gu

CASES = [b for b in gu.bundles() if not b.startswith(("c1_", "actions_", "actions2_"))]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_bit_exact(name):
    meta, forcing, outlet, records, _ = gu.load(name)
    om = gu.run_oracle(meta, forcing)
    assert gu.same(om.outlet, outlet), f"outlet differs, max abs {np.max(np.abs(om.outlet - outlet))}"
    for label, ref in records.items():
        assert gu.same(om.records[label], ref), label


NEXT_CASES = sorted("next/" + p.stem for p in (gu.GOLDEN / "next").glob("*.npz"))


@pytest.mark.parametrize("name", NEXT_CASES)
def test_oracle_matches_reference_on_options_the_engine_lacks(name):
    """tests/golden/next: what is pinned bit for bit in the oracle but not built in the CUDA engine — its structure
    compiler refuses such structures (outside the engine's parity lists). At the moment HBV-96 (models/hbv.py, hbv_96.py:
    infiltration:hbv, et:hbv, capillary:hbv, runoff:hbv, routing:hbv with its unit-hydrograph schedule) and GR4J / GR6J
    (models/gr4j.py, gr6j.py: interception, production store, the two unit hydrographs, routing and exponential stores),
    generated through the reference's own Python classes (tools/ref_hbv_case.py, ref_gr4j_case.py); the previous occupants — snowpack water retention and
    refreezing, HBV's snow routine — moved to tests/golden when the engine learnt them."""
    from hydrobricks_b200 import structure

    meta, forcing, outlet, records, _ = gu.load(name)
    om = gu.run_oracle(meta, forcing)
    assert gu.same(om.outlet, outlet)
    for label, ref in records.items():
        assert gu.same(om.records[label], ref), label
    with pytest.raises(structure.UnsupportedStructure):
        structure.CompiledStructure(meta["structures"], meta["solver"])


# Expected vectors copied as data from the reference's tests (known answers, tolerance 1e-6 there).
KAT_LINEAR = {
    "euler_explicit": ([0.000000, 0.000000, 3.000000, 5.100000, 6.570000, 4.599000, 3.219300, 2.253510, 1.577457,
                        1.104220, 0.772954, 0.541068, 0.378747, 0.265123, 0.185586, 0.129910, 0.090937, 0.063656,
                        0.044559, 0.031191], 0.072780),
    "heun_explicit": ([0.000000, 1.500000, 3.667500, 5.282288, 4.985304, 3.714052, 2.766968, 2.061392, 1.535737,
                       1.144124, 0.852372, 0.635017, 0.473088, 0.352450, 0.262576, 0.195619, 0.145736, 0.108573,
                       0.080887, 0.060261], 0.176056),
    "runge_kutta": ([0.000000, 1.361250, 3.600090, 5.258707, 5.126222, 3.797698, 2.813477, 2.084329, 1.544149,
                     1.143964, 0.847491, 0.627853, 0.465137, 0.344591, 0.255286, 0.189125, 0.140111, 0.103800,
                     0.076899, 0.056969], 0.162852),
}


@pytest.mark.parametrize("solver", list(KAT_LINEAR))
def test_kat_linear_storage(solver):
    """core/tests/src/SolverTest.cpp:148-257."""
    meta, forcing, _, _, _ = gu.load(f"kat_linear_storage_{solver}")
    om = gu.run_oracle(meta, forcing)
    expected, final_content = KAT_LINEAR[solver]
    assert np.allclose(om.outlet, expected, atol=1e-6, rtol=0)
    content = om.records["storage:water_content"][-1, 0]
    assert abs(content - final_content) < 1e-6
    assert abs(30.0 - om.outlet.sum() - content) < 1e-14  # SolverTest.cpp:182


@pytest.mark.parametrize("solver", ["analytic_linear", "implicit_euler", "crank_nicolson", "exponential_euler"])
def test_kat_linear_storage_sequential_solvers(solver):
    """core/tests/src/SolverTest.cpp:259-430: the sequential solvers on the linear-storage fixture against the
    closed forms the reference's tests use (tolerances 1e-9 analytic, 1e-8 the others)."""
    meta, forcing, outlet, _, _ = gu.load(f"kat_linear_storage_{solver}")
    om = gu.run_oracle(meta, forcing)
    assert gu.same(om.outlet, outlet)
    precip = [0.0, 10.0, 10.0, 10.0] + [0.0] * 16
    k = float(np.float32(0.3))
    storage, tol = 0.0, (1e-9 if solver == "analytic_linear" else 1e-8)
    for j, p in enumerate(precip):
        if solver in ("analytic_linear", "exponential_euler"):
            new = storage * np.exp(-k) + p / k * (1.0 - np.exp(-k))
            out = p - (new - storage)
        elif solver == "implicit_euler":
            new = (storage + p) / (1.0 + k)
            out = k * new
        else:
            new = (storage * (1.0 - k / 2) + p) / (1.0 + k / 2)
            out = k * (storage + new) / 2
        assert abs(om.outlet[j] - out) < tol, (j, om.outlet[j], out)
        storage = new
    content = om.records["storage:water_content"][-1, 0]
    assert abs(content - storage) < tol
    assert abs(30.0 - om.outlet.sum() - content) < tol


def test_kat_socont_quick_discharge():
    """core/tests/src/ModelSocontTest.cpp:575-626."""
    meta, forcing, _, _, _ = gu.load("kat_socont_quick_rk4")
    om = gu.run_oracle(meta, forcing)
    expected = [0.0, 0.339098404522742, 6.14339396606552, 15.9871208705154, 18.1637996149964, 18.8240003953163,
                18.3973261932905, 14.3377406789303, 10.4729112050052, 7.92405289891611]
    assert np.allclose(om.outlet, expected, atol=1e-6, rtol=0)


def test_c1_sitter_spatialisation_and_outlet():
    """BASELINE config C1: the bundled catchment, 1981-2020, RK4. Forcing is rebuilt from the station
    series with the reference's formulas (forcing.py:1061-1113) and must reproduce the recorded columns
    bit for bit; the outlet must then match the reference run bit for bit."""
    meta, _, outlet, _, extra = gu.load("c1_sitter_rk4_full")
    elev = np.array([u["elevation"] for u in meta["units"]])
    sp = meta["spatialisation"]
    t = extra["station_t"][:, None] + sp["temperature"]["gradient"] * (elev - sp["temperature"]["ref_elevation"]) / 100
    p = extra["station_p"][:, None] * sp["precipitation"]["correction_factor"]
    p = p * (1 + sp["precipitation"]["gradient"] * (elev - sp["precipitation"]["ref_elevation"]) / 100)
    p[p < 0] = 0
    pet = np.repeat(extra["station_pet"][:, None], len(elev), axis=1)
    assert gu.same(t[:, 0], extra["spatial_first_t"]) and gu.same(t[:, -1], extra["spatial_last_t"])
    assert gu.same(p[:, 0], extra["spatial_first_p"]) and gu.same(p[:, -1], extra["spatial_last_p"])
    om = gu.run_oracle(meta, {"p": p, "t": t, "pet": pet}, labels=[])
    assert gu.same(om.outlet, outlet)


ACTION_CASES = [b for b in gu.bundles("actions_")]


@pytest.mark.parametrize("name", ACTION_CASES)
def test_oracle_actions_match_reference_bit_exact(name):
    """Dated actions (ActionGlacierEvolutionDeltaH, ActionLandCoverChange, ActionGlacierSnowToIceTransformation):
    oracle restatement + the host scheduler (hydrobricks_b200/schedule.py) against the reference run."""
    import sys

    sys.path.insert(0, str(gu.GOLDEN.parent.parent / "tools"))
    import ref_actions_case as rc

    meta, forcing, outlet, records, extra = gu.load(name)
    om = rc.run_oracle(meta, forcing, extra["spatial_dh_areas"], extra["spatial_dh_volumes"], meta["labels"])
    assert gu.same(om.outlet, outlet)
    for label, ref in records.items():
        assert gu.same(om.records[label], ref), label
