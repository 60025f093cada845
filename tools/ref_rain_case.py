"""Socont without the snow routine: SettingsModel::GeneratePrecipitationSplitters(withSnow = false)
(SettingsModel.cpp:518-529) — the `rain` splitter (SplitterRain.cpp:37-39: precipitation x rain_correction_factor to the
land covers) and no snowpacks —, the remaining bricks as core/tests/src/helpers.cpp:9-107 builds them. Driven through the
reference core and the oracle. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def build(hb, solver, nsoil=2, n_units=6, n_steps=500, seed=57, rain_correction=1.15):
    """hb: a host module with the reference's SettingsModel API (the compiled reference or one of this repo's hosts)."""
    import bench

    units = sorted(bench.hydro_units(n_units, seed=seed), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    s = hb.SettingsModel()
    s.log_all(True)
    s.set_solver(solver)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    s.generate_precipitation_splitters(False)
    s.add_land_cover_brick("ground", "generic_land_cover")
    s.select_hydro_unit_brick("ground")
    s.add_brick_process("infiltration", "infiltration:socont", "slow_reservoir")
    s.add_brick_process("runoff", "outflow:rest", "surface_runoff")
    s.set_process_outputs_as_static()
    s.add_hydro_unit_brick("slow_reservoir", "storage")
    s.add_brick_parameter("capacity", 200.0)
    s.add_brick_process("et", "et:socont")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_brick_process("overflow", "overflow", "outlet")
    if nsoil == 2:
        s.add_brick_process("percolation", "percolation:constant", "slow_reservoir_2")
        s.add_hydro_unit_brick("slow_reservoir_2", "storage")
        s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_hydro_unit_brick("surface_runoff", "storage")
    s.add_brick_process("runoff", "runoff:socont", "outlet")
    s.add_logging_to("outlet")
    vals = [("rain", "rain_correction_factor", rain_correction), ("slow_reservoir", "capacity", 160.0),
            ("slow_reservoir", "response_factor", 0.03), ("surface_runoff", "beta", 900.0)]
    if nsoil == 2:
        vals += [("slow_reservoir", "percolation_rate", 0.6), ("slow_reservoir_2", "response_factor", 0.007)]
    for comp, name, v in vals:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    # the temperature series is attached but unused by the reference (no process reads it)
    return s, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    s, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, s, units, forcing)
    meta = {"structures": s.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "reference core driven directly (tools/ref_rain_case.py), rain splitter, no snowpacks"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "heun_explicit", "euler_explicit", "crank_nicolson"):
        for kw in (dict(nsoil=1), dict(nsoil=2)):
            meta, forcing, q, rec = run_reference(ref, solver, **kw)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3],
                  "mean q", float(q.mean()))
