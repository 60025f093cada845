"""The slow-reservoir process order of the reference's C++ test helper (core/tests/src/helpers.cpp:74-86: et, outflow,
percolation, overflow — hbk_structure.slow_process_order = 1) instead of the Python layer's (et, outflow, overflow,
percolation): the order decides the rate-vector layout and the order of the sums in the flux constraint. Two soil
stores, no glacier, built through the SettingsModel API; reference core vs oracle. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def build(hb, solver, n_units=6, n_steps=730, seed=51):
    import bench

    s = hb.SettingsModel()
    s.log_all(True)
    s.set_solver(solver)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    s.generate_precipitation_splitters(True, "snow_rain:linear")
    s.add_land_cover_brick("ground", "generic_land_cover")
    s.generate_snowpacks("melt:degree_day")
    s.select_hydro_unit_brick("ground")
    s.add_brick_process("infiltration", "infiltration:socont", "slow_reservoir")
    s.add_brick_process("runoff", "outflow:rest", "surface_runoff")
    s.set_process_outputs_as_static()
    s.add_hydro_unit_brick("slow_reservoir", "storage")
    s.add_brick_parameter("capacity", 200.0)
    s.add_brick_process("et", "et:socont")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_brick_process("percolation", "percolation:constant", "slow_reservoir_2")
    s.add_brick_process("overflow", "overflow", "outlet")
    s.add_hydro_unit_brick("slow_reservoir_2", "storage")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_hydro_unit_brick("surface_runoff", "storage")
    s.add_brick_process("runoff", "runoff:socont", "outlet")
    s.add_logging_to("outlet")
    for comp, name, v in [("slow_reservoir", "capacity", 40.0), ("slow_reservoir", "response_factor", 0.05),
                          ("slow_reservoir", "percolation_rate", 1.5), ("slow_reservoir_2", "response_factor", 0.01),
                          ("surface_runoff", "beta", 900.0), ("ground_snowpack", "degree_day_factor", 4.0)]:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    units = sorted(bench.hydro_units(n_units, seed=seed), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    return s, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    s, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, s, units, forcing)
    meta = {"structures": s.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "reference core driven directly (tools/ref_slow_order_case.py), helpers.cpp process order"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "analytic_linear"):
        meta, forcing, q, rec = run_reference(ref, solver)
        om = rm.run_oracle(meta, forcing, list(rec))
        bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
        ovf = rec["slow_reservoir:overflow:output"]
        print(solver, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3],
              "overflow days", int((ovf > 0).sum()), "unconverged", om.unconverged)
