"""HBV-96 (python/src/hydrobricks/models/hbv.py, hbv_96.py) through the UNMODIFIED reference — its Python class builds the
structure, its core runs it (oracle/_ref) — and through the oracle restatement. ORACLE ONLY: the CUDA engine does not run
this model family; this pins the restatement of infiltration:hbv, et:hbv, capillary:hbv, runoff:hbv and
routing:hbv (the triangular unit hydrograph with its delivery schedule) for when it does. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

VALUES = [("open_snowpack", "degree_day_factor", 3.5), ("open_snowpack", "water_holding_capacity", 0.12),
          ("open_snowpack", "refreezing_factor", 0.06), ("open", "beta", 2.5), ("soil_moisture", "capacity", 180.0),
          ("soil_moisture", "lp", 0.8), ("upper_zone", "max_capillary_flux", 0.4), ("upper_zone", "percolation_rate", 1.2),
          ("upper_zone", "response_factor", 0.03), ("upper_zone", "alpha", 0.7), ("lower_zone", "response_factor", 0.05),
          ("routing", "maxbas", 3.4)]


def build(hbpy, solver, n_units=6, n_steps=730, seed=71, values=VALUES, **options):
    """hbpy: the reference Python package (tools/refpy.load_reference_python). Returns (settings module object, units,
    forcing)."""
    import bench

    units = sorted(bench.hydro_units(n_units, seed=seed), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    m = hbpy.models.HBV96(solver=solver, record_all=True, land_cover_names=["open"], land_cover_types=["open"], **options)
    s = m.settings.settings
    for comp, name, v in values:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    for u in units:
        u["land_covers"] = [("open", 1.0)]
    return s, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(hbpy, ref, solver, **kw):
    import make_goldens as mg

    s, units, forcing = build(hbpy, solver, **kw)
    q, rec = mg.run_ref_module(ref, s, units, forcing)
    rec = {k: v for k, v in rec.items() if not k.endswith(":internal_storage")}  # routing:hbv's in-transit water (a process value)
    meta = {"structures": s.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": f"reference Python class HBV96 + reference core (tools/ref_hbv_case.py), {kw}"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    hbpy = refpy.load_reference_python(backend="ref")
    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "heun_explicit", "euler_explicit", "crank_nicolson"):
        for kw in (dict(), dict(n_units=3, n_steps=400, seed=5)):
            meta, forcing, q, rec = run_reference(hbpy, ref, solver, **kw)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "max abs", float(np.max(np.abs(om.outlet - q))),
                  "labels differing", len(bad), bad[:4], "mean q", float(q.mean()))
