"""snow_rain:threshold (SplitterSnowRainThreshold.cpp:45-54: all snow at or below the threshold temperature, all rain
above) instead of the linear transition, through the reference core and the oracle. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def build(backend, solver, glacier=True, nsoil=1, n_units=8, n_steps=730, seed=41):
    import bench
    from hydrobricks_b200 import models

    units = sorted(bench.hydro_units(n_units, seed=seed, glacier=glacier), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, record_all=True, snow_rain_process="snow_rain:threshold",
                      glacier_infinite_storage=False, backend=backend)
    s = m.settings
    vals = [("snow_rain_transition", "threshold", 0.75), ("slow_reservoir", "capacity", 180.0),
            ("slow_reservoir", "response_factor", 0.03), ("surface_runoff", "beta", 1200.0),
            ("type:snowpack", "degree_day_factor", 3.5)]
    if nsoil == 2:
        vals += [("slow_reservoir", "percolation_rate", 0.7), ("slow_reservoir_2", "response_factor", 0.006)]
    if glacier:
        vals += [("glacier", "degree_day_factor", 7.0), ("glacier_area_rain_snowmelt_storage", "response_factor", 0.2),
                 ("glacier_area_icemelt_storage", "response_factor", 0.35)]
    for comp, name, v in vals:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    return m, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    m, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, m.settings, units, forcing)
    meta = {"structures": m.settings.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "reference core driven directly (tools/ref_threshold_case.py), snow_rain:threshold splitter"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "euler_explicit", "crank_nicolson"):
        for kw in (dict(glacier=True, nsoil=1), dict(glacier=False, nsoil=2)):
            meta, forcing, q, rec = run_reference(ref, solver, **kw)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3])
