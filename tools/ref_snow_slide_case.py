"""transport:snow_slide (ProcessLateralSnowSlide.cpp): snow above a slope-dependent holding depth slides to the
snowpacks of the connected lower units within the step (instantaneous lateral fluxes). Reference core (oracle/_ref) vs
the oracle restatement. Build container only. The CUDA engine refuses this option (it couples the units of a parameter
set inside a time step); this pins the oracle for when it does not."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def build(backend, solver, glacier=True, nsoil=1, n_units=8, n_steps=730, seed=31, skip_glaciers=True):
    import bench
    from hydrobricks_b200 import models

    rng = np.random.default_rng(seed)
    units = sorted(bench.hydro_units(n_units, seed=seed, glacier=glacier), key=lambda u: -u["elevation"])
    for i, u in enumerate(units):
        u["slope"] = (float(round(rng.uniform(8, 80), 2)), "deg")  # some steeper than max_slope = 75
        # every unit gives to the next one or two lower units (basin order = descending elevation)
        lower = list(range(i + 1, min(i + 3, n_units)))
        w = rng.dirichlet(np.ones(len(lower))) * rng.uniform(0.6, 1.0) if lower else []
        u["lateral"] = [(j, float(x)) for j, x in zip(lower, w)]
    # cold and snowy, so that the holding depth is exceeded
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip * 2.5, temp - 4.0, pet)
    forcing = {"precipitation": p2, "temperature": t2, "pet": e2}
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, record_all=True, snow_redistribution="transport:snow_slide",
                      snow_redistribution_skip_glaciers=skip_glaciers, backend=backend)
    s = m.settings
    for comp, name, v in [("slow_reservoir", "capacity", 180.0), ("slow_reservoir", "response_factor", 0.03),
                          ("surface_runoff", "beta", 1200.0), ("type:snowpack", "degree_day_factor", 3.5)] + \
            ([("slow_reservoir", "percolation_rate", 0.7), ("slow_reservoir_2", "response_factor", 0.006)] if nsoil == 2 else []) + \
            ([("glacier", "degree_day_factor", 7.0), ("glacier_area_rain_snowmelt_storage", "response_factor", 0.2),
              ("glacier_area_icemelt_storage", "response_factor", 0.35)] if glacier else []):
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    return m, units, forcing


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    m, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, m.settings, units, forcing)
    meta = {"structures": m.settings.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "reference core driven directly (tools/ref_snow_slide_case.py), transport:snow_slide"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "euler_explicit", "crank_nicolson"):
        for kw in (dict(glacier=True, nsoil=1), dict(glacier=False, nsoil=2), dict(glacier=True, nsoil=1, skip_glaciers=False)):
            meta, forcing, q, rec = run_reference(ref, solver, **kw)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            snow = rec["ground_snowpack:snow_content"]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3],
                  "max snow", float(np.nanmax(snow)), "mean q", float(q.mean()))
