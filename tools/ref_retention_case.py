"""Snowpack liquid-water retention and refreezing (SettingsModel::GenerateSnowpacksWithWaterRetention,
SettingsModel.cpp:608-634; AddSnowpackRefreezing :574-590; ProcessOutflowSnowHolding.cpp:37-57; ProcessRefreezeDegreeDay.cpp
:72-87) on the Socont structure, with and without the rain routed through the snowpack, run through the UNMODIFIED
reference core (oracle/_ref) and through the oracle restatement. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def build(backend, solver, glacier=True, nsoil=2, refreeze=True, rain_to_snowpack=False, n_units=8, n_steps=730, seed=33,
          extras=None):
    import bench
    from hydrobricks_b200 import models

    units = sorted(bench.hydro_units(n_units, seed=seed, glacier=glacier), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    forcing = {"precipitation": p2, "temperature": t2, "pet": e2}
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, record_all=True, glacier_infinite_storage=False,
                      snow_water_retention_process="outflow:snow_holding",
                      snow_refreezing_process="refreeze:degree_day" if refreeze else None,
                      rain_to_snowpack=rain_to_snowpack, backend=backend, **(extras or {}))
    s = m.settings
    vals = [("slow_reservoir", "capacity", 180.0), ("slow_reservoir", "response_factor", 0.03),
            ("surface_runoff", "beta", 1200.0), ("ground_snowpack", "degree_day_factor", 3.5),
            ("ground_snowpack", "water_holding_capacity", 0.12)]
    if refreeze:
        vals += [("ground_snowpack", "refreezing_factor", 0.07)]
    if nsoil == 2:
        vals += [("slow_reservoir", "percolation_rate", 0.7), ("slow_reservoir_2", "response_factor", 0.006)]
    if glacier:
        vals += [("glacier_snowpack", "degree_day_factor", 4.0), ("glacier_snowpack", "water_holding_capacity", 0.08),
                 ("glacier", "degree_day_factor", 7.0), ("glacier_area_rain_snowmelt_storage", "response_factor", 0.2),
                 ("glacier_area_icemelt_storage", "response_factor", 0.35)]
        if refreeze:
            vals += [("glacier_snowpack", "refreezing_factor", 0.03)]
    for comp, name, v in vals:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    return m, units, forcing


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    m, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, m.settings, units, forcing)
    meta = {"structures": m.settings.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": f"reference core driven directly (tools/ref_retention_case.py), {kw}"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "heun_explicit", "euler_explicit", "crank_nicolson"):
        for kw in (dict(glacier=True, nsoil=2), dict(glacier=False, nsoil=1, refreeze=False),
                   dict(glacier=True, nsoil=1, rain_to_snowpack=True), dict(glacier=False, nsoil=2, rain_to_snowpack=True)):
            meta, forcing, q, rec = run_reference(ref, solver, **kw)
            if "--dump" in sys.argv:
                import json
                print(json.dumps(meta["structures"], indent=1)[:6000])
                sys.exit(0)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3],
                  "mean q", float(q.mean()))
