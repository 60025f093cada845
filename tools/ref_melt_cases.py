"""Melt-process variants of the reference's Socont (`snow_melt_process`): melt:degree_day_aspect (per-unit aspect class
picks one of three degree-day factors, ProcessMeltDegreeDayAspect.cpp) and melt:temperature_index (melt factor +
radiation coefficient x solar radiation, ProcessMeltTemperatureIndex.cpp), run through the UNMODIFIED reference core
(oracle/_ref) and through the oracle restatement. Build container only. The CUDA engine does not implement these two
yet; this pins the oracle for when it does."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

ASPECTS = ["north", "S", "east", "W", "south", "N", "west", "E"]


def build(backend, kind, solver, glacier=True, nsoil=2, n_units=8, n_steps=730, seed=21):
    import bench
    from hydrobricks_b200 import models

    rng = np.random.default_rng(seed)
    units = sorted(bench.hydro_units(n_units, seed=seed, glacier=glacier), key=lambda u: -u["elevation"])
    for i, u in enumerate(units):
        u["aspect_class"] = ASPECTS[i % len(ASPECTS)]
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    doy = np.fmod(np.arange(n_steps), 365.25) / 365.25
    rad = (120 - 100 * np.cos(2 * np.pi * doy))[:, None] * rng.uniform(0.7, 1.2, (1, n_units)) + rng.uniform(0, 20, (n_steps, n_units))
    forcing = {"precipitation": p2, "temperature": t2, "pet": e2}
    if kind == "melt:temperature_index":
        forcing["solar_radiation"] = rad
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, snow_melt_process=kind, record_all=True,
                      glacier_infinite_storage=False, backend=backend)
    s = m.settings
    vals = [("slow_reservoir", "capacity", 180.0), ("slow_reservoir", "response_factor", 0.03),
            ("surface_runoff", "beta", 1200.0)]
    if nsoil == 2:
        vals += [("slow_reservoir", "percolation_rate", 0.7), ("slow_reservoir_2", "response_factor", 0.006)]
    melted = ["ground_snowpack"] + (["glacier_snowpack", "glacier"] if glacier else [])
    for comp in melted:
        ice = comp == "glacier"
        if kind == "melt:degree_day_aspect":
            vals += [(comp, "degree_day_factor_n", 6.0 if ice else 2.5), (comp, "degree_day_factor_s", 9.5 if ice else 5.5),
                     (comp, "degree_day_factor_ew", 7.5 if ice else 3.75), (comp, "melting_temperature", 0.5 if ice else 0.0)]
        else:
            vals += [(comp, "melt_factor", 5.0 if ice else 2.0), (comp, "radiation_coefficient", 0.02 if ice else 0.012),
                     (comp, "melting_temperature", 0.25)]
    if glacier:
        vals += [("glacier_area_rain_snowmelt_storage", "response_factor", 0.2),
                 ("glacier_area_icemelt_storage", "response_factor", 0.35)]
    for comp, name, v in vals:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    return m, units, forcing


def run_reference(ref, kind, solver, **kw):
    import make_goldens as mg

    m, units, forcing = build(ref, kind, solver, **kw)
    q, rec = mg.run_ref_module(ref, m.settings, units, forcing)
    meta = {"structures": m.settings.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": f"reference core driven directly (tools/ref_melt_cases.py), snow_melt_process={kind}"}
    return meta, forcing, np.asarray(q), rec


def run_oracle(meta, forcing, labels):
    from oracle import hb_oracle

    n = len(next(iter(forcing.values())))
    om = hb_oracle.OracleModel(meta["structures"], meta["units"], meta["solver"], n)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    for label in labels:
        om.record(label)
    om.run()
    return om


if __name__ == "__main__":
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    for kind in ("melt:degree_day_aspect", "melt:temperature_index"):
        for solver in ("rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "analytic_linear"):
            for kw in (dict(glacier=True, nsoil=2), dict(glacier=False, nsoil=1)):
                meta, forcing, q, rec = run_reference(ref, kind, solver, **kw)
                om = run_oracle(meta, forcing, list(rec))
                bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
                print(kind, solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad),
                      bad[:3], "mean q", float(q.mean()))
