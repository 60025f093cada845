"""Two glacier covers per hydro unit — glacier_ice + glacier_debris, the land covers of the reference's shipped
Rhone-Gletsch setup (models/socont.py:131-140, parameters.py:1766-1835): each with its own brick, snowpack, ice melt
factor and area fraction, both feeding the same two sub-basin stores. Reference core (oracle/_ref) vs the oracle
restatement. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

COVERS = ["ground", "glacier_ice", "glacier_debris"]


def units_two_glaciers(n_units, seed):
    import bench

    rng = np.random.default_rng(seed)
    units = sorted(bench.hydro_units(n_units, seed=seed, glacier=True), key=lambda u: -u["elevation"])
    for u in units:
        g = dict(u["land_covers"])["glacier"]
        if g > 0:
            debris = round(float(g * rng.uniform(0.1, 0.5)), 4)
            if rng.uniform() < 0.25:
                debris = 0.0  # a unit whose debris-covered part is empty: a null brick
            u["land_covers"] = [("ground", 1.0 - g), ("glacier_ice", g - debris), ("glacier_debris", debris)]
        else:
            u["land_covers"] = [("ground", 1.0), ("glacier_ice", 0.0), ("glacier_debris", 0.0)]
    return units


def build(backend, solver, nsoil=1, n_units=9, n_steps=730, seed=71, infinite=False, transfo=None, melt="melt:degree_day"):
    import bench
    from hydrobricks_b200 import models

    units = units_two_glaciers(n_units, seed)
    if melt == "melt:degree_day_aspect":
        for i, u in enumerate(units):
            u["aspect_class"] = ["north", "S", "east", "W"][i % 4]
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff="socont_runoff", land_cover_names=COVERS,
                      land_cover_types=["ground", "glacier", "glacier"], record_all=True, snow_melt_process=melt,
                      glacier_infinite_storage=infinite, snow_ice_transformation=transfo, backend=backend)
    s = m.settings
    vals = [("slow_reservoir", "capacity", 180.0), ("slow_reservoir", "response_factor", 0.03),
            ("surface_runoff", "beta", 1200.0), ("glacier_area_rain_snowmelt_storage", "response_factor", 0.2),
            ("glacier_area_icemelt_storage", "response_factor", 0.35)]
    if melt == "melt:degree_day":
        vals += [("type:snowpack", "degree_day_factor", 3.5), ("glacier_ice", "degree_day_factor", 8.0),
                 ("glacier_debris", "degree_day_factor", 5.0), ("glacier_debris", "melting_temperature", 0.5),
                 ("glacier_debris_snowpack", "degree_day_factor", 3.0)]
    else:
        for comp, f in (("type:snowpack", 1.0), ("glacier_ice", 2.2), ("glacier_debris", 1.5)):
            vals += [(comp, "degree_day_factor_n", 2.5 * f), (comp, "degree_day_factor_s", 4.5 * f),
                     (comp, "degree_day_factor_ew", 3.5 * f)]
    if nsoil == 2:
        vals += [("slow_reservoir", "percolation_rate", 0.7), ("slow_reservoir_2", "response_factor", 0.006)]
    if transfo == "transform:snow_ice_constant":
        vals += [("glacier_ice_snowpack", "snow_ice_transformation_rate", 0.4),
                 ("glacier_debris_snowpack", "snow_ice_transformation_rate", 0.25)]
    for comp, name, v in vals:
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    return m, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(ref, solver, **kw):
    import make_goldens as mg

    m, units, forcing = build(ref, solver, **kw)
    q, rec = mg.run_ref_module(ref, m.settings, units, forcing)
    meta = {"structures": m.settings.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "reference core driven directly (tools/ref_two_glaciers_case.py), glacier_ice + glacier_debris"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "euler_explicit", "crank_nicolson"):
        for kw in (dict(nsoil=1), dict(nsoil=2, infinite=True), dict(nsoil=1, transfo="transform:snow_ice_constant"),
                   dict(nsoil=1, melt="melt:degree_day_aspect")):
            meta, forcing, q, rec = run_reference(ref, solver, **kw)
            om = rm.run_oracle(meta, forcing, list(rec))
            bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
            print(solver, kw, "outlet mismatches", int(np.sum(om.outlet != q)), "labels differing", len(bad), bad[:3],
                  "mean q", float(q.mean()), "labels", len(rec))
