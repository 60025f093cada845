"""Randomised parity cases for the device arithmetic (TEST INFRASTRUCTURE): small Socont basins with parameter sets
drawn over the reference ranges — one third of them with a soil capacity A below 50 mm, one third with A = 0, where the
capacity / overflow branches of WaterContainer::ApplyConstraints bind — for every solver and structure variant.

    case(seed, solver)             -> structures, units, forcing, parameter values
    python tools/fuzz_cases.py ref <path to a compiled reference _hydrobricks*.so> <solver> <n> <out.npz>
                                   -> outlets of the first n cases through that reference build (used to compare the
                                      reference with its own -march=native -ffp-contract=fast build)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

U, T = 5, 500


def case(seed, solver, backend="python", retention=False):
    """`retention`: liquid water kept in the snowpacks (outflow:snow_holding), refrozen (refreeze:degree_day) for two
    cases out of three and with the rain routed through the snowpack for one out of two, whc / cfr drawn over the
    reference's HBV ranges (models/hbv.py:367-370)."""
    import bench
    from hydrobricks_b200 import models

    glacier = seed % 2 == 0
    nsoil = 1 + (seed // 2) % 2
    quick = "socont_runoff" if (seed // 4) % 3 else "linear_storage"
    rng = np.random.default_rng(seed)
    units = sorted(bench.hydro_units(U, seed=seed, glacier=glacier), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(T, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=nsoil, surface_runoff=quick, land_cover_names=covers,
                      land_cover_types=covers, backend=backend,
                      **({"snow_water_retention_process": "outflow:snow_holding",
                          "snow_refreezing_process": "refreeze:degree_day" if seed % 3 else None,
                          "rain_to_snowpack": (seed // 3) % 2 == 1} if retention else {}))
    vals = {"a_snow": rng.uniform(2, 20), "A": rng.choice([0.0, rng.uniform(0, 50), rng.uniform(0, 3000)]),
            "k_slow_1": 10 ** rng.uniform(-4, 0), "prec_t_start": rng.uniform(-2, 1)}
    vals["prec_t_end"] = vals["prec_t_start"] + rng.uniform(0.1, 3)
    if quick == "socont_runoff":
        vals["beta"] = rng.uniform(100, 30000)
    else:
        vals["k_quick"] = 10 ** rng.uniform(-3, 0)
    if nsoil == 2:
        vals.update({"percol": rng.uniform(0, 10), "k_slow_2": 10 ** rng.uniform(-4, 0)})
    if glacier:
        vals.update({"a_ice": rng.uniform(2, 20), "k_snow": 10 ** rng.uniform(-3, 0), "k_ice": 10 ** rng.uniform(-3, 0)})
    if retention:
        vals["whc"] = rng.uniform(0, 0.2)
        if seed % 3:
            vals["cfr"] = rng.uniform(0, 0.1)
    for alias, v in vals.items():
        comp, name = m.aliases[alias]
        assert m.settings.set_parameter_value(comp, name, float(v)), alias
    forcing = {"precipitation": p2, "temperature": t2, "pet": e2}
    return m, units, forcing, vals


def run_oracle(seed, solver, retention=False):
    from oracle import hb_oracle

    m, units, forcing, vals = case(seed, solver, retention=retention)
    structures = m.settings.get_structure()
    om = hb_oracle.OracleModel(structures, units, solver, T)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    om.record("slow_reservoir:water_content")
    if retention:
        om.record("ground_snowpack:snow_content")
    om.run()
    return {"structures": structures, "units": units, "solver": solver}, forcing, om, vals


def main():
    import importlib.util

    _, _, so, solver, n, out = sys.argv
    spec = importlib.util.spec_from_file_location("_hydrobricks", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.init()
    sys.path.insert(0, str(ROOT / "tools"))
    import make_goldens as mg

    res = {}
    for seed in range(int(n)):
        m, units, forcing, vals = case(seed, solver, backend=ref)
        m.settings.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(T - 1, "D")), 1, "day")
        q, _ = mg.run_ref_module(ref, m.settings, units, forcing)
        res[f"q{seed}"], res[f"A{seed}"] = np.asarray(q), float(vals["A"])
    np.savez(out, **res)


if __name__ == "__main__":
    main()
