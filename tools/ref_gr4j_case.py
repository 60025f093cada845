"""GR4J (python/src/hydrobricks/models/gr4j.py) through the UNMODIFIED reference — its Python class builds the structure, its
core runs it (oracle/_ref) — and through the oracle restatement. ORACLE ONLY: the CUDA engine does not run this model family
(SURVEY.md §8 f4); this pins the restatement of interception:gr4j, production:gr4j, et:gr4j and routing:gr4j (two unit
hydrographs, routing store, groundwater exchange). Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

VALUES = [("production_store", "capacity", 320.0), ("uh_input", "exchange_factor", 0.8), ("uh_input", "routing_capacity", 75.0),
          ("uh_input", "uh_base_time", 2.3)]


GR6J_VALUES = VALUES + [("uh_input", "exchange_threshold", 0.2), ("uh_input", "exp_store_coeff", 5.5)]


def build(hbpy, n_units=5, n_steps=730, seed=81, values=None, model="GR4J", **options):
    import bench

    units = sorted(bench.hydro_units(n_units, seed=seed), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(n_steps, seed=seed)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    m = getattr(hbpy.models, model)(record_all=True, land_cover_names=["open"], land_cover_types=["open"], **options)
    s = m.settings.settings
    for comp, name, v in (values if values is not None else (GR6J_VALUES if model == "GR6J" else VALUES)):
        assert s.set_parameter_value(comp, name, v), (comp, name)
    s.set_timer("1981-01-01", str(np.datetime64("1981-01-01") + np.timedelta64(n_steps - 1, "D")), 1, "day")
    for u in units:
        u["land_covers"] = [("open", 1.0)]
    return m, s, units, {"precipitation": p2, "temperature": t2, "pet": e2}


def run_reference(hbpy, ref, **kw):
    import make_goldens as mg

    m, s, units, forcing = build(hbpy, **kw)
    q, rec = mg.run_ref_module(ref, s, units, forcing)
    rec = {k: v for k, v in rec.items() if not k.endswith(":internal_storage")}  # routing:gr4j's in-transit water (a process value)
    meta = {"structures": s.get_structure(), "units": units, "solver": m.solver, "labels": list(rec),
            "source": f"reference Python class GR4J + reference core (tools/ref_gr4j_case.py), {kw}"}
    return meta, forcing, np.asarray(q), rec


if __name__ == "__main__":
    import refpy
    import ref_melt_cases as rm

    hbpy = refpy.load_reference_python(backend="ref")
    ref = refpy.load_extension("ref")
    ref.init()
    for kw in (dict(), dict(n_units=3, n_steps=400, seed=5), dict(values=[("production_store", "capacity", 150.0),
                                                                         ("uh_input", "exchange_factor", -1.5),
                                                                         ("uh_input", "routing_capacity", 40.0),
                                                                         ("uh_input", "uh_base_time", 0.8)]),
               dict(model="GR6J"), dict(model="GR6J", n_units=3, n_steps=400, seed=5),
               dict(model="GR6J", values=[("production_store", "capacity", 150.0), ("uh_input", "exchange_factor", -1.2),
                                          ("uh_input", "routing_capacity", 40.0), ("uh_input", "uh_base_time", 0.8),
                                          ("uh_input", "exchange_threshold", 0.5), ("uh_input", "exp_store_coeff", 2.0)])):
        meta, forcing, q, rec = run_reference(hbpy, ref, **kw)
        om = rm.run_oracle(meta, forcing, list(rec))
        bad = [l for l in rec if not np.array_equal(om.records[l], rec[l], equal_nan=True)]
        print(meta["solver"], {k: v for k, v in kw.items() if k != "values"}, "outlet mismatches", int(np.sum(om.outlet != q)), "max abs",
              float(np.max(np.abs(om.outlet - q))), "labels differing", len(bad), bad[:4], "mean q", float(q.mean()))
