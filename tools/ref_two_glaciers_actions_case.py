"""Dated actions on a basin with TWO glacier covers per unit (glacier_ice + glacier_debris): delta-h evolution of the
clean ice, dated land-cover edits of the debris-covered part, the annual snow -> ice transfer on both — through the
UNMODIFIED reference core (oracle/_ref). Writes tests/golden/actions2_*.npz (tools/make_goldens.py two_glaciers_actions).
Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

START_MJD = 44605.0  # 1981-01-01
COVERS = ["ground", "glacier_ice", "glacier_debris"]


def case(n_units=10, n_years=5, seed=23):
    import ref_actions_case as ra

    units, forcing, T, gl, areas, volumes = ra.synthetic_case(n_units=n_units, n_years=n_years, seed=seed)
    rng = np.random.default_rng(seed)
    debris = {}
    for i, u in enumerate(units):
        g = dict(u["land_covers"])["glacier"]
        d = round(float(g * rng.uniform(0.15, 0.4)), 4) if g > 0 else 0.0
        debris[i] = d
        u["land_covers"] = [("ground", 1.0 - g), ("glacier_ice", g - d), ("glacier_debris", d)]
    # the delta-h tables follow the clean ice: scale the synthetic tables to its initial areas
    a0 = np.array([units[i]["area"] * dict(units[i]["land_covers"])["glacier_ice"] for i in gl])
    scale = a0 / areas[0]
    areas = np.round(areas * scale, 3)
    areas[0] = a0
    volumes = volumes * scale * 2.5  # thicker ice: the retreat passes several table rows without emptying the glacier
    # dated edits of the debris-covered part: (days after start, position in gl, share of its initial area)
    changes = [(380, 0, 0.8), (900, 1, 1.15), (900.0, 2, 0.5), (1400, 0, 0.0)]
    return units, forcing, T, gl, areas, volumes, changes


def build_settings(backend, solver):
    from hydrobricks_b200 import models

    m = models.Socont(solver=solver, soil_storage_nb=1, surface_runoff="socont_runoff", record_all=True,
                      land_cover_names=COVERS, land_cover_types=["ground", "glacier", "glacier"],
                      glacier_infinite_storage=False, backend=backend)
    for comp, name, v in [("type:snowpack", "degree_day_factor", 4.0), ("slow_reservoir", "capacity", 150),
                          ("slow_reservoir", "response_factor", 0.03), ("surface_runoff", "beta", 900),
                          ("glacier_ice", "degree_day_factor", 8.5), ("glacier_debris", "degree_day_factor", 5.5),
                          ("glacier_area_rain_snowmelt_storage", "response_factor", 0.25),
                          ("glacier_area_icemelt_storage", "response_factor", 0.4)]:
        assert m.settings.set_parameter_value(comp, name, v)
    return m


def add_actions(hb, m, gl, areas, volumes, changes):
    """On a set-up model `m` of host module `hb` (the reference or one of this repo's hosts); returns the actions (the
    model keeps non-owning pointers)."""
    keep = []
    for cover in ("glacier_ice", "glacier_debris"):
        a = hb.ActionGlacierSnowToIceTransformation(9, 30, cover)
        assert m.model.add_action(a)
        keep.append(a)
    dh = hb.ActionGlacierEvolutionDeltaH()
    dh.add_lookup_tables(10, "glacier_ice", np.array([m.units[i]["id"] for i in gl], dtype=np.int32), areas, volumes)
    assert m.model.add_action(dh)
    keep.append(dh)
    lc = hb.ActionLandCoverChange()
    for days, pos, share in changes:
        u = m.units[gl[pos]]
        lc.add_change(START_MJD + days, u["id"], "glacier_debris", share * dict(u["land_covers"])["glacier_debris"] * u["area"])
    assert m.model.add_action(lc)
    keep.append(lc)
    return keep


def run_reference(ref, solver):
    units, forcing, T, gl, areas, volumes, changes = case()
    m = build_settings(ref, solver)
    end = str(np.datetime64("1981-01-01") + np.timedelta64(T - 1, "D"))
    m.setup(units, "1981-01-01", end)
    model = m.model
    keep = add_actions(ref, m, gl, areas, volumes, changes)  # noqa: F841
    time = np.arange(T, dtype=np.float64) + START_MJD
    ids = np.array([u["id"] for u in m.units], dtype=np.int32)
    for k, v in forcing.items():
        assert model.create_time_series(k, time, ids, np.ascontiguousarray(v))
    assert model.attach_time_series_to_hydro_units()
    model.reset()
    model.run()
    rec = {label: model.get_hydro_unit_values(label) for label in model.get_recorded_hydro_unit_labels()}
    frac = {label: model.get_hydro_unit_fractions(label) for label in model.get_recorded_hydro_unit_fraction_labels()}
    meta = {"structures": m.settings.get_structure(), "units": m.units, "solver": solver, "labels": list(rec),
            "actions2": {"gl": gl, "changes": [list(c) for c in changes]},
            "source": "reference core driven directly (tools/ref_two_glaciers_actions_case.py)"}
    return meta, forcing, np.asarray(model.get_outlet_discharge()), rec, frac, areas, volumes


if __name__ == "__main__":
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ("rk4", "heun_explicit"):
        meta, forcing, q, rec, frac, areas, volumes = run_reference(ref, solver)
        print(solver, "mean q", float(q.mean()), "labels", len(rec),
              "ice fraction range", float(np.nanmin(frac["glacier_ice"])), float(np.nanmax(frac["glacier_ice"])),
              "debris fraction range", float(np.nanmin(frac["glacier_debris"])), float(np.nanmax(frac["glacier_debris"])),
              "final ice", float(np.nansum(rec["glacier_ice:ice_content"][-1])), float(np.nansum(rec["glacier_debris:ice_content"][-1])),
              "distinct ice fractions of the top unit", len(set(np.round(frac["glacier_ice"][:, 0], 9))))
