"""GSM-Socont with dated actions (glacier delta-h evolution, land-cover change, annual snow->ice) run through the
UNMODIFIED reference core (oracle/_ref) and through the oracle restatement; used to pin the restatement and to
write tests/golden/actions_*.npz. Build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from hydrobricks_b200 import models, schedule  # noqa: E402


def synthetic_case(n_units=10, n_years=6, seed=11):
    rng = np.random.default_rng(seed)
    elev = np.linspace(1900, 3300, n_units)
    units = []
    for i in range(n_units):
        g = float(np.clip((elev[i] - 2300) / 900 * 0.85, 0, 0.85))
        g = 0.0 if g < 0.05 else round(g, 3)
        units.append({"id": i + 1, "area": float(round(rng.uniform(0.6e6, 1.4e6))), "elevation": float(elev[i]),
                      "slope": (float(round(rng.uniform(8, 30), 2)), "deg"),
                      "land_covers": [("ground", 1 - g), ("glacier", g)]})
    units = sorted(units, key=lambda u: -u["elevation"])
    T = 365 * n_years + 1
    t = np.arange(T)
    doy = np.fmod(t, 365.25) / 365.25
    temp = 3 - 11 * np.cos(2 * np.pi * doy) + rng.normal(0, 2.5, T)
    precip = np.where(rng.random(T) < 0.45, -7 * np.log(1 - rng.random(T)), 0.0)
    pet = np.maximum(0.0, 1.8 - 1.8 * np.cos(2 * np.pi * doy))
    z = np.array([u["elevation"] for u in units])
    forcing = {"precipitation": np.maximum(precip[:, None] * (1 + 0.04 * (z - 2500) / 100), 0),
               "temperature": temp[:, None] - 0.6 * (z - 2500) / 100,
               "pet": np.repeat(pet[:, None], n_units, axis=1)}
    # delta-h lookup tables over the glacier units: 11 rows, lower units vanish first; thin ice so that the
    # glacier actually retreats within a few years
    gl = [i for i, u in enumerate(units) if dict(u["land_covers"])["glacier"] > 0]
    a0 = np.array([units[i]["area"] * dict(units[i]["land_covers"])["glacier"] for i in gl])
    rows = 11
    order = np.argsort([units[i]["elevation"] for i in gl])  # lowest first
    rank = np.empty(len(gl))
    rank[order] = np.arange(len(gl)) / max(len(gl) - 1, 1)
    areas = np.zeros((rows, len(gl)))
    volumes = np.zeros((rows, len(gl)))
    thick0 = 4.0 + 14.0 * rank  # m of ice
    for r in range(rows):
        x = r / (rows - 1)
        keep = np.clip((1 + 0.6 * rank) * (1 - x) - 0.15 * (1 - rank) * x * 3, 0, 1)
        if r == rows - 1:
            keep[:] = 0
        areas[r] = np.round(a0 * keep, 3)
        volumes[r] = areas[r] * thick0 * (1 - 0.7 * x)
    areas[0] = a0
    volumes[0] = a0 * thick0
    return units, forcing, T, gl, areas, volumes


def build_settings(backend, solver, **options):
    """options: further models.Socont options, e.g. the snowpack liquid-water retention of tools/ref_retention_case.py."""
    m = models.Socont(solver=solver, soil_storage_nb=2, surface_runoff="socont_runoff", record_all=True,
                      land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
                      glacier_infinite_storage=False, backend=backend,
                      **{k: v for k, v in options.items() if not k.startswith("_")})
    for comp, name, v in [("type:snowpack", "degree_day_factor", 4.0), ("slow_reservoir", "capacity", 150),
                          ("slow_reservoir", "response_factor", 0.03), ("surface_runoff", "beta", 900),
                          ("glacier", "degree_day_factor", 8.5),
                          ("glacier_area_rain_snowmelt_storage", "response_factor", 0.25),
                          ("glacier_area_icemelt_storage", "response_factor", 0.4),
                          ("slow_reservoir", "percolation_rate", 0.6), ("slow_reservoir_2", "response_factor", 0.008)]:
        m.settings.set_parameter_value(comp, name, v)
    return m


START_MJD = 44605.0  # 1981-01-01
CHANGES = [  # (days after start, unit position in basin order, new glacier area as a share of its initial area)
    (400, 1, 0.90), (800, 2, 0.75), (800.0, 0, 1.05), (1300, 3, 0.0), (1700, 1, 0.60)]


def run_reference(ref, solver, with_delta_h=True, with_changes=True, with_snow_to_ice=True, area_scaling=False,
                  options=None):
    units, forcing, T, gl, areas, volumes = synthetic_case()
    m = build_settings(ref, solver, **(options or {}))
    for comp, name, v in (options or {}).get("_parameters", []):
        assert m.settings.set_parameter_value(comp, name, v), (comp, name)
    end = str(np.datetime64("1981-01-01") + np.timedelta64(T - 1, "D"))
    m.setup(units, "1981-01-01", end)
    model = m.model
    keep = []
    if with_snow_to_ice:
        a = ref.ActionGlacierSnowToIceTransformation(9, 30, "glacier")
        assert model.add_action(a)
        keep.append(a)
    if with_delta_h:
        # area_scaling: ActionGlacierEvolutionAreaScaling, same tables, every unit follows its own column
        a = ref.ActionGlacierEvolutionAreaScaling() if area_scaling else ref.ActionGlacierEvolutionDeltaH()
        a.add_lookup_tables(10, "glacier", np.array([m.units[i]["id"] for i in gl], dtype=np.int32), areas, volumes)
        assert model.add_action(a)
        keep.append(a)
    if with_changes:
        a = ref.ActionLandCoverChange()
        for days, pos, share in CHANGES:
            iu = gl[pos]
            a.add_change(START_MJD + days, m.units[iu]["id"], "glacier", share * areas[0][pos])
        assert model.add_action(a)
        keep.append(a)
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
            "actions": {"delta_h": {"month": 10, "cover": "glacier", "unit_positions": gl,
                                    "area_scaling": bool(area_scaling)} if with_delta_h else None,
                        "changes": [[d, gl[p], s * areas[0][p] / m.units[gl[p]]["area"]] for d, p, s in CHANGES]
                        if with_changes else [],
                        "snow_to_ice": [9, 30] if with_snow_to_ice else None, "start_mjd": START_MJD},
            "source": "reference core driven directly (tools/ref_actions_case.py)"}
    if options:  # what build_settings needs to rebuild this model on another host
        meta["model_options"] = {k: v for k, v in options.items() if not k.startswith("_")}
        meta["model_parameters"] = [list(x) for x in options.get("_parameters", [])]
    return meta, forcing, model.get_outlet_discharge(), rec, frac, areas, volumes


def run_oracle(meta, forcing, areas, volumes, labels):
    from oracle import hb_oracle

    T = len(next(iter(forcing.values())))
    units = meta["units"]
    init = {}
    act = meta["actions"]
    if act["delta_h"]:
        # ActionGlacierEvolutionDeltaH::Init: initial ice w.e. of the table's units (…DeltaH.cpp:21-70)
        for j, iu in enumerate(act["delta_h"]["unit_positions"]):
            init[("glacier", "ice", iu)] = volumes[0][j] * 900.0 / areas[0][j]
    om = hb_oracle.OracleModel(meta["structures"], units, meta["solver"], T, initial_states=init)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    s0 = act["start_mjd"]
    # recursive actions in the order they were added, then the sporadic ones (ActionsManager.cpp:80-104)
    events = []
    if act["snow_to_ice"]:
        for k in schedule.recursive_steps(s0, T, *act["snow_to_ice"]):
            events.append((k, 0, "s2i", None))
    if act["delta_h"]:
        for k in schedule.recursive_steps(s0, T, act["delta_h"]["month"], 1):
            events.append((k, 1, "dh", None))
    for n, (days, iu, frac) in enumerate(sorted(act["changes"], key=lambda c: c[0])):
        events.append((schedule.sporadic_step(s0, s0 + days), 100 + n, "lc", (iu, schedule.check_fraction(frac))))
    events.sort(key=lambda e: (e[0], e[1]))
    dh_steps = [e[0] for e in events if e[2] == "dh"]
    # the oracle keeps the events in insertion order for equal steps: insert in global order
    first_dh = True
    for k, _, kind, payload in events:
        if kind == "s2i":
            om.add_snow_to_ice("glacier", [k])
        elif kind == "dh":
            if first_dh:
                om.set_delta_h("glacier", act["delta_h"]["unit_positions"], areas, volumes, [k],
                               area_scaling=act["delta_h"].get("area_scaling", False))
                first_dh = False
            else:
                om.L.hbo_add_event(om.h, int(k), 3 if act["delta_h"].get("area_scaling") else 2, -1, 0.0)
        else:
            om.add_land_cover_change(k, payload[0], "glacier", payload[1])
    for label in labels:
        om.record(label)
    om.run()
    return om


if __name__ == "__main__":
    import refpy

    ref = refpy.load_extension("ref")
    ref.init()
    for solver in ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson"]:
        for flags in [(True, True, True), (True, False, False), (False, True, False), (False, False, True),
                      (True, False, False, True), (True, True, True, True)]:
            meta, forcing, q, rec, frac, areas, volumes = run_reference(ref, solver, *flags)
            om = run_oracle(meta, forcing, areas, volumes, list(rec))
            bad = int(np.sum(om.outlet != q))
            worst = 0.0
            nbad = 0
            for label in rec:
                a, b = om.records[label], rec[label]
                if not np.array_equal(a, b, equal_nan=True):
                    nbad += 1
                    worst = max(worst, float(np.nanmax(np.abs(a - b))))
            gfr = frac.get("glacier")
            print(solver, flags, "outlet mismatches", bad, "labels differing", nbad, "worst", worst,
                  "glacier fraction range", None if gfr is None else (float(np.nanmin(gfr)), float(np.nanmax(gfr))),
                  "final ice", float(np.nansum(rec["glacier:ice_content"][-1])))
