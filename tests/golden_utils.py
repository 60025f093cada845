"""Helpers shared by the parity tests: load tests/golden/*.npz bundles (made by tools/make_goldens.py from
the unmodified reference) and run the oracle restatement on them."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def bundles(prefix=""):
    return sorted(p.stem for p in GOLDEN.glob(f"{prefix}*.npz"))


def load(name):
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    forcing = {k[len("forcing_"):]: z[k] for k in z.files if k.startswith("forcing_")}
    records = {label: z[f"rec_{i}"] for i, label in enumerate(meta["labels"])}
    extra = {k: z[k] for k in z.files if k.startswith(("station_", "spatial_"))}
    for u in meta["units"]:
        u["land_covers"] = [tuple(x) for x in u["land_covers"]]
        if u.get("slope"):
            u["slope"] = tuple(u["slope"])
    return meta, forcing, z["outlet"], records, extra


def run_oracle(meta, forcing, labels=None):
    from oracle import hb_oracle

    n = len(next(iter(forcing.values())))
    om = hb_oracle.OracleModel(meta["structures"], meta["units"], meta["solver"], n)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    for label in (meta["labels"] if labels is None else labels):
        om.record(label)
    om.run()
    return om


def same(a, b):
    """Bit-exact including NaN placement."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def settings_from_structures(hb, meta, n_steps, solver=None):
    """Rebuild a SettingsModel on the host module `hb` (mirror or compiled) from the structure description a golden
    bundle carries, through the reference's select-then-add API, and give it the bundle's parameter values."""
    s = hb.SettingsModel()
    s.log_all(True)
    s.set_solver(solver or meta["solver"])
    start = np.datetime64("1981-01-01")
    s.set_timer("1981-01-01", str(start + np.timedelta64(n_steps - 1, "D")), 1, "day")
    for i, st in enumerate(meta["structures"]):
        if i > 0:
            s.add_structure()
        tr = next((x for x in st["hydro_unit_splitters"] if x["name"] == "snow_rain_transition"), None)
        if tr is None:  # SettingsModel::GeneratePrecipitationSplitters(withSnow = false): the rain splitter only
            s.generate_precipitation_splitters(False)
        else:
            s.generate_precipitation_splitters(True, tr["kind"])
        for b in st["hydro_unit_bricks"]:
            if b["is_land_cover"]:
                s.add_land_cover_brick(b["name"], b["kind"])
        snowpack_bricks = [b for b in st["hydro_unit_bricks"] if b["is_surface_component"]]
        holding = next((p["kind"] for b in snowpack_bricks for p in b["processes"] if p["name"] == "meltwater"), None)
        if tr is not None and holding:  # GenerateSnowpacksWithWaterRetention / AddSnowpackRefreezing (model_settings.py:248-255)
            rain_splitter = next(sp for sp in st["hydro_unit_splitters"] if sp["name"] == "rain_splitter")
            rain_to_snowpack = any(o["target"].endswith("_snowpack") for o in rain_splitter["outputs"])
            s.generate_snowpacks_with_water_retention(snowpack_bricks[0]["processes"][0]["kind"], holding, rain_to_snowpack)
            refreeze = next((p["kind"] for b in snowpack_bricks for p in b["processes"] if p["name"] == "refreeze"), None)
            if refreeze:
                s.add_snowpack_refreezing(refreeze)
        elif tr is not None:
            s.generate_snowpacks(next(b["processes"][0]["kind"] for b in st["hydro_unit_bricks"]
                                      if b["is_surface_component"]))
        extras = {}  # AddSnowIceTransformation / AddSnowpackSublimation, in the order of model_settings.py:258-263
        for b in st["hydro_unit_bricks"]:
            if b["is_surface_component"]:
                for proc in b["processes"][1:]:
                    extras[proc["name"]] = proc["kind"]
        if "snow_ice_transfo" in extras:
            s.add_snow_ice_transformation(extras["snow_ice_transfo"])
        if "snow_redistribution" in extras:  # on the glacier snowpacks too unless skipGlaciers (SettingsModel.cpp:565-567)
            glacier_snowpacks = [b for b in st["hydro_unit_bricks"] if b["is_surface_component"] and any(
                c["name"] == b["parent"] and c["kind"] == "glacier" for c in st["hydro_unit_bricks"])]
            skip = not any(p["name"] == "snow_redistribution" for b in glacier_snowpacks for p in b["processes"])
            s.add_snow_redistribution(extras["snow_redistribution"], skip)
        if "sublimation" in extras:
            s.add_snowpack_sublimation(extras["sublimation"])
        for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]:
            if b["is_surface_component"]:
                continue
            if b["is_land_cover"]:
                s.select_hydro_unit_brick(b["name"])
            elif b["attach"] == "sub_basin":
                s.add_sub_basin_brick(b["name"], b["kind"])
            else:
                s.add_hydro_unit_brick(b["name"], b["kind"])
            for p in b["parameters"]:
                s.add_brick_parameter(p["name"], p["value"])
            for proc in b["processes"]:
                s.add_brick_process(proc["name"], proc["kind"], proc["outputs"][0]["target"] if proc["outputs"] else "")
                if proc["outputs"] and proc["outputs"][0]["static"]:
                    s.set_process_outputs_as_static()
                if proc["outputs"] and proc["outputs"][0]["instantaneous"]:
                    s.set_process_outputs_as_instantaneous()
        s.add_logging_to("outlet")
    for st in meta["structures"]:
        for el in st["hydro_unit_bricks"] + st["sub_basin_bricks"] + st["hydro_unit_splitters"]:
            for p in el.get("parameters", []):
                if p["name"] not in ("infinite_storage", "no_melt_when_snow_cover"):
                    assert s.set_parameter_value(el["name"], p["name"], p["value"])
            for proc in el.get("processes", []):
                for p in proc["parameters"]:
                    assert s.set_parameter_value(el["name"], p["name"], p["value"])
    return s
