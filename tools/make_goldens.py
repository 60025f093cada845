"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref + the reference Python layer).

Run in the build container only:   python tools/make_goldens.py
Each bundle holds: meta (JSON: structures as exported by SettingsModel.get_structure() with parameter
values, basin-ordered units, solver, labels), the forcing handed to ModelHydro::CreateTimeSeries
[T x U], the outlet discharge [T] and every recorded hydro-unit series [T x U] of the reference run.
The committed bundles are what tests/ compare the oracle restatement and the CUDA engine against on
machines where /root/reference does not exist.
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ref_cases  # noqa: E402
import refpy  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def save(name, meta, forcing, outlet, records, extra=None):
    (OUT / name).parent.mkdir(parents=True, exist_ok=True)
    arrays = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), "outlet": np.asarray(outlet)}
    for k, v in forcing.items():
        arrays["forcing_" + k] = np.asarray(v, dtype=np.float64)
    for i, label in enumerate(meta["labels"]):
        arrays[f"rec_{i}"] = np.asarray(records[label])
    for k, v in (extra or {}).items():
        arrays[k] = np.asarray(v)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    print(f"{name}: T={len(outlet)} labels={len(meta['labels'])} "
          f"{(OUT / (name + '.npz')).stat().st_size / 1024:.0f} KiB")


def run_ref_module(ref, settings, units, forcing, labels_from_model=True):
    """Drive the reference core directly (no Python layer): init_with_basin, time series, run."""
    basin = ref.SettingsBasin()
    for u in units:
        basin.add_hydro_unit(u["id"], u["area"], u.get("elevation", -9999))
        if u.get("slope"):
            basin.add_hydro_unit_property_double("slope", u["slope"][0], u["slope"][1])
        if u.get("aspect_class"):
            basin.add_hydro_unit_property_str("aspect_class", u["aspect_class"])
        for name, frac in u["land_covers"]:
            basin.add_land_cover(name, "", frac)
    for u in units:  # lateral connections (hydro_units.py set_connectivity): (receiver position, fraction)
        for receiver, fraction in u.get("lateral", []):
            basin.add_lateral_connection(u["id"], units[receiver]["id"], fraction, "snow_slide")
    model = ref.ModelHydro()
    model.init_with_basin(settings, basin)
    n = len(next(iter(forcing.values())))
    t = settings.get_timer()
    import pandas as pd

    mjd0 = (pd.Timestamp(t["start"]) - pd.Timestamp("1858-11-17")).days
    time = np.arange(n, dtype=np.float64) + mjd0
    ids = np.array([u["id"] for u in units], dtype=np.int32)
    for k, v in forcing.items():
        assert model.create_time_series(k, time, ids, np.asarray(v, dtype=np.float64))
    assert model.attach_time_series_to_hydro_units()
    model.run()
    rec = {label: model.get_hydro_unit_values(label) for label in model.get_recorded_hydro_unit_labels()}
    return model.get_outlet_discharge(), rec


def kat_linear_storage(ref, solver):
    """core/tests/src/SolverTest.cpp:117-244 (fixture SolverLinearStorage)."""
    s = ref.SettingsModel()
    s.set_solver(solver)
    s.set_timer("2020-01-01", "2020-01-20", 1, "day")
    s.add_hydro_unit_brick("storage", "storage")
    s.add_brick_forcing("precipitation")
    s.add_brick_logging("water_content")
    s.add_brick_process("outflow", "outflow:linear")
    s.set_process_parameter_value("response_factor", 0.3)
    s.add_process_logging("output")
    s.add_process_output("outlet")
    s.add_logging_to("outlet")
    units = [{"id": 1, "area": 100.0, "elevation": -9999, "slope": None, "land_covers": []}]
    precip = np.array([0.0, 10.0, 10.0, 10.0] + [0.0] * 16).reshape(20, 1)
    forcing = {"precipitation": precip}
    q, rec = run_ref_module(ref, s, units, forcing)
    meta = {"structures": s.get_structure(), "units": units, "solver": solver, "labels": list(rec),
            "source": "core/tests/src/SolverTest.cpp:117-244"}
    return meta, forcing, q, rec


def kat_socont_quick(ref):
    """core/tests/src/ModelSocontTest.cpp:536-626 (QuickDischargeIsCorrect) with the structure of
    core/tests/src/helpers.cpp:9-107 (GenerateStructureSocont, 1 soil store, socont_runoff)."""
    s = ref.SettingsModel()
    s.log_all(True)
    s.set_solver("runge_kutta")
    s.set_timer("2020-01-01", "2020-01-10", 1, "day")
    s.generate_precipitation_splitters(True)
    s.add_land_cover_brick("ground", "generic_land_cover")
    s.generate_snowpacks("melt:degree_day")
    s.add_sub_basin_brick("glacier_area_rain_snowmelt_storage", "storage")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_sub_basin_brick("glacier_area_icemelt_storage", "storage")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.select_hydro_unit_brick("ground")
    s.add_brick_process("infiltration", "infiltration:socont", "slow_reservoir")
    s.add_brick_process("runoff", "outflow:rest", "surface_runoff")
    s.set_process_outputs_as_static()
    s.add_hydro_unit_brick("slow_reservoir", "storage")
    s.add_brick_parameter("capacity", 200.0)
    s.add_brick_process("et", "et:socont")
    s.add_brick_process("outflow", "outflow:linear", "outlet")
    s.add_brick_process("overflow", "overflow", "outlet")
    s.add_hydro_unit_brick("surface_runoff", "storage")
    s.add_brick_process("runoff", "runoff:socont", "outlet")
    s.add_logging_to("outlet")
    s.set_parameter_value("surface_runoff", "beta", 301)
    s.set_parameter_value("slow_reservoir", "capacity", 0)
    units = [{"id": 1, "area": 38.9 * 1000000, "elevation": -9999, "slope": (0.3, "m/m"),
              "land_covers": [("ground", 1.0)]}]
    forcing = {
        "precipitation": np.array([0.0, 13.8, 59.3, 34.2, 13.7, 26.1, 9.8, 0.0, 0.0, 0.0]).reshape(10, 1),
        "temperature": np.full((10, 1), 5.0),
        "pet": np.zeros((10, 1)),
    }
    q, rec = run_ref_module(ref, s, units, forcing)
    meta = {"structures": s.get_structure(), "units": units, "solver": "rk4", "labels": list(rec),
            "source": "core/tests/src/ModelSocontTest.cpp:536-626"}
    return meta, forcing, q, rec


def snow_ice_bundles(hb):
    """continuous snow -> ice transformation on the glacier snowpack (transform:snow_ice_swat /
    transform:snow_ice_constant, ProcessTransformSnowToIce{Swat,Constant}.cpp), finite glacier ice."""
    cases = [("rk4", "swat", 2, {"snow_ice_basal_acc_coeff": 0.004}),
             ("rk4", "constant", 1, {"snow_ice_rate": 0.8}),
             ("heun_explicit", "swat", 1, {"snow_ice_basal_acc_coeff": 0.004}),
             ("euler_explicit", "constant", 2, {"snow_ice_rate": 0.8})]
    for solver, algo, nb, pr in cases:
        m, hu, f, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-12-31", n_units=10, soil_storage_nb=nb,
                                           infinite_storage=False, params=pr,
                                           snow_ice_transformation="transform:snow_ice_" + algo)
        structures, units, fdata, q, rec = ref_cases.collect(m, hu, f)
        meta = {"structures": structures, "units": units, "solver": m.solver, "labels": list(rec),
                "source": "reference Python layer, 10 synthetic bands with slope/glacier, Rhone-Gletsch meteo "
                          "1981-1982, snow_ice_transformation=transform:snow_ice_" + algo}
        save(f"gsm_socont_{solver}_snowice_{algo}_{nb}store", meta, fdata, q, rec)
    # snow sublimation (sublimation:constant / sublimation:pet, ProcessSublimation{Constant,PET}.cpp) on every snowpack
    cases = [("rk4", "constant", 1, dict(with_glacier=False), {"sublimation_rate": 0.6}),
             ("rk4", "pet", 2, dict(infinite_storage=False, snow_ice_transformation="transform:snow_ice_swat"),
              {"sublimation_pet_factor": 0.5, "snow_ice_basal_acc_coeff": 0.004}),
             ("heun_explicit", "constant", 1, dict(infinite_storage=False), {"sublimation_rate": 0.6}),
             ("euler_explicit", "pet", 2, dict(with_glacier=False), {"sublimation_pet_factor": 0.5})]
    for solver, algo, nb, kw, pr in cases:
        m, hu, f, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-12-31", n_units=10, soil_storage_nb=nb,
                                           params=pr, snow_sublimation_process="sublimation:" + algo, **kw)
        structures, units, fdata, q, rec = ref_cases.collect(m, hu, f)
        meta = {"structures": structures, "units": units, "solver": m.solver, "labels": list(rec),
                "source": "reference Python layer, 10 synthetic bands, Rhone-Gletsch meteo 1981-1982, "
                          "snow_sublimation_process=sublimation:" + algo + " " + json.dumps(kw)}
        tag = ("glacier" if kw.get("with_glacier", True) else "noglacier") + ("_snowice" if "snow_ice_transformation" in kw else "")
        save(f"gsm_socont_{solver}_sublim_{algo}_{tag}_{nb}store", meta, fdata, q, rec)


def sequential_bundles(hb, ref):
    """the sequential solver family (base/SolverSequential.cpp; crank_nicolson is the default of the
    reference's Python layer, models/model.py:80)."""
    for solver in ["crank_nicolson", "implicit_euler", "exponential_euler", "analytic_linear"]:
        save(f"kat_linear_storage_{solver}", *kat_linear_storage(ref, solver))
    cases = [("crank_nicolson", "glacier_2store", dict(soil_storage_nb=2)),
             ("crank_nicolson", "glacier_finite_1store", dict(soil_storage_nb=1, infinite_storage=False)),
             ("crank_nicolson", "glacier_linear_1store", dict(soil_storage_nb=1, surface_runoff="linear_storage")),
             ("implicit_euler", "glacier_2store", dict(soil_storage_nb=2)),
             ("exponential_euler", "glacier_finite_1store", dict(soil_storage_nb=1, infinite_storage=False)),
             ("exponential_euler", "noglacier_2store", dict(soil_storage_nb=2, with_glacier=False)),
             ("analytic_linear", "glacier_2store", dict(soil_storage_nb=2)),
             ("analytic_linear", "glacier_linear_1store", dict(soil_storage_nb=1, surface_runoff="linear_storage"))]
    for solver, tag, kw in cases:
        m, hu, f, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-12-31", n_units=10, **kw)
        structures, units, fdata, q, rec = ref_cases.collect(m, hu, f)
        meta = {"structures": structures, "units": units, "solver": m.solver, "labels": list(rec),
                "source": "reference Python layer, 10 synthetic bands with slope/glacier, Rhone-Gletsch meteo "
                          "1981-1982, sequential solver " + solver}
        save(f"gsm_socont_{solver}_{tag}", meta, fdata, q, rec)


def action_bundles(ref, only_new=False):
    """Dated actions (SURVEY §8 a21, config C4 in small): delta-h glacier evolution, land-cover changes, snow->ice;
    the area-scaling glacier evolution (ActionGlacierEvolutionAreaScaling) and actions under a sequential solver."""
    import ref_actions_case as rc

    cases = [("rk4", (True, True, True), False), ("rk4", (True, False, False), False), ("rk4", (False, True, False), False),
             ("heun_explicit", (True, True, True), False), ("euler_explicit", (True, True, True), False)]
    new = [("rk4", (True, False, False), True), ("heun_explicit", (True, True, True), True),
           ("crank_nicolson", (True, True, True), False)]
    for solver, flags, area_scaling in ([] if only_new is None else new if only_new else cases + new):
        meta, forcing, q, rec, frac, areas, volumes = rc.run_reference(ref, solver, *flags, area_scaling=area_scaling)
        tag = "".join(n for n, f in zip(("as" if area_scaling else "dh", "lc", "s2i"), flags) if f)
        extra = {"spatial_dh_areas": areas, "spatial_dh_volumes": volumes}
        for k, v in frac.items():
            extra["spatial_fraction_" + k] = v
        keep = ["glacier:ice_content", "glacier:water_content", "ground:water_content",
                "glacier_snowpack:snow_content", "ground_snowpack:snow_content", "slow_reservoir:water_content",
                "surface_runoff:water_content", "glacier:melt:output", "glacier:outflow_rain_snowmelt:output"]
        meta["labels"] = keep
        save(f"actions_{solver}_{tag}", meta, forcing, q, {k: rec[k] for k in keep}, extra=extra)
    # the same actions with liquid water in the snowpacks: the land-cover transfer moves it like the snow
    retention = dict(snow_water_retention_process="outflow:snow_holding", snow_refreezing_process="refreeze:degree_day",
                     rain_to_snowpack=True, _parameters=[("type:snowpack", "water_holding_capacity", 0.15),
                                                         ("type:snowpack", "refreezing_factor", 0.06)])
    meta, forcing, q, rec, frac, areas, volumes = rc.run_reference(ref, "rk4", True, True, True, options=retention)
    extra = {"spatial_dh_areas": areas, "spatial_dh_volumes": volumes}
    for k, v in frac.items():
        extra["spatial_fraction_" + k] = v
    keep = ["glacier:ice_content", "glacier:water_content", "ground:water_content", "glacier_snowpack:snow_content",
            "ground_snowpack:snow_content", "glacier_snowpack:water_content", "ground_snowpack:water_content",
            "ground_snowpack:meltwater:output", "glacier_snowpack:refreeze:output", "slow_reservoir:water_content"]
    meta["labels"] = keep
    save("actions_rk4_dhlcs2i_retention", meta, forcing, q, {k: rec[k] for k in keep}, extra=extra)


def melt_variant_bundles(ref):
    """snow_melt_process variants melt:degree_day_aspect and melt:temperature_index (in the engine since round 2)."""
    import ref_melt_cases as rm

    for kind, tag in (("melt:degree_day_aspect", "aspect"), ("melt:temperature_index", "tindex")):
        for solver, kw, name in (("rk4", dict(glacier=True, nsoil=2), "glacier_2store"),
                                 ("crank_nicolson", dict(glacier=False, nsoil=1), "noglacier_1store")):
            meta, forcing, q, rec = rm.run_reference(ref, kind, solver, **kw)
            save(f"melt_{tag}_{solver}_{name}", meta, forcing, q, rec)


def snow_slide_bundles(ref):
    """transport:snow_slide: lateral snow fluxes between the units inside a step (the engine's sequential-over-units
    kernel)."""
    import ref_snow_slide_case as rs

    for solver, kw, name in (("rk4", dict(glacier=True, nsoil=1), "glacier_1store"),
                             ("crank_nicolson", dict(glacier=False, nsoil=2), "noglacier_2store")):
        meta, forcing, q, rec = rs.run_reference(ref, solver, **kw)
        save(f"snow_slide_{solver}_{name}", meta, forcing, q, rec)


def threshold_splitter_bundles(ref):
    """snow_rain:threshold instead of the linear transition."""
    import ref_threshold_case as rt

    for solver, kw, name in (("rk4", dict(glacier=True, nsoil=1), "glacier_1store"),
                             ("heun_explicit", dict(glacier=False, nsoil=2), "noglacier_2store")):
        meta, forcing, q, rec = rt.run_reference(ref, solver, **kw)
        save(f"splitter_threshold_{solver}_{name}", meta, forcing, q, rec)


def rain_only_bundles(ref):
    """The rain splitter of GeneratePrecipitationSplitters(withSnow = false): no snow routine at all."""
    import ref_rain_case as rr

    for solver, kw, name in (("rk4", dict(nsoil=2), "2store"), ("crank_nicolson", dict(nsoil=1), "1store")):
        meta, forcing, q, rec = rr.run_reference(ref, solver, **kw)
        save(f"splitter_rain_{solver}_nosnow_{name}", meta, forcing, q, rec)


def snow_retention_bundles(ref):
    """Snowpack liquid-water retention (outflow:snow_holding), refreezing (refreeze:degree_day) and the rain routed through
    the snowpack (tools/ref_retention_case.py)."""
    import ref_retention_case as rr

    for solver, kw, name in (("rk4", dict(glacier=True, nsoil=2), "refreeze_glacier_2store"),
                             ("rk4", dict(glacier=False, nsoil=1, refreeze=False), "holding_1store"),
                             ("heun_explicit", dict(glacier=True, nsoil=1, rain_to_snowpack=True), "rain_glacier_1store"),
                             ("euler_explicit", dict(glacier=False, nsoil=2, rain_to_snowpack=True), "rain_2store"),
                             ("crank_nicolson", dict(glacier=True, nsoil=1), "refreeze_glacier_1store"),
                             ("rk4", dict(glacier=True, nsoil=1, rain_to_snowpack=True,
                                          extras=dict(snow_ice_transformation="transform:snow_ice_constant",
                                                      snow_sublimation_process="sublimation:pet")), "rain_snowice_sublim_1store")):
        meta, forcing, q, rec = rr.run_reference(ref, solver, **kw)
        save(f"snow_retention_{solver}_{name}", meta, forcing, q, rec)


def hbv_bundles(ref):
    """HBV-96 through the reference's own Python class (tools/ref_hbv_case.py): ORACLE ONLY — the engine does not run this
    family; the vectors go to tests/golden/next (outside the engine's parity lists)."""
    import ref_hbv_case as rh

    hbpy = refpy.load_reference_python(backend="ref")
    for solver in ("rk4", "crank_nicolson", "euler_explicit"):
        meta, forcing, q, rec = rh.run_reference(hbpy, ref, solver, n_units=4, n_steps=600, seed=9)
        save(f"next/hbv96_{solver}_open", meta, forcing, q, rec)


def gr_bundles(ref):
    """GR4J / GR6J through the reference's own Python classes (tools/ref_gr4j_case.py): ORACLE ONLY, tests/golden/next."""
    import ref_gr4j_case as rg

    hbpy = refpy.load_reference_python(backend="ref")
    for model in ("GR4J", "GR6J"):
        meta, forcing, q, rec = rg.run_reference(hbpy, ref, model=model, n_units=4, n_steps=600, seed=9)
        save(f"next/{model.lower()}_open", meta, forcing, q, rec)


def two_glaciers_bundles(ref):
    """glacier_ice + glacier_debris in one unit (the engine's twin units, hbk_set_unit_twins)."""
    import ref_two_glaciers_case as rt

    for solver, kw, name in (("rk4", dict(nsoil=1), "finite_1store"), ("rk4", dict(nsoil=2, infinite=True), "infinite_2store"),
                             ("crank_nicolson", dict(nsoil=1, transfo="transform:snow_ice_constant"), "snowice_1store"),
                             ("heun_explicit", dict(nsoil=1, melt="melt:degree_day_aspect"), "aspect_1store")):
        meta, forcing, q, rec = rt.run_reference(ref, solver, **kw)
        save(f"two_glaciers_{solver}_{name}", meta, forcing, q, rec)


def two_glaciers_action_bundles(ref):
    """Dated actions on glacier_ice + glacier_debris units: delta-h on the clean ice, land-cover edits of the debris cover,
    snow -> ice on both (tools/ref_two_glaciers_actions_case.py)."""
    import ref_two_glaciers_actions_case as rt

    for solver in ("rk4", "heun_explicit"):
        meta, forcing, q, rec, frac, areas, volumes = rt.run_reference(ref, solver)
        extra = {"spatial_dh_areas": areas, "spatial_dh_volumes": volumes}
        for k, v in frac.items():
            extra["spatial_fraction_" + k] = v
        keep = ["glacier_ice:ice_content", "glacier_debris:ice_content", "glacier_ice:water_content", "ground:water_content",
                "glacier_ice_snowpack:snow_content", "glacier_debris_snowpack:snow_content", "ground_snowpack:snow_content",
                "slow_reservoir:water_content", "glacier_ice:melt:output", "glacier_debris:melt:output",
                "glacier_debris:outflow_rain_snowmelt:output"]
        meta["labels"] = keep
        save(f"actions2_{solver}_two_glaciers", meta, forcing, q, {k: rec[k] for k in keep}, extra=extra)


def slow_order_bundles(ref):
    """hbk_structure.slow_process_order = 1 (the process order of core/tests/src/helpers.cpp), two soil stores."""
    import ref_slow_order_case as ro

    for solver in ("rk4", "crank_nicolson"):
        meta, forcing, q, rec = ro.run_reference(ref, solver)
        save(f"slow_order_{solver}_noglacier_2store", meta, forcing, q, rec)


def main():
    hb = refpy.load_reference_python("ref")
    ref = sys.modules["hydrobricks._hydrobricks"]
    if len(sys.argv) > 1 and sys.argv[1] == "snow_ice":
        snow_ice_bundles(hb)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "melt":
        melt_variant_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "slow_order":
        slow_order_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "threshold":
        threshold_splitter_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "two_glaciers_actions":
        two_glaciers_action_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "two_glaciers":
        two_glaciers_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rain_only":
        rain_only_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "hbv":
        hbv_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gr":
        gr_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "snow_retention":
        snow_retention_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "snow_slide":
        snow_slide_bundles(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "actions_retention":
        action_bundles(ref, only_new=None)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "actions_new":
        action_bundles(ref, only_new=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sequential":
        sequential_bundles(hb, ref)
        return

    for solver in ["euler_explicit", "heun_explicit", "runge_kutta"]:
        save(f"kat_linear_storage_{solver}", *kat_linear_storage(ref, solver))
    save("kat_socont_quick_rk4", *kat_socont_quick(ref))

    def bundle(name, model, hu, forcing, n_years_note):
        structures, units, fdata, q, rec = ref_cases.collect(model, hu, forcing)
        meta = {"structures": structures, "units": units, "solver": model.solver, "labels": list(rec),
                "source": n_years_note}
        save(name, meta, fdata, q, rec)

    # RK4 (the headline solver) gets every structure variant; Heun/Euler a representative subset, to keep
    # the committed fixtures small.
    for solver in ["rk4", "heun_explicit", "euler_explicit"]:
        full = solver == "rk4"
        for nb in (1, 2) if full else (1,):
            m, hu, f, _ = ref_cases.sitter_socont(hb, solver=solver, soil_storage_nb=nb, end_date="1981-12-31")
            bundle(f"socont_sitter_{solver}_{nb}store", m, hu, f,
                   "reference Python layer, ch_sitter_appenzell 35 bands, 1981, linear_storage, record_all")
        variants = [("glacier_1store", dict(soil_storage_nb=1)),
                    ("glacier_2store", dict(soil_storage_nb=2)),
                    ("noglacier_1store", dict(soil_storage_nb=1, with_glacier=False)),
                    ("glacier_finite_1store", dict(soil_storage_nb=1, infinite_storage=False)),
                    ("glacier_linear_1store", dict(soil_storage_nb=1, surface_runoff="linear_storage"))]
        for tag, kw in variants if full else [variants[1], variants[3]]:
            m, hu, f, _ = ref_cases.gsm_socont(hb, solver=solver, end_date="1982-12-31", n_units=10, **kw)
            bundle(f"gsm_socont_{solver}_{tag}", m, hu, f,
                   "reference Python layer, 10 synthetic bands with slope/glacier, Rhone-Gletsch meteo 1981-1982")

    action_bundles(ref)

    # C1: full bundled catchment, 1981-2020, outlet + station series only (fused spatialisation input).
    m, hu, f, params = ref_cases.sitter_socont(hb, solver="rk4", soil_storage_nb=1, end_date="2020-12-31")
    structures, units, fdata, q, rec = ref_cases.collect(m, hu, f)
    station = {str(n): np.asarray(d) for n, d in zip(f.data1D.data_name, f.data1D.data)}
    meta = {"structures": structures, "units": units, "solver": "rk4", "labels": [],
            "spatialisation": {"temperature": {"method": "additive_elevation_gradient", "ref_elevation": 1250,
                                               "gradient": -0.6},
                               "precipitation": {"method": "multiplicative_elevation_gradient",
                                                 "ref_elevation": 1250, "gradient": 0.05,
                                                 "correction_factor": 1},
                               "pet": {"method": "constant"}},
            "source": "reference Python layer, ch_sitter_appenzell 35 bands, 1981-2020 (BASELINE config C1)"}
    # spot-check columns of the spatialised forcing so the fused path can be verified without 12 MB
    extra = {"station_" + k: v for k, v in station.items()}
    for k, v in fdata.items():
        extra["spatial_first_" + k] = v[:, 0]
        extra["spatial_last_" + k] = v[:, -1]
    save("c1_sitter_rk4_full", meta, {}, q, {}, extra=extra)
    snow_ice_bundles(hb)
    sequential_bundles(hb, ref)
    melt_variant_bundles(ref)
    snow_slide_bundles(ref)
    threshold_splitter_bundles(ref)
    rain_only_bundles(ref)
    snow_retention_bundles(ref)
    hbv_bundles(ref)
    gr_bundles(ref)
    two_glaciers_bundles(ref)
    two_glaciers_action_bundles(ref)
    slow_order_bundles(ref)


if __name__ == "__main__":
    main()
