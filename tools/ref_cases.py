"""Cases run through the UNMODIFIED reference (Python layer + compiled core, oracle/_ref) to pin the
oracle restatement and to generate tests/golden/*.npz. Build container only (needs /root/reference).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refpy  # noqa: E402

CATCH = Path("/root/reference/tests/files/catchments")


def units_from_hydro_units(hydro_units, land_cover_names):
    """Basin-order unit table (descending elevation, hydro_units.py:693-736) for the oracle."""
    df = hydro_units.hydro_units.copy()
    df = df.sort_values(by=("elevation", "m"), ascending=False)
    units = []
    for _, row in df.iterrows():
        u = {"id": int(row["id"].values[0]), "area": float(row["area"].values[0]),
             "elevation": float(row["elevation"].values[0]), "slope": None, "land_covers": []}
        if "slope" in df.columns.get_level_values(0):
            unit = [c[1] for c in df.columns if c[0] == "slope"][0]
            u["slope"] = (float(row["slope"].values[0]), unit)
        for name in land_cover_names:
            frac = float(row["fraction-" + name].values[0])
            u["land_covers"].append((name, 0.0 if np.isnan(frac) else frac))
        units.append(u)
    return units


def sitter_socont(hb, solver="rk4", soil_storage_nb=1, surface_runoff="linear_storage", end_date="2020-12-31",
                  params=None):
    import hydrobricks.models as models

    C = CATCH / "ch_sitter_appenzell"
    socont = models.Socont(soil_storage_nb=soil_storage_nb, surface_runoff=surface_runoff, record_all=True,
                           solver=solver)
    parameters = socont.generate_parameters()
    values = {"a_snow": 3, "k_quick": 0.05, "A": 200, "k_slow_1": 0.001}
    if soil_storage_nb == 2:
        values.update({"percol": 0.5, "k_slow_2": 0.005})
    values.update(params or {})
    parameters.set_values(values)
    parameters.add_data_parameter("precip_correction_factor", 1)
    parameters.add_data_parameter("precip_gradient", 0.05)
    parameters.add_data_parameter("temp_gradients", -0.6)
    hydro_units = hb.HydroUnits()
    hydro_units.load_from_csv(C / "hydro_units_elevation.csv", column_elevation="elevation", column_area="area")
    forcing = hb.Forcing(hydro_units)
    forcing.load_station_data_from_csv(
        C / "meteo.csv", column_time="date", time_format="%d/%m/%Y",
        content={"precipitation": "precip(mm/day)", "temperature": "temp(C)", "pet": "pet_sim(mm/day)"})
    forcing.spatialize_from_station_data(variable="temperature", ref_elevation=1250,
                                         gradient=parameters.get("temp_gradients"))
    forcing.spatialize_from_station_data(variable="pet")
    forcing.correct_station_data(variable="precipitation",
                                 correction_factor=parameters.get("precip_correction_factor"))
    forcing.spatialize_from_station_data(variable="precipitation", ref_elevation=1250,
                                         gradient=parameters.get("precip_gradient"))
    tmp = tempfile.mkdtemp()
    socont.setup(spatial_structure=hydro_units, output_path=tmp, start_date="1981-01-01", end_date=end_date)
    socont.run(parameters=parameters, forcing=forcing)
    return socont, hydro_units, forcing, parameters


def synthetic_units_csv(path, n_units=12, seed=3, with_glacier=True, area=1.0e6):
    """Elevation bands 1700..3300 m with slope (deg) and a glacier fraction ramp above 2400 m."""
    rng = np.random.default_rng(seed)
    elev = np.linspace(1700, 3300, n_units)
    slope = rng.uniform(5, 35, n_units)
    areas = area * rng.uniform(0.4, 1.6, n_units)
    with open(path, "w") as f:
        if with_glacier:
            f.write("id,elevation,area_ground,area_glacier,slope\n-,m,m2,m2,deg\n")
        else:
            f.write("id,fraction-ground,elevation,area,slope\n-,fraction,m,m2,deg\n")
        for i in range(n_units):
            g = float(np.clip((elev[i] - 2400) / 800 * 0.8, 0, 0.8)) if with_glacier else 0.0
            g = round(g, 4)
            if with_glacier:
                f.write(f"{i + 1},{elev[i]:.1f},{areas[i] * (1 - g):.1f},{areas[i] * g:.1f},{slope[i]:.3f}\n")
            else:
                f.write(f"{i + 1},1,{elev[i]:.1f},{areas[i]:.1f},{slope[i]:.3f}\n")


def gsm_socont(hb, solver="rk4", soil_storage_nb=1, surface_runoff="socont_runoff", with_glacier=True,
               end_date="1990-12-31", n_units=12, params=None, infinite_storage=True, record_all=True,
               spinup_days=0, snow_ice_transformation=None, snow_sublimation_process=None):
    import hydrobricks.models as models

    tmp = tempfile.mkdtemp()
    csv = Path(tmp) / "hydro_units.csv"
    synthetic_units_csv(csv, n_units=n_units, with_glacier=with_glacier)
    names = ["ground", "glacier"] if with_glacier else ["ground"]
    types = ["ground", "glacier"] if with_glacier else ["ground"]
    kw = dict(soil_storage_nb=soil_storage_nb, surface_runoff=surface_runoff, record_all=record_all, solver=solver,
              land_cover_names=names, land_cover_types=types)
    if with_glacier:
        kw["glacier_infinite_storage"] = infinite_storage
    if snow_ice_transformation:
        kw["snow_ice_transformation"] = snow_ice_transformation
    if snow_sublimation_process:
        kw["snow_sublimation_process"] = snow_sublimation_process
    socont = models.Socont(**kw)
    parameters = socont.generate_parameters()
    values = {"a_snow": 4.5, "k_quick": 0.3, "A": 120, "k_slow_1": 0.02, "beta": 800}
    if surface_runoff == "linear_storage":
        values.pop("beta")
    else:
        values.pop("k_quick")
    if with_glacier:
        values.update({"a_ice": 7.0, "k_snow": 0.2, "k_ice": 0.35})
    if soil_storage_nb == 2:
        values.update({"percol": 0.8, "k_slow_2": 0.006})
    values.update(params or {})
    parameters.set_values(values)
    parameters.add_data_parameter("precip_gradient", 0.04)
    parameters.add_data_parameter("temp_gradients", -0.55)
    hydro_units = hb.HydroUnits(land_cover_types=types, land_cover_names=names)
    if with_glacier:
        hydro_units.load_from_csv(csv, column_elevation="elevation",
                                  columns_areas={"ground": "area_ground", "glacier": "area_glacier"},
                                  other_columns={"slope": "slope"})
    else:
        hydro_units.load_from_csv(csv, column_elevation="elevation", column_area="area",
                                  other_columns={"slope": "slope"})
    forcing = hb.Forcing(hydro_units)
    forcing.load_station_data_from_csv(
        CATCH / "ch_rhone_gletsch" / "meteo.csv", column_time="date", time_format="%d/%m/%Y",
        content={"precipitation": "precip(mm/day)", "temperature": "temp(C)", "pet": "pet_sim(mm/day)"})
    forcing.spatialize_from_station_data(variable="temperature", ref_elevation=2702,
                                         gradient=parameters.get("temp_gradients"))
    forcing.spatialize_from_station_data(variable="pet")
    forcing.spatialize_from_station_data(variable="precipitation", ref_elevation=2702,
                                         gradient=parameters.get("precip_gradient"))
    socont.setup(spatial_structure=hydro_units, output_path=tmp, start_date="1981-01-01", end_date=end_date,
                 **({"spinup": spinup_days} if spinup_days else {}))
    socont.run(parameters=parameters, forcing=forcing)
    return socont, hydro_units, forcing, parameters


def collect(model, hydro_units, forcing):
    """Everything the oracle / engine needs to repeat the run + everything the reference recorded."""
    structures = model.settings.settings.get_structure()
    units = units_from_hydro_units(hydro_units, model.land_cover_names)
    n = len(model.get_outlet_discharge())
    # Forcing as handed to ModelHydro::CreateTimeSeries (model.py:354-384): data[T x U] in the order of
    # spatial_structure.get_ids(); re-order the columns to basin order.
    ids = model.spatial_structure.get_ids().to_numpy().flatten()
    order = [int(np.where(ids == u["id"])[0][0]) for u in units]
    time = forcing.data2D.time.to_numpy()
    t0 = int(np.where(time == np.datetime64(model.start_date if hasattr(model, "start_date") else time[0]))[0][0]) \
        if False else 0
    fdata = {}
    for name, data in zip(forcing.data2D.data_name, forcing.data2D.data):
        fdata[str(name)] = np.asarray(data)[t0:t0 + n][:, order]
    rec = {label: model.model.get_hydro_unit_values(label) for label in model.model.get_recorded_hydro_unit_labels()}
    return structures, units, fdata, model.get_outlet_discharge(), rec


if __name__ == "__main__":
    hb = refpy.load_reference_python("ref")
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import hb_oracle

    cases = [("sitter", dict(soil_storage_nb=1, surface_runoff="linear_storage")),
             ("sitter", dict(soil_storage_nb=2, surface_runoff="linear_storage")),
             ("gsm", dict(soil_storage_nb=1, surface_runoff="socont_runoff", with_glacier=True)),
             ("gsm", dict(soil_storage_nb=2, surface_runoff="socont_runoff", with_glacier=True)),
             ("gsm", dict(soil_storage_nb=1, surface_runoff="socont_runoff", with_glacier=False)),
             ("gsm", dict(soil_storage_nb=1, surface_runoff="socont_runoff", with_glacier=True,
                          infinite_storage=False))]
    solvers = sys.argv[1:] or ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "implicit_euler",
                               "exponential_euler", "analytic_linear"]
    for solver in solvers:
        for kind, kw in cases:
            fn = sitter_socont if kind == "sitter" else gsm_socont
            model, hu, forcing, params = fn(hb, solver=solver, end_date="1990-12-31", **kw)
            nb, sr = kind, kw
            structures, units, fdata, q_ref, rec_ref = collect(model, hu, forcing)
            om = hb_oracle.OracleModel(structures, units, solver, len(q_ref))
            for k, v in fdata.items():
                om.set_forcing(k, v)
            for label in rec_ref:
                om.record(label)
            q = om.run()
            nbad = int(np.sum(q != q_ref))
            worst = max(float(np.nanmax(np.abs(om.records[l] - rec_ref[l]))) for l in rec_ref)
            print(solver, nb, sr, "outlet mismatches:", nbad, "max abs diff:", float(np.max(np.abs(q - q_ref))),
                  "worst record diff:", worst, "unconverged:", om.unconverged)
