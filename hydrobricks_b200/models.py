"""Host mirror of the reference's model classes for the hot path: `Socont` with the option names, parameter
aliases and structure-generation call sequence of python/src/hydrobricks/models/socont.py:111-215,
models/model.py:1040-1105, models/model_settings.py:130-347 and modules/glacier.py:87-131, driving any
`_hydrobricks`-compatible backend module (this package's, or the compiled reference in tests).

It exists so that the engine can be used — and benchmarked, and tested on the GPU box — without the
reference's Python package; with that package installed, inject hydrobricks_b200._hydrobricks instead and use
hb.models.Socont unchanged.
"""
from __future__ import annotations

import numpy as np

def default_backend():
    """The compiled C++/pybind11 host module (hydrobricks_b200/native/_hydrobricks*.so); the pure-Python mirror
    hydrobricks_b200._hydrobricks offers the same surface (backend="python") — both drive the same C-ABI / CUDA
    engine, neither computes on the CPU."""
    from .native import _hydrobricks as native

    return native


def resolve_backend(backend):
    if backend is None or backend == "native":
        return default_backend()
    if backend == "python":
        from . import _hydrobricks as mirror

        return mirror
    return backend

RAIN_SNOWMELT_STORAGE = "glacier_area_rain_snowmelt_storage"
ICEMELT_STORAGE = "glacier_area_icemelt_storage"


class Socont:
    def __init__(self, solver="crank_nicolson", soil_storage_nb=1, surface_runoff="socont_runoff",
                 land_cover_names=None, land_cover_types=None, record_all=False, glacier_infinite_storage=True,
                 snow_melt_process="melt:degree_day", snow_ice_transformation=None, snow_sublimation_process=None,
                 snow_redistribution=None, snow_redistribution_skip_glaciers=False, snow_rain_process=None,
                 snow_water_retention_process=None, snow_refreezing_process=None, rain_to_snowpack=False,
                 backend=None):
        self.backend = resolve_backend(backend)
        self.solver = solver
        self.options = {"soil_storage_nb": soil_storage_nb, "surface_runoff": surface_runoff,
                        "snow_melt_process": snow_melt_process, "glacier_infinite_storage": glacier_infinite_storage,
                        "snow_ice_transformation": snow_ice_transformation,
                        "snow_sublimation_process": snow_sublimation_process,
                        "snow_redistribution": snow_redistribution,
                        "snow_rain_process": snow_rain_process,  # model_settings.py:194-200; None -> snow_rain:linear
                        "snow_redistribution_skip_glaciers": snow_redistribution_skip_glaciers,
                        # models/model.py:1184-1217 -> model_settings.py:139-141: liquid water kept in the snowpack
                        # ('outflow:snow_holding'), refrozen ('refreeze:degree_day'), rain routed through the snowpack
                        "snow_water_retention_process": snow_water_retention_process,
                        "snow_refreezing_process": snow_refreezing_process,
                        "rain_to_snowpack": rain_to_snowpack}
        if snow_refreezing_process and not snow_water_retention_process:  # model_settings.py:233-239
            raise ValueError("Snow refreezing requires a snow water retention process.")
        if rain_to_snowpack and not snow_water_retention_process:  # model_settings.py:240-247
            raise ValueError("Routing the rain to the snowpacks requires a snow water retention process.")
        self.land_cover_names = list(land_cover_names or ["open"])
        self.land_cover_types = list(land_cover_types or ["open"])
        self.record_all = record_all
        if soil_storage_nb not in (1, 2):
            raise ValueError("There can be only one or two groundwater storages.")
        if surface_runoff not in ("socont_runoff", "linear_storage"):
            raise ValueError(f"The surface runoff option {surface_runoff} is not recognised in Socont.")
        self.settings = self.backend.SettingsModel()
        self.settings.log_all(record_all)
        self.settings.set_solver(solver)
        self.model = None
        self._define_structure()
        self._generate_structure()
        self._define_parameter_aliases()

    # models/socont.py:111-215 + modules/glacier.py:87-131
    def _define_structure(self):
        st = {}
        glaciers = [n for n, t in zip(self.land_cover_names, self.land_cover_types) if t == "glacier"]
        for name in glaciers:
            st[name] = {"attach_to": "hydro_unit", "kind": "land_cover",
                        "parameters": {"no_melt_when_snow_cover": True,
                                       "infinite_storage": self.options["glacier_infinite_storage"]},
                        "processes": {
                            "outflow_rain_snowmelt": {"kind": "outflow:direct", "target": RAIN_SNOWMELT_STORAGE,
                                                      "instantaneous": True},
                            "melt": {"kind": self.options["snow_melt_process"], "target": ICEMELT_STORAGE,
                                     "instantaneous": True}}}
        if glaciers:
            for name in (RAIN_SNOWMELT_STORAGE, ICEMELT_STORAGE):
                st[name] = {"attach_to": "sub_basin", "kind": "storage",
                            "processes": {"outflow": {"kind": "outflow:linear", "target": "outlet"}}}
        soil = next((n for n, t in zip(self.land_cover_names, self.land_cover_types) if t != "glacier"), "open")
        st[soil] = {"attach_to": "hydro_unit", "kind": "land_cover",
                    "processes": {"infiltration": {"kind": "infiltration:socont", "target": "slow_reservoir"},
                                  "runoff": {"kind": "outflow:rest", "target": "surface_runoff"}}}
        st["slow_reservoir"] = {"attach_to": "hydro_unit", "kind": "storage", "parameters": {"capacity": 200},
                                "processes": {"et": {"kind": "et:socont"},
                                              "outflow": {"kind": "outflow:linear", "target": "outlet"},
                                              "overflow": {"kind": "overflow", "target": "outlet"}}}
        if self.options["soil_storage_nb"] == 2:
            st["slow_reservoir"]["processes"]["percolation"] = {"kind": "percolation:constant",
                                                                "target": "slow_reservoir_2"}
            st["slow_reservoir_2"] = {"attach_to": "hydro_unit", "kind": "storage",
                                      "processes": {"outflow": {"kind": "outflow:linear", "target": "outlet"}}}
        kind = "runoff:socont" if self.options["surface_runoff"] == "socont_runoff" else "outflow:linear"
        st["surface_runoff"] = {"attach_to": "hydro_unit", "kind": "storage",
                                "processes": {"runoff": {"kind": kind, "target": "outlet"}}}
        self.structure = st
        self._glacier_names = glaciers

    def _variants(self):  # models/model.py:1129-1168
        if not self._glacier_names:
            return [(self.land_cover_names, self.land_cover_types, self.structure)]
        names = [n for n, t in zip(self.land_cover_names, self.land_cover_types) if t != "glacier"]
        types = [t for t in self.land_cover_types if t != "glacier"]
        base = {k: v for k, v in self.structure.items() if k not in self._glacier_names}
        return [(names, types, base), (self.land_cover_names, self.land_cover_types, self.structure)]

    def _generate_structure(self):  # models/model.py:1040-1105, model_settings.py:130-347
        s = self.settings
        for i, (names, types, structure) in enumerate(self._variants()):
            if i > 0:
                s.add_structure()
            s.generate_precipitation_splitters(True, self.options["snow_rain_process"] or "snow_rain:linear")
            for t, n in zip(types, names):
                s.add_land_cover_brick(n, "generic_land_cover" if t in ("ground", "generic_land_cover", "open",
                                                                        "forest", "lake") else t)
            if self.options["snow_water_retention_process"]:  # model_settings.py:248-255
                s.generate_snowpacks_with_water_retention(self.options["snow_melt_process"],
                                                          self.options["snow_water_retention_process"],
                                                          self.options["rain_to_snowpack"])
                if self.options["snow_refreezing_process"]:
                    s.add_snowpack_refreezing(self.options["snow_refreezing_process"])
            else:
                s.generate_snowpacks(self.options["snow_melt_process"])
            if self.options["snow_ice_transformation"]:  # model_settings.py:258-259
                s.add_snow_ice_transformation(self.options["snow_ice_transformation"])
            if self.options["snow_redistribution"]:  # model_settings.py:260-261
                s.add_snow_redistribution(self.options["snow_redistribution"],
                                          self.options["snow_redistribution_skip_glaciers"])
            if self.options["snow_sublimation_process"]:  # model_settings.py:262-263
                s.add_snowpack_sublimation(self.options["snow_sublimation_process"])
            for key, brick in structure.items():
                if brick["kind"] == "land_cover":
                    s.select_hydro_unit_brick(key)
                elif brick["attach_to"] == "hydro_unit":
                    s.add_hydro_unit_brick(key, brick["kind"])
                else:
                    s.add_sub_basin_brick(key, brick["kind"])
                for pname, value in brick.get("parameters", {}).items():
                    s.add_brick_parameter(pname, value)
                for pname, pd in brick.get("processes", {}).items():
                    s.add_brick_process(pname, pd["kind"], pd.get("target", ""), False)
                    if pd["kind"] in ("outflow:direct", "outflow:rest"):
                        s.set_process_outputs_as_static()
                    if pd.get("instantaneous"):
                        s.set_process_outputs_as_instantaneous()
            s.add_logging_to("outlet")

    def _define_parameter_aliases(self):
        """models/socont.py:217-262 + parameters.py aliases: alias -> (component, parameter name)."""
        soil = next((n for n, t in zip(self.land_cover_names, self.land_cover_types) if t != "glacier"), "open")
        a = {"A": ("slow_reservoir", "capacity"), "k_slow": ("slow_reservoir", "response_factor"),
             "k_slow_1": ("slow_reservoir", "response_factor"), "k_slow_2": ("slow_reservoir_2", "response_factor"),
             "percol": ("slow_reservoir", "percolation_rate"), "k_quick": ("surface_runoff", "response_factor"),
             "beta": ("surface_runoff", "beta"), "melt_t_snow": ("type:snowpack", "melting_temperature"),
             "prec_t_start": ("snow_rain_transition", "transition_start"),
             "prec_t_end": ("snow_rain_transition", "transition_end")}
        glaciers = ",".join(self._glacier_names)
        smp = self.options["snow_melt_process"]  # parameters.py:1755-1820 (glacier ice), 1940-1975 (snowpacks)
        if smp == "melt:degree_day":
            a["a_snow"] = ("type:snowpack", "degree_day_factor")
            if glaciers:
                a["a_ice"] = (glaciers, "degree_day_factor")
        elif smp == "melt:degree_day_aspect":
            for suffix in ("n", "s", "ew"):
                a[f"a_snow_{suffix}"] = ("type:snowpack", f"degree_day_factor_{suffix}")
                if glaciers:
                    a[f"a_ice_{suffix}"] = (glaciers, f"degree_day_factor_{suffix}")
        elif smp == "melt:temperature_index":
            shared = "type:snowpack,type:glacier" if glaciers else "type:snowpack"  # one melt factor for snow and ice
            a["melt_factor"] = a["mf"] = (shared, "melt_factor")
            a["r_snow"] = ("type:snowpack", "radiation_coefficient")
            if glaciers:
                a["r_ice"] = (glaciers, "radiation_coefficient")
        else:
            raise ValueError(f"The snow melt process option {smp} is not recognised.")
        if glaciers:
            a["melt_t_ice"] = (glaciers, "melting_temperature")
        if len(self._glacier_names) > 1:  # parameters.py:1766-1835: one alias per glacier cover, by name and by index
            per_cover = {"melt:degree_day": [("a_ice", "degree_day_factor")],
                         "melt:degree_day_aspect": [(f"a_ice_{x}", f"degree_day_factor_{x}") for x in ("n", "s", "ew")],
                         "melt:temperature_index": [("r_ice", "radiation_coefficient")]}[smp]
            for i, name in enumerate(self._glacier_names):
                for alias, pname in per_cover + [("melt_t_ice", "melting_temperature")]:
                    a[f"{alias}_{name.replace('-', '_')}"] = a[f"{alias}_{i}"] = (name, pname)
        if self.options["snow_sublimation_process"] == "sublimation:constant":  # parameters.py:286-305, 2006-2018
            a["sublimation_rate"] = ("type:snowpack", "sublimation_rate")
        elif self.options["snow_sublimation_process"] == "sublimation:pet":
            a["sublimation_pet_factor"] = ("type:snowpack", "sublimation_pet_factor")
            a["sublimation_factor"] = ("type:snowpack", "sublimation_pet_factor")
        if self.options["snow_water_retention_process"]:  # models/hbv.py:367: cwh / whc, cfr
            a["whc"] = a["cwh"] = ("type:snowpack", "water_holding_capacity")
        if self.options["snow_refreezing_process"]:
            a["cfr"] = ("type:snowpack", "refreezing_factor")
        if self._glacier_names:
            a.update({"k_snow": (RAIN_SNOWMELT_STORAGE, "response_factor"),
                      "k_ice": (ICEMELT_STORAGE, "response_factor")})
            snowpacks = ",".join(n + "_snowpack" for n in self._glacier_names)  # parameters.py:2021-2070
            if self.options["snow_ice_transformation"] in ("transform:snow_ice_constant",
                                                           "transformation:snow_ice_constant"):
                a["snow_ice_rate"] = (snowpacks, "snow_ice_transformation_rate")
            elif self.options["snow_ice_transformation"]:
                a["snow_ice_basal_acc_coeff"] = (snowpacks, "snow_ice_transformation_basal_acc_coeff")
        self.aliases = a
        self.soil_cover = soil

    # -- the reference's Model surface (models/model.py:118-332, 450-600) ----------------------------------
    def setup(self, units, start_date, end_date, spinup_days=0, device=0):
        """units: list of dicts {"id","area","elevation","slope":(value,unit)|None,"land_covers":[(name,frac)]}
        in any order; sorted by descending elevation like HydroUnits (hydro_units.py:693-736)."""
        self.units = sorted(units, key=lambda u: -u["elevation"])
        self.settings.set_timer(start_date, end_date, 1, "day")
        if spinup_days:
            self.settings.set_spinup_days(spinup_days)
        basin = self.backend.SettingsBasin()
        for u in self.units:
            basin.add_hydro_unit(u["id"], u["area"], u["elevation"])
            if u.get("slope"):
                basin.add_hydro_unit_property_double("slope", u["slope"][0], u["slope"][1])
            if u.get("aspect_class"):  # hydro_units.py: the property melt:degree_day_aspect reads
                basin.add_hydro_unit_property_str("aspect_class", u["aspect_class"])
            for name, frac in u["land_covers"]:
                basin.add_land_cover(name, "", frac)
        for u in self.units:  # hydro_units.py set_connectivity: (receiver position in basin order, fraction)
            for receiver, fraction in u.get("lateral", []):
                basin.add_lateral_connection(u["id"], self.units[receiver]["id"], fraction, "snow_slide")
        self.model = self.backend.ModelHydro()
        try:
            self.model.init_with_basin(self.settings, basin, device)
        except TypeError:
            self.model.init_with_basin(self.settings, basin)
        self._basin = basin
        self._n_steps = int((np.datetime64(end_date) - np.datetime64(start_date)).astype(int)) + 1

    def set_parameters(self, values: dict):
        for alias, value in values.items():
            component, name = self.aliases.get(alias, (None, None))
            if component is None:
                component, name = alias.rsplit(":", 1)  # "type:snowpack:coeff" -> ("type:snowpack", "coeff")
            if not self.settings.set_parameter_value(component, name, value):
                raise ValueError(f"Failed setting parameter {alias}")
        self.model.update_parameters(self.settings)

    def set_forcing(self, time_mjd, precipitation, temperature, pet, solar_radiation=None):
        """Pre-spatialised forcing [T x U] in the basin order of self.units (model.py:354-384); solar_radiation only
        with snow_melt_process="melt:temperature_index"."""
        ids = np.array([u["id"] for u in self.units], dtype=np.int32)
        self.model.clear_time_series()
        series = [("precipitation", precipitation), ("temperature", temperature), ("pet", pet)]
        if solar_radiation is not None:
            series.append(("solar_radiation", solar_radiation))
        for name, data in series:
            if not self.model.create_time_series(name, np.asarray(time_mjd, dtype=np.float64), ids, data):
                raise ValueError("Failed adding time series.")
        if not self.model.attach_time_series_to_hydro_units():
            raise ValueError("Attaching time series failed.")

    def run(self, parameters: dict | None = None):
        if parameters:
            self.set_parameters(parameters)
        self.model.reset()
        self.model.run()

    def get_outlet_discharge(self):
        return self.model.get_outlet_discharge()

    # -- batched extras --------------------------------------------------------------------------------
    def parameter_matrix(self, sets: dict, n_sets: int):
        """{alias: array[n_sets] | scalar} -> float32 [n_sets][HBK_N_PARAMS] (engine.PARAM_SLOTS order)."""
        base = np.asarray(self.model.base_parameters(), dtype=np.float32)
        m = np.repeat(base[None, :], n_sets, axis=0)
        for alias, values in sets.items():
            component, name = self.aliases.get(alias, (None, None))
            if component is None:
                component, name = alias.rsplit(":", 1)  # "type:snowpack:coeff" -> ("type:snowpack", "coeff")
            slots = self.model.parameter_slot(component, name)
            if not slots:
                raise ValueError(f"parameter {alias} is not part of this structure")
            for slot in slots:
                m[:, slot] = np.asarray(values, dtype=np.float32)
        return m

    def run_sets(self, matrix):
        """Integrate the parameter sets [n_sets x HBK_N_PARAMS] in one launch and leave the outlet on the device
        (evaluate_batch scores it there; get_outlet_discharge_batch fetches it)."""
        self._last_batch = len(matrix)
        self.model.set_parameter_sets(matrix)
        self.model.run()

    def run_batch(self, matrix):
        self._last_batch = len(matrix)
        return self.model.run_batch(matrix)

    def evaluate_batch(self, metric, observations, warmup=0):
        """Objective score (engine.SCORES: "nse", "kge_2012", "kge_2009", "rmse", "mae", ...) of every parameter set of the last run_batch against the observed
        outlet discharge, computed on the device (python evaluation/metrics.py:40-62, trainer.py:1091-1141)."""
        return self.engine(getattr(self, "_last_batch", 1)).compute_scores(metric, observations, warmup)

    def engine(self, n_sets=0):
        """ctypes view of the underlying hbk_engine (device pointers, stream, statistics) for either backend."""
        from . import engine as _engine

        if hasattr(self.model, "_engine"):
            return self.model._engine
        n_engine = self.model.engine_unit_count() if hasattr(self.model, "engine_unit_count") else len(self.units)
        eng = _engine.Engine.from_handle(self.model.engine_handle(), n_engine, self._n_steps, n_sets)
        if n_engine > len(self.units):  # twin units carrying a second glacier cover (hbk_set_unit_twins)
            from . import structure as _structure

            twin_of = list(self.model.twin_of())
            eng.layout = _structure.UnitLayout(len(self.units), twin_of[len(self.units):])
        return eng
