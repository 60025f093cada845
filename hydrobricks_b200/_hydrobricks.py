"""Host-side mirror of the reference's `_hydrobricks` extension module (core/bindings/Binding.cpp:164-413)
for the ModelHydro::Run() hot path, on top of the C-ABI of include/hbk.h.

Same class names, method names, argument meaning and error behaviour as the pybind11 module the
reference Python layer imports (`from hydrobricks._hydrobricks import ModelHydro, ...`,
python/src/hydrobricks/models/model.py:13), so `hb.models.Socont(...).setup/run/get_outlet_discharge` work
unchanged with this module injected as `hydrobricks._hydrobricks` (tools/refpy.py, backend "b200").
Batched extras the reference does not have: ModelHydro.run_batch / get_outlet_discharge_batch.

SettingsModel / SettingsBasin are declarative (cold path); ModelHydro.init_with_basin compiles them with
structure.CompiledStructure and creates the CUDA engine. Structures outside the Socont family raise
ValueError there — nothing is ever computed on the CPU.
"""
from __future__ import annotations

import copy

import numpy as np

from . import engine as _engine
from . import schedule as _schedule
from . import structure as _structure

_FORCING_NAMES = {"precipitation", "pet", "temperature", "temperature_min", "temperature_max", "solar_radiation",
                  "r_solar"}

# Process::RegisterSettings of the reference (core/src/processes/Process*.cpp RegisterProcessSettings):
# default parameters and forcing of every process type the engine runs.
_PROCESS_REGISTRY = {
    "melt:degree_day": ([("degree_day_factor", 3.0), ("melting_temperature", 0.0)], ["temperature"]),
    # ProcessMeltDegreeDayAspect.cpp:13-19, ProcessMeltTemperatureIndex.cpp:14-20
    "melt:degree_day_aspect": ([("degree_day_factor_n", 3.0), ("degree_day_factor_s", 3.0), ("degree_day_factor_ew", 3.0),
                                ("melting_temperature", 0.0)], ["temperature"]),
    "melt:temperature_index": ([("melt_factor", 3.0), ("melting_temperature", 0.0),
                                ("radiation_coefficient", 0.0007)], ["temperature", "solar_radiation"]),
    "et:socont": ([], ["pet"]),
    "infiltration:socont": ([], []),
    "runoff:socont": ([("beta", 500.0)], []),
    "outflow:linear": ([("response_factor", 0.2)], []),
    "outflow:direct": ([], []),
    "outflow:rest": ([], []),
    "overflow": ([], []),
    "outflow:overflow": ([], []),
    "percolation:constant": ([("percolation_rate", 0.1)], []),
    # ProcessTransformSnowToIce{Constant,Swat}::RegisterProcessSettings; both spellings (Process.cpp:65-66)
    "transform:snow_ice_constant": ([("snow_ice_transformation_rate", 0.5)], []),
    "transformation:snow_ice_constant": ([("snow_ice_transformation_rate", 0.5)], []),
    "transform:snow_ice_swat": ([("snow_ice_transformation_basal_acc_coeff", 0.0014), ("north_hemisphere", 1)], []),
    "transformation:snow_ice_swat": ([("snow_ice_transformation_basal_acc_coeff", 0.0014), ("north_hemisphere", 1)], []),
    # ProcessLateralSnowSlide::RegisterProcessSettings (ProcessLateralSnowSlide.cpp:24-31)
    "transport:snow_slide": ([("coeff", 3178.4), ("exp", -1.998), ("min_slope", 10.0), ("max_slope", 75.0),
                              ("min_snow_holding_depth", 50.0), ("max_snow_depth", 20000.0)], []),
    # ProcessSublimation{Constant,PET}::RegisterProcessSettings
    "sublimation:constant": ([("sublimation_rate", 0.1)], []),
    "sublimation:pet": ([("sublimation_pet_factor", 0.2)], ["pet"]),
    # ProcessOutflowSnowHolding.cpp:15-17, ProcessRefreezeDegreeDay.cpp:14-17
    "outflow:snow_holding": ([("water_holding_capacity", 0.1)], []),
    "refreeze:degree_day": ([("refreezing_factor", 0.05)], ["temperature"]),
}

_log_level = "message"


def init():
    """InitHydrobricksForPython (Utils.h:16)."""
    return True


def init_log(path):
    return None


def close_log():
    return None


def set_max_log_level():
    global _log_level
    _log_level = "max"


def set_debug_log_level():
    global _log_level
    _log_level = "debug"


def set_message_log_level():
    global _log_level
    _log_level = "message"


class LogNull:
    pass


class ParameterModifierType:
    """base/ParameterModifier.h: what the modifier's values are indexed by."""
    Yearly, Monthly, Dates = "Yearly", "Monthly", "Dates"


class ParameterModifier:
    """base/ParameterModifier.cpp: a parameter value looked up by the year, the month or the exact date of a modified
    Julian date (NaN when the date has no entry). A standalone object, as in the reference's binding
    (Binding.cpp:284-303): nothing attaches it to a running model, time-varying parameters inside a run are out of
    scope."""

    def __init__(self, type=ParameterModifierType.Yearly):  # noqa: A002 (the binding's argument name)
        self._type, self._dates, self._values = type, [], []

    def set_type(self, type):  # noqa: A002
        self._type = type

    def get_type(self):
        return self._type

    def _require(self, wanted, method):
        if self._type != wanted:
            raise RuntimeError(f"ParameterModifier::{method} - Modifier type is not {wanted}.")

    def set_yearly_values(self, year_start, year_end, values):
        self._require(ParameterModifierType.Yearly, "SetYearlyValues")
        self._values = [float(np.float32(v)) for v in values]
        self._dates = [float(y) for y in range(int(year_start), int(year_end) + 1)]
        return len(self._values) == len(self._dates)

    def set_monthly_values(self, values):
        self._require(ParameterModifierType.Monthly, "SetMonthlyValues")
        self._values = [float(np.float32(v)) for v in values]
        self._dates = []
        return len(self._values) == 12

    def set_dates_and_values(self, dates, values):
        self._require(ParameterModifierType.Dates, "SetDatesAndValues")
        self._dates = [float(d) for d in dates]
        self._values = [float(np.float32(v)) for v in values]
        return len(self._values) == len(self._dates)

    def update_value(self, date):
        day = np.datetime64("1858-11-17") + np.timedelta64(int(np.floor(date)), "D")
        year = int(str(day)[:4])
        month = int(str(day)[5:7])
        if self._type == ParameterModifierType.Monthly:
            return self._values[month - 1] if len(self._values) == 12 else float("nan")
        key = float(year) if self._type == ParameterModifierType.Yearly else float(date)
        for d, v in zip(self._dates, self._values):
            if d == key:
                return v
        return float("nan")

    def updates_on_year_change(self):
        return self._type == ParameterModifierType.Yearly

    def updates_on_month_change(self):
        return self._type == ParameterModifierType.Monthly

    def updates_on_date_change(self):
        return self._type == ParameterModifierType.Dates


class Parameter:
    """base/Parameter.h, Parameter.cpp: name + float32 value + optional modifier."""

    def __init__(self, name, value):
        self.name = name
        self.value = float(np.float32(value))
        self._modifier = None

    def get_name(self):
        return self.name

    def set_name(self, name):
        self.name = name

    def get_value(self):
        return self.value

    def set_value(self, value):
        self.value = float(np.float32(value))

    def set_modifier(self, modifier):
        self._modifier = modifier

    def has_modifier(self):
        return self._modifier is not None

    def update_from_modifier(self, date):
        if self._modifier is None:
            return False
        value = self._modifier.update_value(date)
        if value != value:  # NaN: no entry for this date
            return False
        self.value = value
        return True

    def __repr__(self):
        return f"<_hydrobricks.Parameter named '{self.name}'>"


def _param(name, value):
    return {"name": name, "value": float(np.float32(value))}


class SettingsModel:
    """base/SettingsModel.{h,cpp}: select-then-add builder of the model structure."""

    def __init__(self):
        self._log_all = False
        self._record_fractions = False
        self._structures = [self._new_structure(1)]
        self._sel_structure = self._structures[0]
        self._sel_brick = None
        self._sel_process = None
        self._sel_splitter = None
        self.solver = ""
        self.timer = {"start": "", "end": "", "time_step": 1, "time_step_unit": "", "spinup_days": 0}

    @staticmethod
    def _new_structure(sid):
        return {"id": sid, "log_items": [], "hydro_unit_bricks": [], "sub_basin_bricks": [],
                "hydro_unit_splitters": [], "sub_basin_splitters": []}

    # -- options -------------------------------------------------------------------------------------
    def log_all(self, log_all=True):
        self._log_all = bool(log_all)

    def record_fractions(self, record=True):
        self._record_fractions = bool(record)

    def set_solver(self, name):
        self.solver = name

    def set_timer(self, start_date, end_date, time_step, time_step_unit):
        self.timer.update(start=start_date, end=end_date, time_step=int(time_step), time_step_unit=time_step_unit)

    def set_spinup_days(self, days):
        if days < 0:
            raise ValueError("The spin-up duration cannot be negative.")  # SettingsModel.cpp:38-43
        self.timer["spinup_days"] = int(days)

    def get_solver_name(self):
        return self.solver

    def get_timer(self):
        return dict(self.timer)

    # -- structures ----------------------------------------------------------------------------------
    def add_structure(self):
        sid = max(s["id"] for s in self._structures) + 1
        self._structures.append(self._new_structure(sid))
        self._sel_structure = self._structures[-1]
        self._sel_brick = self._sel_process = self._sel_splitter = None
        return sid

    def select_structure(self, sid):
        for s in self._structures:
            if s["id"] == sid:
                self._sel_structure = s
                self._sel_brick = s["hydro_unit_bricks"][0] if s["hydro_unit_bricks"] else None
                return True
        return False

    # -- bricks --------------------------------------------------------------------------------------
    def _add_brick(self, where, name, kind, attach):
        brick = {"name": name, "kind": kind, "parent": None, "attach": attach, "computed_directly": False,
                 "is_land_cover": False, "is_surface_component": False, "forcing": [], "parameters": [],
                 "log_items": [], "processes": []}
        self._sel_structure[where].append(brick)
        self._sel_brick = brick
        if self._log_all:  # SettingsModel.cpp:55-62
            self.add_brick_logging("water_content")
            if kind == "glacier":
                self.add_brick_logging("ice_content")
            elif kind == "snowpack":
                self.add_brick_logging("snow_content")
        return brick

    def add_hydro_unit_brick(self, name, kind="storage"):
        self._add_brick("hydro_unit_bricks", name, kind, "hydro_unit")

    def add_sub_basin_brick(self, name, kind="storage"):
        self._add_brick("sub_basin_bricks", name, kind, "sub_basin")

    def add_land_cover_brick(self, name, kind):
        brick = self._add_brick("hydro_unit_bricks", name, kind, "hydro_unit")
        brick["is_land_cover"] = True
        if self._select_splitter_if_found("rain_splitter"):  # SettingsModel.cpp:85-94
            self._add_splitter_output(name)

    def _add_surface_component_brick(self, name, kind):
        brick = self._add_brick("hydro_unit_bricks", name, kind, "hydro_unit")
        brick["is_surface_component"] = True

    def select_hydro_unit_brick(self, name):
        for b in self._sel_structure["hydro_unit_bricks"]:
            if b["name"] == name:
                self._sel_brick, self._sel_process = b, None
                return
        raise RuntimeError(f"The hydro unit brick '{name}' was not found")

    def add_brick_parameter(self, name, value, kind="constant"):
        if kind != "constant":
            raise NotImplementedError(f"SettingsModel::AddBrickParameter - Parameter type '{kind}' not supported")
        self._sel_brick["parameters"].append(_param(name, value))

    def add_brick_forcing(self, name):
        if name not in _FORCING_NAMES:
            raise ValueError(f"The provided forcing '{name}' is not yet supported.")
        self._sel_brick["forcing"].append(name)

    def set_current_brick_computed_directly(self):
        self._sel_brick["computed_directly"] = True

    def add_brick_logging(self, name):
        if name not in self._sel_brick["log_items"]:
            self._sel_brick["log_items"].append(name)

    def add_logging_to(self, name):
        if name not in self._sel_structure["log_items"]:
            self._sel_structure["log_items"].append(name)

    # -- processes -----------------------------------------------------------------------------------
    def add_brick_process(self, name, kind, target="", log=False):
        if kind not in _PROCESS_REGISTRY:
            raise RuntimeError(
                f"Process type '{kind}' is not available in the B200 engine (Socont family only: "
                f"{', '.join(sorted(_PROCESS_REGISTRY))}).")
        proc = {"name": name, "kind": kind, "forcing": [], "outputs": [], "parameters": [], "log_items": []}
        self._sel_brick["processes"].append(proc)
        self._sel_process = proc
        if target:
            if ":" in target:  # "brick:content" (SettingsModel.cpp:186-193)
                tgt, flux_type = target.split(":", 1)
                self.add_process_output(tgt, flux_type)
            else:
                self.add_process_output(target)
        if (log or self._log_all) and "transport:" not in kind:
            self.add_process_logging("output")
        params, forcing = _PROCESS_REGISTRY[kind]
        for pname, pvalue in params:
            self.add_process_parameter(pname, pvalue)
        for f in forcing:
            self.add_process_forcing(f)

    def select_process(self, name):
        for p in self._sel_brick["processes"]:
            if p["name"] == name:
                self._sel_process = p
                return
        raise RuntimeError(f"The process '{name}' was not found.")

    def add_process_output(self, target, flux_type="water"):
        self._sel_process["outputs"].append({"target": target, "flux_type": flux_type, "instantaneous": False,
                                             "static": False})

    def add_process_parameter(self, name, value, kind="constant"):
        if kind != "constant":
            raise NotImplementedError(f"SettingsModel::AddProcessParameter - Parameter type '{kind}' not supported")
        for p in self._sel_process["parameters"]:
            if p["name"] == name:
                p["value"] = float(np.float32(value))
                return
        self._sel_process["parameters"].append(_param(name, value))

    def set_process_parameter_value(self, name, value, kind="constant"):
        for p in self._sel_process["parameters"]:
            if p["name"] == name:
                p["value"] = float(np.float32(value))
                return
        raise RuntimeError(f"SettingsModel::SetProcessParameterValue - Parameter '{name}' not found")

    def add_process_forcing(self, name):
        if name not in _FORCING_NAMES:
            raise ValueError(f"The provided forcing '{name}' is not yet supported.")
        self._sel_process["forcing"].append(name)

    def add_process_logging(self, name):
        if name not in self._sel_process["log_items"]:
            self._sel_process["log_items"].append(name)

    def set_process_outputs_as_instantaneous(self):
        for o in self._sel_process["outputs"]:
            o["instantaneous"] = True

    def set_process_outputs_as_static(self):
        for o in self._sel_process["outputs"]:
            o["static"] = True

    # -- splitters -----------------------------------------------------------------------------------
    def _add_hydro_unit_splitter(self, name, kind):
        s = {"name": name, "kind": kind, "attach": "hydro_unit", "forcing": [], "outputs": [], "parameters": [],
             "log_items": []}
        self._sel_structure["hydro_unit_splitters"].append(s)
        self._sel_splitter = s

    def _select_splitter_if_found(self, name):
        for s in self._sel_structure["hydro_unit_splitters"]:
            if s["name"] == name:
                self._sel_splitter = s
                return True
        return False

    def _add_splitter_output(self, target, flux_type="water"):
        self._sel_splitter["outputs"].append({"target": target, "flux_type": flux_type, "instantaneous": False,
                                              "static": False})

    def generate_precipitation_splitters(self, with_snow=True, splitter_type="snow_rain:linear"):
        """SettingsModel::GeneratePrecipitationSplitters (SettingsModel.cpp:491-530)."""
        if not with_snow:  # SplitterRain (SplitterRain.cpp:37-39): all precipitation is rain, no snowpacks
            self._add_hydro_unit_splitter("rain", "rain")
            self._sel_splitter["forcing"] += ["precipitation"]
            self._add_splitter_output("rain_splitter")
            self._sel_splitter["parameters"].append(_param("rain_correction_factor", 1.0))
            self._add_hydro_unit_splitter("rain_splitter", "multi_fluxes")
            return
        if splitter_type not in ("snow_rain:linear", "snow_rain:threshold"):
            raise RuntimeError(f"Splitter type '{splitter_type}' is not available in the B200 engine.")
        self._add_hydro_unit_splitter("snow_rain_transition", splitter_type)
        self._sel_splitter["forcing"] += ["precipitation", "temperature"]
        self._add_splitter_output("rain_splitter")
        self._add_splitter_output("snow_splitter", "snow")
        if splitter_type == "snow_rain:threshold":
            self._sel_splitter["parameters"].append(_param("threshold", 0.0))
        else:
            self._sel_splitter["parameters"].append(_param("transition_start", 0.0))
            self._sel_splitter["parameters"].append(_param("transition_end", 2.0))
        self._sel_splitter["parameters"].append(_param("rain_correction_factor", 1.0))
        self._sel_splitter["parameters"].append(_param("snow_correction_factor", 1.0))
        self._add_hydro_unit_splitter("snow_splitter", "multi_fluxes")
        self._add_hydro_unit_splitter("rain_splitter", "multi_fluxes")

    def generate_snowpacks(self, snow_melt_process):
        """SettingsModel::GenerateSnowpacks (SettingsModel.cpp:532-545)."""
        covers = [b["name"] for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for cover in covers:
            if not self._select_splitter_if_found("snow_splitter"):
                raise RuntimeError("The hydro unit splitter 'snow_splitter' was not found")
            self._add_splitter_output(cover + "_snowpack", "snow")
            self._add_surface_component_brick(cover + "_snowpack", "snowpack")
            self._sel_brick["parent"] = cover
            self.add_brick_process("melt", snow_melt_process, cover)

    def _unsupported(self, what):
        raise RuntimeError(f"{what} is not available in the B200 engine (SURVEY.md §8f lists it as next).")

    def generate_snowpacks_with_water_retention(self, snow_melt_process, outflow_process, rain_to_snowpack=False):
        """SettingsModel::GenerateSnowpacksWithWaterRetention (SettingsModel.cpp:608-634): the melt goes to the snowpack's
        own water container (instantaneous), `outflow_process` (outflow:snow_holding) releases it to the land cover, and
        with rain_to_snowpack the rain splitter feeds the snowpack instead of the cover."""
        covers = [b["name"] for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for cover in covers:
            if not self._select_splitter_if_found("snow_splitter"):
                raise RuntimeError("The hydro unit splitter 'snow_splitter' was not found")
            self._add_splitter_output(cover + "_snowpack", "snow")
            self._add_surface_component_brick(cover + "_snowpack", "snowpack")
            self._sel_brick["parent"] = cover
            self.add_brick_process("melt", snow_melt_process)
            self._sel_process["outputs"].append({"target": cover + "_snowpack", "flux_type": "water",
                                                 "instantaneous": True, "static": False})  # OutputProcessToSameBrick
            self.add_brick_process("meltwater", outflow_process)
            self.add_process_output(cover)
            if rain_to_snowpack:
                if not self._select_splitter_if_found("rain_splitter"):
                    raise RuntimeError("The hydro unit splitter 'rain_splitter' was not found")
                for out in self._sel_splitter["outputs"]:  # ChangeSplitterOutputTargetIfFound
                    if out["target"] == cover:
                        out["target"] = cover + "_snowpack"
                        break

    def generate_canopy_interception(self, *a, **k):
        self._unsupported("canopy interception")

    def add_snowpack_refreezing(self, refreezing_process="refreeze:degree_day"):
        """SettingsModel::AddSnowpackRefreezing (SettingsModel.cpp:574-590): on every snowpack, from its water container back
        to its snow container, instantaneous."""
        covers = [b["name"] for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for name in covers:
            self.select_hydro_unit_brick(name + "_snowpack")
            self.add_brick_process("refreeze", refreezing_process, name + "_snowpack:snow")
            self.set_process_outputs_as_instantaneous()

    def add_snowpack_sublimation(self, sublimation_process):
        """SettingsModel::AddSnowpackSublimation (SettingsModel.cpp:586-598): on every snowpack, from the snow
        container to the atmosphere (no target)."""
        covers = [b["name"] for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for name in covers:
            self.select_hydro_unit_brick(name + "_snowpack")
            self.add_brick_process("sublimation", sublimation_process)

    def add_snow_ice_transformation(self, transformation_process="transform:snow_ice_swat"):
        """SettingsModel::AddSnowIceTransformation (SettingsModel.cpp:547-559): on the snowpack of every glacier
        cover, a process that moves snow into the glacier's ice container."""
        covers = [(b["name"], b["kind"]) for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for name, kind in covers:
            self.select_hydro_unit_brick(name + "_snowpack")
            if kind == "glacier":
                self.add_brick_process("snow_ice_transfo", transformation_process, name + ":ice")

    def add_snow_redistribution(self, redistribution_process="transport:snow_slide", skip_glaciers=False):  # SettingsModel.h:391
        """SettingsModel::AddSnowRedistribution (SettingsModel.cpp:560-572): on the snowpack of every land cover (the
        glacier covers only without skip_glaciers), a lateral process whose instantaneous snow fluxes go to the
        snowpacks of the connected units."""
        if redistribution_process != "transport:snow_slide":
            raise RuntimeError(f"Process type '{redistribution_process}' is not available in the B200 engine.")
        covers = [(b["name"], b["kind"]) for b in self._sel_structure["hydro_unit_bricks"] if b["is_land_cover"]]
        for name, kind in covers:
            if skip_glaciers and kind == "glacier":
                continue
            self.select_hydro_unit_brick(name + "_snowpack")
            self.add_brick_process("snow_redistribution", redistribution_process, "lateral:snow")
            self.set_process_outputs_as_instantaneous()

    # -- parameters ----------------------------------------------------------------------------------
    def set_parameter_value(self, component, name, value):
        """SettingsModel::SetParameterValue (SettingsModel.cpp:923-1031)."""
        if "," in component:
            for c in [c.strip() for c in component.split(",") if c.strip()]:
                if not self.set_parameter_value(c, name, value):
                    return False
            return True
        value = float(np.float32(value))
        found = False

        def set_in_brick(brick):
            for p in brick["parameters"]:
                if p["name"] == name:
                    p["value"] = value
                    return
            for proc in brick["processes"]:
                for p in proc["parameters"]:
                    if p["name"] == name:
                        p["value"] = value
                        return
            raise RuntimeError(f"The parameter '{name}' was not found.")

        for st in self._structures:
            bricks = {b["name"]: b for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]}
            splitters = {s["name"]: s for s in st["hydro_unit_splitters"] + st["sub_basin_splitters"]}
            if component in bricks:
                set_in_brick(bricks[component])
                found = True
            elif component in splitters:
                for p in splitters[component]["parameters"]:
                    if p["name"] == name:
                        p["value"] = value
                        break
                else:
                    raise RuntimeError(f"SettingsModel::SetSplitterParameterValue - Parameter '{name}' not found")
                found = True
            elif "type:" in component:
                kind = component.split(":", 1)[1]
                for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]:
                    if b["kind"] == kind or (kind == "land_cover" and b["kind"] in ("generic_land_cover", "glacier")):
                        set_in_brick(b)
                found = True  # SetParameterValueInSelectedStructure falls through to `return true`
        return found

    # -- export --------------------------------------------------------------------------------------
    def get_structure(self):
        """Binding.cpp:119-160 (plus "parameters"/"log_items" per element)."""
        return copy.deepcopy(self._structures)

    def get_hydro_unit_log_labels(self):
        """SettingsModel::GetHydroUnitLogLabels (SettingsModel.cpp:808-841): union over variants, in order."""
        labels = []
        for st in self._structures:
            for b in st["hydro_unit_bricks"]:
                for item in b["log_items"]:
                    labels.append(f"{b['name']}:{item}")
                for p in b["processes"]:
                    for item in p["log_items"]:
                        labels.append(f"{b['name']}:{p['name']}:{item}")
        return list(dict.fromkeys(labels))

    def get_land_cover_names(self):
        names = []
        for st in self._structures:
            names += [b["name"] for b in st["hydro_unit_bricks"] if b["is_land_cover"]]
        return list(dict.fromkeys(names))


class SettingsBasin:
    """base/SettingsBasin.{h,cpp}."""

    def __init__(self):
        self.units = []
        self.connections = []

    def add_hydro_unit(self, id, area, elevation=-9999):  # noqa: A002
        self.units.append({"id": int(id), "area": float(area), "elevation": float(elevation), "slope": None,
                           "properties": {}, "land_covers": [], "lateral": []})

    def add_land_cover(self, name, kind, fraction):
        self.units[-1]["land_covers"].append((name, float(fraction)))

    def add_hydro_unit_property_str(self, name, value):
        self.units[-1]["properties"][name] = value

    def add_hydro_unit_property_double(self, name, value, unit):
        self.units[-1]["properties"][name] = (float(value), unit)
        if name == "slope":
            self.units[-1]["slope"] = (float(value), unit)

    def add_lateral_connection(self, giver_hydro_unit_id, receiver_hydro_unit_id, fraction, type=""):  # noqa: A002
        """SettingsBasin::AddLateralConnection: kept by unit id; resolved to basin positions when the model is built
        (ModelBuilder.cpp:476-508)."""
        self.connections.append((int(giver_hydro_unit_id), int(receiver_hydro_unit_id), float(fraction), type))

    def get_lateral_connection_count(self):
        return len(self.connections)

    def clear(self):
        self.units = []
        self.connections = []


class SubBasin:
    def __init__(self):
        self.settings = None

    def init(self, spatial_structure):
        if not spatial_structure.units:
            raise ValueError("Basin initialization failed: no hydro unit")
        self.settings = spatial_structure


class TimeSeries:
    def __init__(self, name, time, ids, data):
        self.name, self.time, self.ids, self.data = name, time, ids, data

    @staticmethod
    def create(data_name, time, ids, data):
        return TimeSeries(data_name, np.asarray(time, dtype=np.float64), np.asarray(ids, dtype=np.int32),
                          np.asarray(data, dtype=np.float64))


def _parse_date_mjd(text):
    """UtilsDateTime: 'YYYY-MM-DD' -> modified Julian date."""
    d = np.datetime64(text[:10], "D")
    return float((d - np.datetime64("1858-11-17", "D")).astype(int))


class Action:
    def __init__(self):
        pass


class ModelHydro:
    """base/ModelHydro.{h,cpp} on the CUDA engine. One instance = one structure + basin + period; `run` is
    the reference's single-parameter-set call, `run_batch` integrates many parameter sets in one launch."""

    def __init__(self):
        self._engine = None
        self._cs = None
        self._settings = None
        self._units = None
        self._series = {}
        self._forcing_loaded = False
        self._actions = []
        self._n_steps = 0
        self._labels = []
        self._ran = False
        self._mjd0 = 0.0
        self._device = 0
        self._pending_sets = None
        self._shard = False

    # -- setup ---------------------------------------------------------------------------------------
    def init_with_basin(self, model_settings, basin_settings, device=0):
        try:
            t = model_settings.timer
            if not model_settings.solver:
                raise ValueError("Model settings are not valid.")  # SettingsModel::IsValid
            if t["time_step_unit"] not in ("day", "days", "d"):
                raise _structure.UnsupportedStructure("only daily time steps (time_step_unit='day')")
            self._mjd0 = _parse_date_mjd(t["start"])
            dt = float(t["time_step"])
            # TimeMachine::GetTimeStepCount (TimeMachine.cpp:61-64): 1 + (end - start) / dt
            self._n_steps = int(1 + (_parse_date_mjd(t["end"]) - self._mjd0) / dt)
            self._settings = model_settings
            self._units = [dict(u) for u in basin_settings.units]
            pos = {u["id"]: i for i, u in enumerate(self._units)}
            for u in self._units:
                u["lateral"] = []
            for giver, receiver, fraction, _kind in getattr(basin_settings, "connections", []):
                if giver not in pos or receiver not in pos:
                    raise KeyError(f"lateral connection: hydro unit {giver if giver not in pos else receiver} not found")
                self._units[pos[giver]]["lateral"].append((pos[receiver], fraction))
            self._cs, self._engine = _structure.build_engine(model_settings.get_structure(), self._units,
                                                             model_settings.solver, self._n_steps, device, dt,
                                                             start_mjd=self._mjd0)
            self._engine.set_spinup_steps(int(t["spinup_days"] / dt))
            self._labels = [l for l in model_settings.get_hydro_unit_log_labels()]
            self._device = device
        except (_structure.UnsupportedStructure, KeyError) as err:
            raise ValueError(f"Model initialization failed: {err}") from err

    def is_valid(self):
        return self._engine is not None

    # -- unit shards (one block of a large basin per GPU) -----------------------------------
    def set_unit_shard(self, basin_area_m2, reduce=None):
        """This model holds a block of the units of a larger basin of the given total area (hbk_set_unit_shard).
        reduce(device_pointer, n): in-place sum over all shards (parallel.make_reduce_callback); with it run() covers
        spin-up and dated actions and returns the outlet of the whole basin. Actions are registered with the tables and
        changes of the WHOLE basin on every shard; the entries of units outside this block are skipped."""
        self._engine.set_unit_shard(float(basin_area_m2))
        self._engine.set_shard_reduce(reduce)
        self._shard = basin_area_m2 > 0

    # -- actions (dated state edits, applied between kernel segments) -----------------------------------
    def add_action(self, action):
        self._actions.append(action)
        return True

    def get_action_count(self):
        return len(self._actions)

    def get_sporadic_action_item_count(self):
        return sum(getattr(a, "get_change_count", lambda: 0)() for a in self._actions)

    def _schedule_actions(self):
        """Translate the registered actions into engine events (kernel boundaries): recursive actions in
        registration order, then the sporadic items in date order (ActionsManager.cpp:24-104)."""
        eng = self._engine
        eng.clear_actions()
        if not self._actions:
            return
        dt = self._cs.hbk.time_step_days
        pos = {u["id"]: i for i, u in enumerate(self._units)}
        glacier = self._cs.glacier_name
        layout = self._cs.layout

        def cover_index(name, what):
            """0 for the first glacier cover, 1 for the second (carried by the twin units)."""
            if name == glacier:
                return 0
            if name is not None and name == self._cs.glacier2_name:
                return 1
            raise ValueError(f"{what}: land cover '{name}' is not a glacier cover")

        def engine_unit(u, cover):
            if cover == 0:
                return u
            if layout.twin_col[u] < 0:
                raise ValueError(f"The land cover '{self._cs.glacier2_name}' was not found in hydro unit "
                                 f"{self._units[u]['id']}")
            return layout.twin_col[u]

        events = []  # (step, order, callable)
        sporadic = []
        have_tables = False
        for order, action in enumerate(self._actions):
            if isinstance(action, ActionGlacierSnowToIceTransformation):
                c = cover_index(action.land_cover, "snow-to-ice")
                for k in _schedule.recursive_steps(self._mjd0, self._n_steps, action.month, action.day, dt):
                    events.append((k, order, lambda k=k, c=c: eng.add_snow_to_ice(k, c)))
            elif isinstance(action, ActionGlacierEvolutionDeltaH):
                # ActionGlacierEvolutionAreaScaling shares the lookup-table interface and Init of the delta-h action
                # (ActionGlacierEvolutionAreaScaling.cpp:9-68); only its Apply rule differs
                area_scaling = isinstance(action, ActionGlacierEvolutionAreaScaling)
                if have_tables:
                    raise ValueError("only one glacier-evolution action per model")
                have_tables = True
                c = cover_index(action.land_cover, "delta-h")
                if self._shard:
                    # this block's columns of the lookup tables; the row sums of the whole table, summed in table order
                    # like ActionGlacierEvolutionDeltaH.cpp:123-126, come with them
                    cols = [j for j, i in enumerate(action.hu_ids) if int(i) in pos]
                    units = [pos[int(action.hu_ids[j])] for j in cols]
                    vol = np.asarray(action.volumes, dtype=np.float64)
                    eng.set_delta_h_tables(units, np.asarray(action.areas, dtype=np.float64)[:, cols], vol[:, cols])
                    eng.set_delta_h_shard_totals(np.cumsum(vol, axis=1)[:, -1])
                else:
                    try:
                        units = [engine_unit(pos[int(i)], c) for i in action.hu_ids]
                    except KeyError as err:
                        raise ValueError(f"The hydro unit {err} was not found") from err
                    eng.set_delta_h_tables(units, action.areas, action.volumes)
                for k in _schedule.recursive_steps(self._mjd0, self._n_steps, action.month, 1, dt):
                    events.append((k, order, (lambda k=k: eng.add_area_scaling_update(k)) if area_scaling
                                   else (lambda k=k: eng.add_delta_h_update(k))))
            elif isinstance(action, ActionLandCoverChange):
                for date, unit_id, cover, area in action.changes:
                    sporadic.append((date, order, unit_id, cover, area))
            else:
                raise ValueError(f"Action {type(action).__name__} is not available in the B200 engine")
        for n, (date, order, unit_id, cover, area) in enumerate(sorted(sporadic, key=lambda c: c[0])):
            if unit_id not in pos:
                if self._shard:  # a unit of another block: only the kernel boundary (all shards run the same segments)
                    k = _schedule.sporadic_step(self._mjd0, date, dt)
                    events.append((k, 10000 + n, lambda k=k: eng.add_land_cover_change(k, -1, 0.0)))
                    continue
                raise ValueError(f"The hydro unit {unit_id} was not found")
            u = pos[unit_id]
            if cover not in (glacier, self._cs.glacier2_name) or cover is None:
                # changing the generic cover itself is undone by FixLandCoverFractionsTotal (HydroUnit.cpp:459-490)
                if cover == self._cs.ground_name:
                    continue
                raise ValueError(f"Land cover '{cover}' was not found.")
            fraction = _schedule.check_fraction(area / self._units[u]["area"])
            k = _schedule.sporadic_step(self._mjd0, date, dt)
            ue = engine_unit(u, cover_index(cover, "land cover change"))
            events.append((k, 10000 + n, lambda k=k, u=ue, f=fraction: eng.add_land_cover_change(k, u, f)))
        for _, _, add in sorted(events, key=lambda e: (e[0], e[1])):
            add()

    # -- forcing -------------------------------------------------------------------------------------
    def add_time_series(self, ts):
        return self.create_time_series(ts.name, ts.time, ts.ids, ts.data)

    def create_time_series(self, data_name, time, ids, data):
        """ModelHydro::CreateTimeSeries (ModelHydro.cpp:306-322): data[T x U] float64, one column per id."""
        data = np.asarray(data, dtype=np.float64)
        time = np.asarray(time, dtype=np.float64)
        ids = np.asarray(ids).astype(np.int64).ravel()
        if data.ndim != 2 or data.shape != (len(time), len(ids)):
            return False
        key = data_name.lower()
        if key not in _engine.VAR:
            raise ValueError(f"Unknown forcing variable '{data_name}'")
        self._series[_engine.VAR[key]] = (time, ids, data)
        return True

    def clear_time_series(self):
        self._series = {}
        self._forcing_loaded = False

    def attach_time_series_to_hydro_units(self):
        """ModelHydro::AttachTimeSeriesToHydroUnits (ModelHydro.cpp:324-346): match columns to units by id and
        align the series start with the modelling period (TimeSeriesData cursor, ModelHydro.cpp:348-357)."""
        unit_ids = np.array([u["id"] for u in self._units])
        for var, (time, ids, data) in self._series.items():
            pos = {int(i): k for k, i in enumerate(ids)}
            try:
                cols = [pos[int(i)] for i in unit_ids]
            except KeyError:
                return False
            start = int(round(self._mjd0 - time[0]))
            if start < 0 or start + self._n_steps > len(time):
                raise ValueError("The forcing data do not cover the modelling period.")
            self._engine.set_forcing(var, np.ascontiguousarray(data[start:start + self._n_steps][:, cols]))
        self._forcing_loaded = len(self._series) > 0
        return True

    def forcing_loaded(self):
        return self._forcing_loaded

    def set_station_forcing(self, variable, series, ref_elevation=0.0, gradient=0.0, correction=1.0):
        """Batched extra: station series + fused elevation-gradient spatialisation (forcing.py:1061-1113);
        gradient / correction may be arrays with one value per parameter set (set the parameter sets first)."""
        if np.ndim(gradient) > 0 or np.ndim(correction) > 0:
            if self._pending_sets is None:
                raise ValueError("set_parameter_sets() must precede per-set spatialisation gradients")
            self._engine.set_parameters(self._pending_sets)
        self._engine.set_station_forcing(variable.lower(), series, ref_elevation, gradient, correction)
        self._station = getattr(self, "_station", set()) | {_engine.VAR[variable.lower()]}
        self._forcing_loaded = len(self._station) >= (4 if self._cs.hbk.melt_kind == _engine.MELT_TEMPERATURE_INDEX else 3)

    # -- parameters ----------------------------------------------------------------------------------
    def update_parameters(self, model_settings):
        """ModelHydro::UpdateParameters (ModelHydro.cpp:76-84): re-read every parameter value."""
        cs = _structure.CompiledStructure(model_settings.get_structure(), model_settings.solver,
                                          self._cs.hbk.time_step_days)
        self._cs.params = cs.params
        self._pending_sets = cs.params[None, :].copy()

    def set_parameter_sets(self, values):
        """Batched extra: values[n_sets][HBK_N_PARAMS] float32 in engine.PARAM_SLOTS order."""
        self._pending_sets = np.ascontiguousarray(values, dtype=np.float32)

    def parameter_slot(self, component, name):
        return self._cs.resolve_all(component, name)

    def base_parameters(self):
        """float32[HBK_N_PARAMS]: the current single-set parameter vector (engine.PARAM_SLOTS order)."""
        return self._cs.params.copy()

    def engine_handle(self):
        return self._engine.h.value

    def get_state(self, state_slot):
        return self._engine.get_state(state_slot)

    # -- run -----------------------------------------------------------------------------------------
    def reset(self):
        self._ran = False

    def save_as_initial_state(self):
        self._engine.save_as_initial_state()

    def run(self):
        if not self._forcing_loaded:
            raise ValueError("Time step processing failed: no forcing attached.")
        self._schedule_actions()
        if self._pending_sets is None:
            self._pending_sets = self._cs.params[None, :].copy()
        try:
            self._engine.set_parameters(self._pending_sets)
            slots = [self._cs.labels[l] for l in self._labels if l in self._cs.labels]
            if self._labels and (self._settings._log_all or self._settings._record_fractions):
                slots += ["fraction_ground"] + (["fraction_glacier"] if self._cs.glacier_name else [])
            self._engine.set_record_mask(slots)
            self._engine.run()
        except _engine.HbkError as err:
            raise ValueError(str(err)) from err
        self._ran = True

    def run_batch(self, parameter_sets, record=False):
        """Integrate n parameter sets in one launch; returns outlet[n_sets][T]."""
        self.set_parameter_sets(parameter_sets)
        labels, self._labels = self._labels, (self._labels if record else [])
        try:
            self.run()
        finally:
            self._labels = labels
        return self._engine.get_outlet()

    # -- results -------------------------------------------------------------------------------------
    def get_outlet_discharge(self):
        return self._engine.get_outlet(0, 1)[0]

    def get_outlet_discharge_batch(self):
        return self._engine.get_outlet()

    def get_total_outlet_discharge(self):
        return float(self.get_outlet_discharge().sum())

    def get_recorded_hydro_unit_labels(self):
        return list(self._labels)

    def get_recorded_hydro_unit_fraction_labels(self):
        return self._settings.get_land_cover_names() if self._settings._log_all or self._settings._record_fractions else []

    def get_hydro_unit_values(self, label):
        if label not in self._cs.labels:
            raise RuntimeError(f"Label '{label}' is not recorded.")
        return self._cs.recorded(self._engine, label)

    def get_hydro_unit_fractions(self, label):
        """Logger fraction series (Logger.cpp:111-119): recorded per step, so action-driven changes show."""
        slot = {self._cs.ground_name: "fraction_ground", self._cs.glacier_name: "fraction_glacier",
                self._cs.glacier2_name: "fraction_glacier"}.get(label)
        if slot is None or label is None:
            raise RuntimeError(f"Fraction label '{label}' is not recorded.")
        # the second glacier cover's fraction is the glacier fraction of the twin units
        return self._cs.layout.gather(self._engine.get_recorded(slot, 0), label == self._cs.glacier2_name)

    def get_hydro_unit_ids(self):
        return [u["id"] for u in self._units]

    def get_hydro_unit_areas(self):
        return np.array([u["area"] for u in self._units], dtype=np.float64)

    def _fractions(self):
        """Initial land-cover fractions per ENGINE unit (a twin: no ground, the second glacier cover's fraction)."""
        layout = self._cs.layout
        g = [dict(u["land_covers"]).get(self._cs.ground_name, 1.0) for u in self._units]
        gl = [dict(u["land_covers"]).get(self._cs.glacier_name, 0.0) if self._cs.glacier_name else 0.0
              for u in self._units]
        for u in layout.twin_of[layout.n_basin:]:
            g.append(0.0)
            gl.append(dict(self._units[u]["land_covers"]).get(self._cs.glacier2_name, 0.0))
        return np.array(g), np.array(gl)

    def get_total_et(self):
        """Logger::GetTotalET (Logger.cpp:168-216): area-weighted sum of the ET flux records."""
        areas = self.get_hydro_unit_areas()
        et = self._cs.layout.gather(self._engine.get_recorded("slow_et", 0), False)
        return float((np.nan_to_num(et) * (areas / areas.sum())[None, :]).sum())

    def _storage_total(self, slots_with_fraction):
        areas = self.get_hydro_unit_areas()
        w = self._cs.layout.expand(areas / areas.sum())  # a twin carries a cover of its unit: the unit's weight
        total = 0.0
        for slot, frac in slots_with_fraction:
            end = self._engine.get_state(slot)[0]
            total += float((end * frac * w).sum())
        return total

    def get_total_water_storage_changes(self):
        """Logger::GetTotalWaterStorageChanges (Logger.cpp:220-290): final - initial (initial = 0 here unless
        an initial state was saved) of every water container, land-cover stores weighted by their fraction."""
        g, gl = self._fractions()
        one = np.ones_like(g)
        total = self._storage_total([("ground_water", g), ("glacier_water", gl), ("slow", one), ("slow2", one),
                                     ("quick", one)])
        if self._cs.hbk.snow_retention:  # liquid water kept in the snowpacks: '<cover>_snowpack:water_content'
            total += self._storage_total([("ground_snowpack_water", g), ("glacier_snowpack_water", gl)])
        if self._cs.hbk.has_glacier:
            total += float(self._engine.get_basin_state()[0].sum())
        return total

    def get_total_snow_storage_changes(self):
        g, gl = self._fractions()
        return self._storage_total([("ground_snow", g), ("glacier_snow", gl)])

    def get_total_glacier_storage_changes(self):
        g, gl = self._fractions()
        return self._storage_total([("ice", gl)])

    def dump_outputs(self, path):
        """ModelHydro::DumpOutputs -> ResultWriter::WriteNetCDF (ResultWriter.cpp:8-87): the same variables under the
        same names in <path>/results.nc (netCDF classic, through scipy.io) and <path>/results.npz; read either back with
        results.Results."""
        from . import results

        return results.write_results(self, path, self._cs.assign_structure_variants_ids(
            [u["land_covers"] for u in self._units]))

    def start_mjd(self):
        return self._mjd0

    def time_step_days(self):
        return self._cs.hbk.time_step_days


class ActionLandCoverChange(Action):
    def __init__(self):
        super().__init__()
        self.changes = []

    def add_change(self, date, hydro_unit_id, land_cover, area):
        self.changes.append((float(date), int(hydro_unit_id), land_cover, float(area)))

    def get_change_count(self):
        return len(self.changes)

    def get_land_cover_count(self):
        return len({c[2] for c in self.changes})


class ActionGlacierEvolutionDeltaH(Action):
    def __init__(self):
        super().__init__()
        self.month = 10
        self.land_cover = ""
        self.hu_ids = np.zeros(0, dtype=np.int32)
        self.areas = np.zeros((0, 0))
        self.volumes = np.zeros((0, 0))

    def add_lookup_tables(self, month_num, land_cover, hu_ids, areas, volumes):
        self.month, self.land_cover = int(month_num), land_cover
        self.hu_ids = np.asarray(hu_ids, dtype=np.int32)
        self.areas = np.asarray(areas, dtype=np.float64)
        self.volumes = np.asarray(volumes, dtype=np.float64)

    def get_land_cover_name(self):
        return self.land_cover

    def get_hydro_unit_ids(self):
        return self.hu_ids

    def get_lookup_table_area(self):
        return self.areas

    def get_lookup_table_volume(self):
        return self.volumes


class ActionGlacierEvolutionAreaScaling(ActionGlacierEvolutionDeltaH):
    """Same table interface as the delta-h action (ActionGlacierEvolutionAreaScaling.h:23-24); its Apply rule is
    outside the engine's scope and ModelHydro.run() refuses it."""


class ActionGlacierSnowToIceTransformation(Action):
    def __init__(self, month, day, land_cover):
        super().__init__()
        self.month, self.day, self.land_cover = int(month), int(day), land_cover

    def get_land_cover_name(self):
        return self.land_cover
