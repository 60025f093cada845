"""Python front of the CPU restatement (oracle/hb_oracle.cpp). TEST INFRASTRUCTURE ONLY.

Builds the flat brick/process/flux graph from a structure description exactly the way the
reference's ModelBuilder does (core/src/base/ModelBuilder.cpp:36-240, 448-727) and runs it through
libhboracle.so. The structure description is the list of structure-variant dicts that
`SettingsModel.get_structure()` returns (core/bindings/Binding.cpp:119-160) extended with the
"parameters" of each element (oracle/ref_binding.cpp and hydrobricks_b200 both export them).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import math
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

BRICK_KIND = {"storage": 0, "generic_land_cover": 1, "glacier": 2, "snowpack": 3}
PROC_KIND = {
    "melt:degree_day": 0,
    "et:socont": 1,
    "infiltration:socont": 2,
    "runoff:socont": 3,
    "outflow:linear": 4,
    "outflow:direct": 5,
    "outflow:rest": 6,
    "overflow": 7,
    "outflow:overflow": 7,
    "percolation:constant": 8,
    "transform:snow_ice_constant": 9,  # aliases of Process.cpp:65-66
    "transformation:snow_ice_constant": 9,
    "transform:snow_ice_swat": 10,
    "transformation:snow_ice_swat": 10,
    "sublimation:constant": 11,
    "sublimation:pet": 12,
    "melt:degree_day_aspect": 0,  # the degree-day factor of the unit's aspect class (ProcessMeltDegreeDayAspect.cpp:38-58)
    "melt:temperature_index": 13,
    "transport:snow_slide": 14,
    "outflow:snow_holding": 15,
    "refreeze:degree_day": 16,
    # HBV-96 (oracle only)
    "infiltration:hbv": 17,
    "et:hbv": 18,
    "capillary:hbv": 19,
    "runoff:hbv": 20,
    "routing:hbv": 21,
    # GR4J (oracle only)
    "interception:gr4j": 22,
    "production:gr4j": 23,
    "et:gr4j": 24,
    "routing:gr4j": 25,
    "routing:gr6j": 26,
}
PROC_PARAMS = {
    0: ["degree_day_factor", "melting_temperature"],
    3: ["beta"],
    4: ["response_factor"],
    8: ["percolation_rate"],
    9: ["snow_ice_transformation_rate"],
    10: ["snow_ice_transformation_basal_acc_coeff", "north_hemisphere"],
    11: ["sublimation_rate"],
    12: ["sublimation_pet_factor"],
    13: ["melt_factor", "melting_temperature", "radiation_coefficient"],
    14: ["coeff", "exp", "min_slope", "max_slope", "min_snow_holding_depth", "max_snow_depth"],
    15: ["water_holding_capacity"],
    16: ["refreezing_factor"],
    17: ["beta"],
    18: ["lp", "et_correction_factor"],
    19: ["max_capillary_flux"],
    20: ["response_factor", "alpha"],
    21: ["maxbas"],
    25: ["exchange_factor", "routing_capacity", "uh_base_time"],
    26: ["exchange_factor", "routing_capacity", "uh_base_time", "exchange_threshold", "exp_store_coeff"],
}
ASPECT_DDF = {"north": "degree_day_factor_n", "N": "degree_day_factor_n", "south": "degree_day_factor_s",
              "S": "degree_day_factor_s", "east": "degree_day_factor_ew", "west": "degree_day_factor_ew",
              "E": "degree_day_factor_ew", "W": "degree_day_factor_ew"}
SPLITTER_KIND = {"snow_rain:linear": 0, "snow_rain_transition": 0, "multi_fluxes": 1, "rain": 2, "snow_rain:threshold": 3}
FORCING = {"precipitation": 0, "p": 0, "temperature": 1, "t": 1, "pet": 2,
           "solar_radiation": 3, "r_solar": 3, "r": 3}  # TimeSeries.cpp:134-146
SOLVERS = {"euler_explicit": 0, "heun_explicit": 1, "rk4": 2, "runge_kutta": 2,
           "crank_nicolson": 3, "trapezoidal": 3, "implicit_euler": 4, "euler_implicit": 4, "exponential_euler": 5,
           "analytic_linear": 6, "analytic": 6}
F_TO_BRICK, F_TO_BRICK_INSTANT, F_TO_OUTLET, F_TO_ATMOSPHERE, F_SIMPLE, F_FORCING = range(6)
WATER, SNOW, ICE = 0, 1, 2
CONTENT = {"water": WATER, "snow": SNOW, "ice": ICE}
PRECISION = 1e-8


def build(force: bool = False) -> Path:
    """Compile libhboracle.so (g++ only; seconds)."""
    import os

    if os.environ.get("HBO_LIB"):  # tools/count_flops.py: the instrumented build (make -C oracle count)
        return Path(os.environ["HBO_LIB"])
    so = HERE / "libhboracle.so"
    src = HERE / "hb_oracle.cpp"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, (HERE / "hb_oracle.h").stat().st_mtime):
        subprocess.check_call(["make", "-C", str(HERE), "-B", "libhboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        c_int, c_dbl, c_flt, c_vp = ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_void_p
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        L.hbo_create.restype = c_vp
        L.hbo_create.argtypes = [c_int, c_int, c_dbl, c_int, c_int]
        L.hbo_destroy.argtypes = [c_vp]
        L.hbo_add_brick.argtypes = [c_vp, c_int, c_int, c_int, c_int]
        L.hbo_set_area_fraction.argtypes = [c_vp, c_int, c_dbl]
        L.hbo_add_container.argtypes = [c_vp, c_int, c_int, c_flt, c_int, c_int, c_dbl]
        L.hbo_set_related_snowpack.argtypes = [c_vp, c_int, c_int]
        L.hbo_set_parent.argtypes = [c_vp, c_int, c_int]
        L.hbo_set_target_brick.argtypes = [c_vp, c_int, c_int]
        L.hbo_add_capillary_target.argtypes = [c_vp, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int)]
        L.hbo_add_process.argtypes = [c_vp, c_int, c_int, c_int, fp, c_int, c_int, c_dbl, c_flt]
        L.hbo_add_flux.argtypes = [c_vp, c_int, c_int, c_int, c_dbl, c_int, c_int, c_int]
        L.hbo_process_add_output.argtypes = [c_vp, c_int, c_int]
        L.hbo_container_add_input.argtypes = [c_vp, c_int, c_int]
        L.hbo_add_splitter.argtypes = [c_vp, c_int, c_int, fp, c_int]
        L.hbo_splitter_add_output.argtypes = [c_vp, c_int, c_int]
        L.hbo_splitter_add_input.argtypes = [c_vp, c_int, c_int]
        L.hbo_add_outlet_flux.argtypes = [c_vp, c_int]
        L.hbo_set_forcing.argtypes = [c_vp, c_int, dp]
        L.hbo_set_lateral_slope.argtypes = [c_vp, c_int, ctypes.c_float]
        L.hbo_add_lateral_output.argtypes = [c_vp, c_int, c_int, ctypes.c_double]
        L.hbo_set_unit_area.argtypes = [c_vp, c_int, ctypes.c_double]
        L.hbo_add_record.argtypes = [c_vp, c_int, c_int]
        L.hbo_set_start_mjd.argtypes = [c_vp, c_dbl]
        ipt = ctypes.POINTER(ctypes.c_int)
        L.hbo_add_event.argtypes = [c_vp, c_int, c_int, c_int, c_dbl]
        L.hbo_add_snow_to_ice_unit.argtypes = [c_vp, c_int, c_int]
        L.hbo_set_delta_h.argtypes = [c_vp, c_int, ipt, dp, c_int, dp, dp]
        L.hbo_set_cover_links.argtypes = [c_vp, c_int, c_int, c_int]
        L.hbo_finalize_graph.argtypes = [c_vp]
        L.hbo_run.argtypes = [c_vp, dp, dp]
        L.hbo_last_error.restype = ctypes.c_char_p
        L.hbo_last_error.argtypes = [c_vp]
        L.hbo_unconverged_sweeps.restype = ctypes.c_long
        L.hbo_unconverged_sweeps.argtypes = [c_vp]
        _LIB = L
    return _LIB


def _params(elem) -> dict:
    return {p["name"]: np.float32(p["value"]) for p in elem.get("parameters", [])}


def _farr(values):
    a = (ctypes.c_float * max(len(values), 1))()
    for i, v in enumerate(values):
        a[i] = float(v)
    return a


def slope_m_per_m(value: float, unit: str) -> np.float32:
    """HydroUnitProperty::GetValue(... "m/m") then HydroUnit::GetPropertyFloat (float32 cast).
    HydroUnitProperty.cpp:19-49, HydroUnit.cpp:146-148."""
    if unit in ("m/m", "m_per_m"):
        return np.float32(value)
    if unit in ("degrees", "degree", "deg", "°"):
        return np.float32(math.tan(value * 3.1415926535897932384626433832795 / 180.0))
    raise ValueError(f"unsupported slope unit {unit!r}")


def set_parameter(structures, component: str, name: str, value) -> bool:
    """SettingsModel::SetParameterValue on the exported dicts (SettingsModel.cpp:923-1031)."""
    if "," in component:
        return all(set_parameter(structures, c.strip(), name, value) for c in component.split(",") if c.strip())
    found = False
    value = float(np.float32(value))

    def set_in_brick(brick):
        for p in brick.get("parameters", []):
            if p["name"] == name:
                p["value"] = value
                return True
        for proc in brick["processes"]:
            for p in proc.get("parameters", []):
                if p["name"] == name:
                    p["value"] = value
                    return True
        raise KeyError(f"parameter {name!r} not found in brick {brick['name']!r}")

    for st in structures:
        bricks = {b["name"]: b for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]}
        splitters = {s["name"]: s for s in st["hydro_unit_splitters"] + st["sub_basin_splitters"]}
        if component in bricks:
            set_in_brick(bricks[component])
            found = True
        elif component in splitters:
            for p in splitters[component]["parameters"]:
                if p["name"] == name:
                    p["value"] = value
            found = True
        elif component.startswith("type:"):
            kind = component.split(":", 1)[1]
            for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]:
                if b["kind"] == kind or (kind == "land_cover" and b["kind"] in ("generic_land_cover", "glacier")):
                    set_in_brick(b)
            found = True  # SetParameterValueInSelectedStructure falls through to `return true`
    return found


class OracleModel:
    """One model instance = one parameter set (the values inside `structures`).

    units: list of dicts {"id", "area", "elevation", "slope": (value, unit) or None,
                          "land_covers": [(name, fraction), ...]} in basin order.
    """

    def __init__(self, structures, units, solver: str, n_steps: int, dt: float = 1.0, spinup_steps: int = 0,
                 initial_states: dict | None = None, start_mjd: float = 44605.0):
        self.L = lib()
        self.n_units = len(units)
        self.n_steps = int(n_steps)
        self.h = self.L.hbo_create(self.n_units, self.n_steps, float(dt), SOLVERS[solver], int(spinup_steps))
        self.L.hbo_set_start_mjd(self.h, float(start_mjd))  # 44605 = 1981-01-01
        self.labels = []  # (label, unit index or -1) per record
        self._rec = {}
        self.initial_states = initial_states or {}
        self._build(structures, units)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hbo_destroy(self.h)
            self.h = None

    # -- graph construction (ModelBuilder.cpp) ----------------------------------------------------
    @staticmethod
    def assign_structures(structures, units):
        """ModelBuilder::AssignHydroUnitStructures (ModelBuilder.cpp:36-97)."""
        if len(structures) <= 1:
            return [1] * len(units)
        covers = {st["id"]: {b["name"] for b in st["hydro_unit_bricks"] if b["is_land_cover"]} for st in structures}
        out = []
        for u in units:
            present = {n for n, f in u["land_covers"] if abs(f) > PRECISION}
            matched = next((i for i in sorted(covers) if covers[i] == present), -1)
            if matched < 0:
                supersets = [(len(c), i) for i, c in covers.items() if c >= present]
                if supersets:
                    size = min(s for s, _ in supersets)
                    matched = min(i for s, i in supersets if s == size)
                else:
                    size = max(len(c) for c in covers.values())
                    matched = min(i for i, c in covers.items() if len(c) == size)
            out.append(matched)
        return out

    def _add_brick(self, bs, unit, unit_info):
        L, h = self.L, self.h
        kind = BRICK_KIND[bs["kind"]]
        needs_solver = kind == 0 and not bs.get("computed_directly", False)
        b = L.hbo_add_brick(h, kind, unit, int(needs_solver), -1)
        params = _params(bs)
        key = (bs["name"], unit)
        cap = params.get("capacity", np.float32("nan"))
        init = self.initial_states
        water = L.hbo_add_container(h, b, WATER, float(cap), 0, 0, float(init.get((bs["name"], "water", unit), 0.0)))
        second = -1
        if kind == 3:
            second = L.hbo_add_container(h, b, SNOW, float("nan"), 0, 0, float(init.get((bs["name"], "snow", unit), 0.0)))
        elif kind == 2:
            infinite = "infinite_storage" in params and abs(float(params["infinite_storage"])) > np.finfo(np.float32).eps
            no_melt = "no_melt_when_snow_cover" in params and float(params["no_melt_when_snow_cover"]) > 0
            second = L.hbo_add_container(h, b, ICE, float("nan"), int(infinite), int(no_melt),
                                         float(init.get((bs["name"], "ice", unit), 0.0)))
        info = {"id": b, "kind": kind, "water": water, "second": second, "settings": bs, "processes": [],
                "needs_solver": needs_solver, "unit": unit}
        self._bricks[key] = info
        # Brick forcing connections (ModelBuilder.cpp:693-703)
        for f in bs.get("forcing", []):
            fl = L.hbo_add_flux(h, F_FORCING, 0, 0, 1.0, water, FORCING[f], unit)
            L.hbo_container_add_input(h, water, fl)
        for ps in bs["processes"]:
            ptype = PROC_KIND[ps["kind"]]
            if ptype in (0, 13):
                if kind not in (2, 3):
                    raise ValueError("melt process on unsupported brick")
                container = second
            elif ptype in (9, 10, 11, 12, 14):  # Process.cpp:301-346: the snowpack's snow container
                if kind != 3:
                    raise ValueError("transformation / sublimation process on unsupported brick")
                container = second
            else:
                container = water
            pp = _params(ps)
            if ps["kind"] == "melt:degree_day_aspect":
                aspect = (unit_info or {}).get("aspect_class")
                if aspect not in ASPECT_DDF:
                    raise ValueError(f"Invalid aspect: {aspect}")
                name = ASPECT_DDF[aspect]
                if name.endswith("_ew") and name not in pp:
                    name = "degree_day_factor_we"
                vals = [pp[name], pp["melting_temperature"]]
            else:
                vals = [pp[n] for n in PROC_PARAMS.get(ptype, [])]
            area = unit_info["area"] if unit_info else 0.0
            slope = np.float32(0)
            if ptype == 3:
                slope = slope_m_per_m(*unit_info["slope"])
            p = L.hbo_add_process(h, b, ptype, container, _farr(vals), len(vals), -1, float(area), float(slope))
            info["processes"].append({"id": p, "settings": ps, "type": ptype, "container": container})
        return info

    def _build(self, structures, units):
        L, h = self.L, self.h
        self._bricks = {}
        self._splitters = {}
        self._flux_of = {}
        st_by_id = {st["id"]: st for st in structures}
        st_ids = self.assign_structures(structures, units)
        self.structure_ids = st_ids
        basin_area = 0.0
        for u in units:
            basin_area += u["area"]
        base = st_by_id[1]

        # CreateSubBasinComponents (ModelBuilder.cpp:108-150)
        for bs in base["sub_basin_bricks"]:
            self._add_brick(bs, -1, None)
        self._build_brick_fluxes(base["sub_basin_bricks"], -1, None, basin_area)

        # CreateHydroUnitsComponents, first loop (ModelBuilder.cpp:152-196)
        for iu, u in enumerate(units):
            st = st_by_id[st_ids[iu]]
            hb = st["hydro_unit_bricks"]
            order = [b for b in hb if b["is_surface_component"]] + [b for b in hb if b["is_land_cover"]] + \
                    [b for b in hb if not b["is_surface_component"] and not b["is_land_cover"]]
            for bs in order:
                self._add_brick(bs, iu, u)
            for ss in st["hydro_unit_splitters"]:
                sp = _params(ss)
                kind = SPLITTER_KIND[ss["kind"]]
                if kind == 0:
                    vals = [sp["transition_start"], sp["transition_end"], sp.get("rain_correction_factor", 1.0),
                            sp.get("snow_correction_factor", 1.0)]
                elif kind == 3:
                    vals = [sp["threshold"], 0.0, sp.get("rain_correction_factor", 1.0),
                            sp.get("snow_correction_factor", 1.0)]
                elif kind == 2:
                    vals = [0.0, 0.0, sp.get("rain_correction_factor", 1.0), 1.0]
                else:
                    vals = [0.0, 0.0, 1.0, 1.0]
                s = L.hbo_add_splitter(h, iu, kind, _farr(vals), 4)
                self._splitters[(ss["name"], iu)] = {"id": s, "settings": ss}
            # LinkSurfaceComponentsParents (ModelBuilder.cpp:294-306) + Glacier::SurfaceComponentAdded
            for bs in hb:
                if bs["is_surface_component"] and bs.get("parent"):
                    child = self._bricks[(bs["name"], iu)]
                    parent = self._bricks[(bs["parent"], iu)]
                    child["parent"] = parent
                    self._set_parent(child, parent)

        # links used by the land-cover-change actions (HydroUnit::GetGenericLandCover, '<cover>_snowpack')
        for iu, u in enumerate(units):
            st = st_by_id[st_ids[iu]]
            covers = [b["name"] for b in st["hydro_unit_bricks"] if b["is_land_cover"]]
            generic = next((n for n in ("open", "ground", "generic", "generic_land_cover") if n in covers), None)
            for name in covers:
                sp = self._bricks.get((name + "_snowpack", iu))
                L.hbo_set_cover_links(h, self._bricks[(name, iu)]["id"],
                                      self._bricks[(generic, iu)]["id"] if generic else -1, sp["id"] if sp else -1)
        self._units = units

        # Land-cover fractions (SubBasin::AssignFractions, SubBasin.cpp:57-101)
        for iu, u in enumerate(units):
            for name, fraction in u["land_covers"]:
                if (name, iu) in self._bricks:
                    L.hbo_set_area_fraction(h, self._bricks[(name, iu)]["id"], float(fraction))
                    self._bricks[(name, iu)]["fraction"] = float(fraction)

        for iu, u in enumerate(units):
            L.hbo_set_unit_area(h, iu, float(u["area"]))
        # second loop (ModelBuilder.cpp:198-204)
        for iu, u in enumerate(units):
            st = st_by_id[st_ids[iu]]
            self._build_brick_fluxes(st["hydro_unit_bricks"], iu, u, basin_area)
            self._build_splitter_fluxes(st["hydro_unit_splitters"], iu, u, basin_area)
        L.hbo_finalize_graph(h)

    def _set_parent(self, child, parent):
        self.L.hbo_set_parent(self.h, child["id"], parent["id"])
        if parent["kind"] == 2 and child["kind"] == 3:
            self.L.hbo_set_related_snowpack(self.h, parent["second"], child["id"])

    def _find_brick(self, name, unit):
        if (name, unit) in self._bricks:
            return self._bricks[(name, unit)], False
        if (name, -1) in self._bricks:
            return self._bricks[(name, -1)], unit >= 0
        return None, False

    def _attach_in(self, brick, flux, content):
        """Brick::AttachFluxIn / Snowpack::AttachFluxIn / Glacier::AttachFluxIn."""
        if content == WATER:
            c = brick["water"]
        elif content == SNOW and brick["kind"] == 3:
            c = brick["second"]
        elif content == ICE and brick["kind"] == 2:
            c = brick["second"]
        else:
            raise ValueError("flux content type not accepted by the target brick")
        self.L.hbo_container_add_input(self.h, c, flux)
        return c

    def _build_brick_fluxes(self, brick_settings, unit, u, basin_area):
        """BuildSubBasinBricksFluxes / BuildHydroUnitBricksFluxes (ModelBuilder.cpp:399-578)."""
        L, h = self.L, self.h
        can_have_fraction = lambda b: b["kind"] in (1, 2, 3)  # noqa: E731
        for bs in brick_settings:
            brick = self._bricks[(bs["name"], unit)]
            for proc in brick["processes"]:
                ps = proc["settings"]
                # LinkHydroUnitProcessesTargetBricks (infiltration needs its target)
                if proc["type"] in (2, 17):
                    target, _ = self._find_brick(ps["outputs"][0]["target"], unit)
                    L.hbo_set_target_brick(h, proc["id"], target["id"])
                if proc["type"] == 19:
                    # capillary:hbv links every output's target with the land covers that feed it
                    # (ModelBuilder::LinkHydroUnitProcessesTargetBricks / FindLandCoversFeeding, ModelBuilder.cpp:340-348)
                    for out in ps["outputs"]:
                        target, _ = self._find_brick(out["target"], unit)
                        feeding = [self._bricks[(b["name"], unit)]["id"] for b in brick_settings if b["is_land_cover"] and
                                   any(o["target"] == out["target"] for pp in b["processes"] for o in pp["outputs"])]
                        arr = (ctypes.c_int * max(len(feeding), 1))(*feeding)
                        L.hbo_add_capillary_target(h, proc["id"], target["id"], len(feeding), arr)
                if proc["type"] in (1, 11, 12, 18, 22, 24):  # ToAtmosphere (ET, sublimation, interception)
                    fl = L.hbo_add_flux(h, F_TO_ATMOSPHERE, 0, 0, 1.0, -1, -1, unit)
                    L.hbo_process_add_output(h, proc["id"], fl)
                    self._flux_of[(bs["name"], ps["name"], unit)] = fl
                    continue
                for out in ps["outputs"]:
                    content = CONTENT[out["flux_type"]]
                    if out["target"] == "lateral":
                        # ModelBuilder.cpp:476-508: one flux per (lateral connection of the unit, snowpack of the
                        # receiving unit), weighted with the connection's fraction
                        if content != SNOW:
                            raise ValueError("Lateral flux type not yet supported")
                        value, slope_unit = u["slope"]
                        if slope_unit not in ("deg", "degree", "degrees", "°"):
                            raise ValueError("snow slide: slope in degrees expected")
                        L.hbo_set_lateral_slope(h, proc["id"], float(np.float32(value)))
                        for receiver, fraction in u.get("lateral", []):
                            for (bname, bunit), tb in self._bricks.items():
                                if bunit != receiver or tb["kind"] != 3:
                                    continue
                                kind = F_TO_BRICK_INSTANT if out["instantaneous"] else F_TO_BRICK
                                fl = L.hbo_add_flux(h, kind, int(out["static"]), 1, 1.0, tb["second"], -1, unit)
                                self._attach_in(tb, fl, SNOW)
                                L.hbo_process_add_output(h, proc["id"], fl)
                                L.hbo_add_lateral_output(h, proc["id"], tb["id"], float(fraction))
                        continue
                    if out["target"] == "outlet":
                        if unit >= 0:
                            fl = L.hbo_add_flux(h, F_TO_OUTLET, 1, int(can_have_fraction(brick)),
                                                u["area"] / basin_area, -1, -1, unit)
                        else:
                            fl = L.hbo_add_flux(h, F_TO_OUTLET, 0, 0, 1.0, -1, -1, unit)
                        L.hbo_add_outlet_flux(h, fl)
                    else:
                        target, to_sub = self._find_brick(out["target"], unit)
                        if target is not None:
                            kind = F_TO_BRICK_INSTANT if out["instantaneous"] else F_TO_BRICK
                            weighting = unit >= 0 and can_have_fraction(brick) and not can_have_fraction(target)
                            frac = u["area"] / basin_area if to_sub else 1.0
                            static = to_sub or (unit >= 0 and out["static"])
                            tc = target["water"] if content == WATER else target["second"]
                            fl = L.hbo_add_flux(h, kind, int(static), int(weighting), frac, tc, -1, unit)
                            self._attach_in(target, fl, content)
                        elif (out["target"], unit) in self._splitters:
                            fl = L.hbo_add_flux(h, F_SIMPLE, 1, int(can_have_fraction(brick)), 1.0, -1, -1, unit)
                            L.hbo_splitter_add_input(h, self._splitters[(out["target"], unit)]["id"], fl)
                        else:
                            raise ValueError(f"The target {out['target']} to attach the flux was no found")
                    L.hbo_process_add_output(h, proc["id"], fl)
                    self._flux_of[(bs["name"], ps["name"], unit)] = fl

    def _build_splitter_fluxes(self, splitter_settings, unit, u, basin_area):
        """BuildHydroUnitSplittersFluxes (ModelBuilder.cpp:622-691)."""
        L, h = self.L, self.h
        for ss in splitter_settings:
            sp = self._splitters[(ss["name"], unit)]
            for out in ss["outputs"]:
                content = CONTENT[out["flux_type"]]
                if out["target"] == "outlet":
                    fl = L.hbo_add_flux(h, F_TO_OUTLET, 0, 0, u["area"] / basin_area, -1, -1, unit)
                    L.hbo_add_outlet_flux(h, fl)
                else:
                    target, to_sub = self._find_brick(out["target"], unit)
                    if target is not None:
                        frac = u["area"] / basin_area if to_sub else 1.0
                        tc = target["water"] if content == WATER else target["second"]
                        fl = L.hbo_add_flux(h, F_TO_BRICK, 1, 0, frac, tc, -1, unit)
                        self._attach_in(target, fl, content)
                    elif (out["target"], unit) in self._splitters:
                        fl = L.hbo_add_flux(h, F_SIMPLE, 1, 0, 1.0, -1, -1, unit)
                        L.hbo_splitter_add_input(h, self._splitters[(out["target"], unit)]["id"], fl)
                    else:
                        raise ValueError(f"The target {out['target']} to attach the flux was no found")
                L.hbo_splitter_add_output(h, sp["id"], fl)

    # -- dated actions (applied before time step `step`, step >= 1) --------------------------------
    def add_land_cover_change(self, step, unit_index, cover, fraction):
        b = self._bricks[(cover, unit_index)]
        self.L.hbo_add_event(self.h, int(step), 0, b["id"], float(fraction))

    def add_snow_to_ice(self, cover, steps):
        if not getattr(self, "_s2i_done", False):
            for iu in range(self.n_units):
                if (cover, iu) in self._bricks:
                    self.L.hbo_add_snow_to_ice_unit(self.h, self._bricks[(cover, iu)]["id"],
                                                    self._bricks[(cover + "_snowpack", iu)]["id"])
            self._s2i_done = True
        for s in steps:
            self.L.hbo_add_event(self.h, int(s), 1, -1, 0.0)

    def set_delta_h(self, cover, unit_indices, areas, volumes, steps, area_scaling=False):
        areas = np.ascontiguousarray(areas, dtype=np.float64)
        volumes = np.ascontiguousarray(volumes, dtype=np.float64)
        n = len(unit_indices)
        bricks = (ctypes.c_int * n)(*[self._bricks[(cover, iu)]["id"] for iu in unit_indices])
        ua = np.ascontiguousarray([self._units[iu]["area"] for iu in unit_indices], dtype=np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        self.L.hbo_set_delta_h(self.h, n, bricks, ua.ctypes.data_as(dp), areas.shape[0], areas.ctypes.data_as(dp),
                               volumes.ctypes.data_as(dp))
        for s in steps:  # ActionGlacierEvolutionAreaScaling shares the tables and Init of the delta-h action
            self.L.hbo_add_event(self.h, int(s), 3 if area_scaling else 2, -1, 0.0)

    # -- recording (Logger.cpp:93-120; label scheme SettingsModel.cpp:808-841) --------------------
    def record(self, label: str):
        """label: 'brick:water_content' | 'brick:snow_content' | 'brick:ice_content' |
        'brick:process:output'. Registers one record per unit that has it; sub-basin bricks -> unit -1."""
        parts = label.split(":")
        recs = []
        units = list(range(self.n_units)) + [-1]
        for iu in units:
            b = self._bricks.get((parts[0], iu))
            if b is None:
                continue
            if len(parts) == 2:
                item = parts[1]
                if item in ("water_content", "water"):
                    rid = self.L.hbo_add_record(self.h, 0, b["water"])
                elif item in ("snow_content", "snow", "ice_content", "ice"):
                    rid = self.L.hbo_add_record(self.h, 0, b["second"])
                else:
                    raise KeyError(label)
            else:
                rid = self.L.hbo_add_record(self.h, 1, self._flux_of[(parts[0], parts[1], iu)])
            recs.append((iu, rid))
        self._rec[label] = recs
        return recs

    def set_forcing(self, var: str, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        assert a.shape == (self.n_steps, self.n_units), a.shape
        self.L.hbo_set_forcing(self.h, FORCING[var], a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    def run(self):
        n_rec = sum(len(v) for v in self._rec.values())
        outlet = np.zeros(self.n_steps)
        rec = np.zeros((max(n_rec, 1), self.n_steps))
        dp = ctypes.POINTER(ctypes.c_double)
        rc = self.L.hbo_run(self.h, outlet.ctypes.data_as(dp), rec.ctypes.data_as(dp))
        if rc != 0:
            raise RuntimeError(self.L.hbo_last_error(self.h).decode())
        self.outlet = outlet
        self.records = {}
        for label, recs in self._rec.items():
            hu = np.full((self.n_steps, self.n_units), np.nan)
            sub = None
            for iu, rid in recs:
                if iu >= 0:
                    hu[:, iu] = rec[rid]
                else:
                    sub = rec[rid].copy()
            self.records[label] = sub if sub is not None else hu
        self.unconverged = self.L.hbo_unconverged_sweeps(self.h)
        return outlet
