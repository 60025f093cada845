"""Structure compiler: model-structure description -> the engine's per-structure process table.

Input is the list of structure-variant dicts that `SettingsModel.get_structure()` exports
(reference core/bindings/Binding.cpp:119-160, with the parameter values of every element). The compiler
recognises the Socont family (python/src/hydrobricks/models/socont.py:111-215, modules/glacier.py:87-131,
core/tests/src/helpers.cpp:9-107) by the ROLE of each brick, not by its name:

  snow_rain_transition splitter (snow_rain:linear | snow_rain:threshold) -> rain / snow multi_fluxes splitters
  one generic land cover     {infiltration:socont -> S, outflow:rest -> Q}            (direct brick)
  optional glacier land cover {outflow:direct -> B1 (instantaneous), melt:<kind> -> B2 (instantaneous)}
  one snowpack per cover      {melt:<kind> -> cover}      <kind> = degree_day | degree_day_aspect | temperature_index,
                                                          the same on every snowpack and on the ice (socont.py:141)
                              [+ transform:snow_ice_* on a glacier] [+ transport:snow_slide -> lateral] [+ sublimation:*]
  S  storage, capacity        {et:socont, outflow:linear -> outlet, overflow -> outlet, [percolation:constant -> S2]}
  S2 storage                  {outflow:linear -> outlet}
  Q  storage                  {runoff:socont -> outlet | outflow:linear -> outlet}
  B1, B2 sub-basin storages   {outflow:linear -> outlet}

and emits (i) the hbk_structure flags that select the kernel instantiation, (ii) the float32 parameter
vector in HBK_P_* slot order, (iii) the label -> recorder-slot map. Anything else raises
UnsupportedStructure: the engine never silently computes something different from the reference.
"""
from __future__ import annotations

import math

import numpy as np

from . import engine as _e

PRECISION = 1e-8


class UnsupportedStructure(ValueError):
    pass


def _params(elem):
    return {p["name"]: np.float32(p["value"]) for p in elem.get("parameters", [])}


def _proc_kinds(brick):
    return [p["kind"] for p in brick["processes"]]


def _target(proc):
    outs = proc.get("outputs", [])
    return outs[0]["target"] if outs else None


class CompiledStructure:
    def __init__(self, structures, solver: str, time_step_days: float = 1.0, start_mjd: float = 44605.0):
        if solver not in _e.SOLVER:
            raise UnsupportedStructure(
                f"Incorrect solver name: {solver!r}. Valid solver names: " + ", ".join(_e.SOLVER))
        self.solver = solver
        self.structures = structures
        self.labels = {}
        self.params = np.zeros(_e.N_PARAMS, dtype=np.float32)
        self.param_index = {}  # (component, parameter name) -> slot
        self.start_mjd = float(start_mjd)
        self._compile(structures, time_step_days)

    # ------------------------------------------------------------------------------------------------
    def _set(self, slot, component, name, value):
        i = _e.PARAM_SLOTS.index(slot)
        self.params[i] = np.float32(value)
        self.param_index[(component, name)] = i

    def _compile(self, structures, dt):
        if not structures:
            raise UnsupportedStructure("empty structure")
        by_id = {s["id"]: s for s in structures}
        base = by_id[1]
        glacier_st = None
        for s in structures:
            if any(b["kind"] == "glacier" for b in s["hydro_unit_bricks"]):
                glacier_st = s
        if len(structures) > 2:
            raise UnsupportedStructure("more than two structure variants")
        full = glacier_st or base  # the variant that carries every brick
        self.glacier_structure_id = glacier_st["id"] if glacier_st else None
        self.variant_covers = {s["id"]: {b["name"] for b in s["hydro_unit_bricks"] if b["is_land_cover"]}
                               for s in structures}

        bricks = {b["name"]: b for b in full["hydro_unit_bricks"]}
        basin_bricks = {b["name"]: b for b in base["sub_basin_bricks"]}
        splitters = {s["name"]: s for s in full["hydro_unit_splitters"]}

        st = _e.HbkStructure()
        st.solver = _e.SOLVER[self.solver]
        st.time_step_days = float(dt)
        st.snow_ice_kind = _e.SNOW_ICE_NONE
        st.sublimation_kind = _e.SUBLIMATION_NONE
        st.north_hemisphere = 1
        st.start_mjd = self.start_mjd

        # --- splitters
        if base["sub_basin_splitters"]:
            raise UnsupportedStructure("sub-basin splitters")
        tr = [s for s in splitters.values() if s["kind"] in ("snow_rain:linear", "snow_rain:threshold")]
        multi = [s for s in splitters.values() if s["kind"] == "multi_fluxes"]
        rain_only = [s for s in splitters.values() if s["kind"] == "rain"]
        self.no_snow = len(rain_only) == 1 and not tr
        if self.no_snow:
            # SettingsModel::GeneratePrecipitationSplitters(withSnow = false) (SettingsModel.cpp:518-529): SplitterRain
            # (SplitterRain.cpp:37-39) sends precipitation x rain_correction_factor to the land covers; no snowpacks.
            # On the device this is the threshold splitter with a threshold no temperature reaches (-FLT_MAX): rain =
            # precipitation x rcf, snow = 0, and the snowpack code only ever moves zeros.
            if len(multi) != 1 or len(splitters) != 2:
                raise UnsupportedStructure("expected the rain splitter + the rain multi_fluxes splitter")
            tr = rain_only[0]
            st.splitter_kind = _e.SPLIT_THRESHOLD
            self.params[_e.PARAM_SLOTS.index("transition_start")] = -np.finfo(np.float32).max
            self.params[_e.PARAM_SLOTS.index("snow_correction")] = 1.0
            self._set("rain_correction", tr["name"], "rain_correction_factor",
                      _params(tr).get("rain_correction_factor", 1.0))
            self.splitter_name = tr["name"]
            tgt = splitters.get(tr["outputs"][0]["target"]) if len(tr["outputs"]) == 1 else None
            if tgt is None or tgt["kind"] != "multi_fluxes":
                raise UnsupportedStructure("the rain splitter must feed a multi_fluxes splitter")
            rain_targets, snow_targets = [o["target"] for o in tgt["outputs"]], []
        else:
            if len(tr) != 1 or len(multi) != 2 or len(splitters) != 3:
                raise UnsupportedStructure("expected snow_rain_transition + rain/snow multi_fluxes splitters")
            tr = tr[0]
            sp = _params(tr)
            if tr["kind"] == "snow_rain:linear":
                st.splitter_kind = _e.SPLIT_LINEAR
                self._set("transition_start", tr["name"], "transition_start", sp["transition_start"])
                self._set("transition_end", tr["name"], "transition_end", sp["transition_end"])
            else:
                st.splitter_kind = _e.SPLIT_THRESHOLD
                self._set("transition_start", tr["name"], "threshold", sp["threshold"])
            self._set("rain_correction", tr["name"], "rain_correction_factor", sp.get("rain_correction_factor", 1.0))
            self._set("snow_correction", tr["name"], "snow_correction_factor", sp.get("snow_correction_factor", 1.0))
            self.splitter_name = tr["name"]
            rain_targets = snow_targets = None
            for out in tr["outputs"]:
                tgt = splitters.get(out["target"])
                if tgt is None or tgt["kind"] != "multi_fluxes":
                    raise UnsupportedStructure("snow_rain_transition must feed multi_fluxes splitters")
                names = [o["target"] for o in tgt["outputs"]]
                if out["flux_type"] == "snow":
                    snow_targets = names
                else:
                    rain_targets = names

        # --- land covers
        covers = [b for b in full["hydro_unit_bricks"] if b["is_land_cover"]]
        ground = [b for b in covers if b["kind"] == "generic_land_cover"]
        glacier = [b for b in covers if b["kind"] == "glacier"]
        if len(ground) != 1 or len(glacier) > 2 or len(ground) + len(glacier) != len(covers):
            raise UnsupportedStructure("expected one generic land cover and at most two glacier covers")
        ground = ground[0]
        # a second glacier cover (glacier_ice + glacier_debris, models/socont.py:131-140) is carried by twin units
        # (hbk_set_unit_twins): the same bricks once more with the HBK_P_GLACIER2_* / HBK_P_ICE2_* parameter slots
        glacier2 = glacier[1] if len(glacier) > 1 else None
        glacier = glacier[0] if glacier else None
        self.ground_name = ground["name"]
        self.glacier_name = glacier["name"] if glacier else None
        self.glacier2_name = glacier2["name"] if glacier2 else None
        self.twin_labels = set()  # log labels of the second glacier cover: recorded in the twin units' columns
        # rain goes to every land cover — or, with rain_to_snowpack, to the snowpack of every land cover (checked once
        # the snowpacks are known, SettingsModel.cpp:624-632)
        rain_to_snowpack = sorted(rain_targets or []) == sorted(c["name"] + "_snowpack" for c in covers)
        if not rain_to_snowpack and sorted(rain_targets or []) != sorted(c["name"] for c in covers):
            raise UnsupportedStructure("rain must go to every land cover")
        if _proc_kinds(ground) != ["infiltration:socont", "outflow:rest"]:
            raise UnsupportedStructure(f"land cover {ground['name']}: expected infiltration:socont + outflow:rest")
        slow_name = _target(ground["processes"][0])
        quick_name = _target(ground["processes"][1])
        if not ground["processes"][1]["outputs"][0]["static"]:
            raise UnsupportedStructure("outflow:rest output must be static")
        g = self.ground_name
        self.labels.update({f"{g}:water_content": "ground_water",
                            f"{g}:{ground['processes'][0]['name']}:output": "ground_infiltration",
                            f"{g}:{ground['processes'][1]['name']}:output": "ground_runoff"})

        # --- snowpacks
        transfo_kinds = {"transform:snow_ice_constant": _e.SNOW_ICE_CONSTANT,
                         "transformation:snow_ice_constant": _e.SNOW_ICE_CONSTANT,
                         "transform:snow_ice_swat": _e.SNOW_ICE_SWAT, "transformation:snow_ice_swat": _e.SNOW_ICE_SWAT}

        sublim_kinds = {"sublimation:constant": _e.SUBLIMATION_CONSTANT, "sublimation:pet": _e.SUBLIMATION_PET}
        st.melt_kind = -1
        st.snow_slide_kind = _e.SNOW_SLIDE_NONE
        slide_on = {}  # snowpack name -> its transport:snow_slide process

        def set_melt(component, proc, slot_a, slot_tm, slot_b, slot_c):
            """The melt process of a snowpack or of the glacier ice: its kind (one per structure) and its parameters —
            degree_day_factor | degree_day_factor_n, _s, _ew (or _we, ProcessMeltDegreeDayAspect.cpp:48-55) |
            melt_factor, radiation_coefficient."""
            kind = _e.MELT_KIND.get(proc["kind"])
            if kind is None:
                raise UnsupportedStructure(f"{component}: melt process {proc['kind']}")
            if st.melt_kind not in (-1, kind):
                raise UnsupportedStructure("different melt processes on the snowpacks / the glacier ice")
            st.melt_kind = kind
            pp = _params(proc)
            self._set(slot_tm, component, "melting_temperature", pp["melting_temperature"])
            if kind == _e.MELT_DEGREE_DAY:
                self._set(slot_a, component, "degree_day_factor", pp["degree_day_factor"])
            elif kind == _e.MELT_DEGREE_DAY_ASPECT:
                self._set(slot_a, component, "degree_day_factor_n", pp["degree_day_factor_n"])
                self._set(slot_b, component, "degree_day_factor_s", pp["degree_day_factor_s"])
                ew = "degree_day_factor_ew" if "degree_day_factor_ew" in pp else "degree_day_factor_we"
                if ew not in pp:
                    raise UnsupportedStructure("Missing parameter 'degree_day_factor_ew' or 'degree_day_factor_we'")
                self._set(slot_c, component, ew, pp[ew])
            else:
                self._set(slot_a, component, "melt_factor", pp["melt_factor"])
                self._set(slot_b, component, "radiation_coefficient", pp["radiation_coefficient"])

        retention = {}  # snowpack name -> (outflow:snow_holding process, refreeze:degree_day process or None)

        def snowpack_of(cover):
            """One snowpack per cover: melt:degree_day [, outflow:snow_holding [, refreeze:degree_day]] [, snow -> ice
            transformation on a glacier] [, snow slide] [, sublimation], in the order model_settings.py:248-263 adds them.
            Returns (brick, transformation process, sublimation process)."""
            sps = [b for b in full["hydro_unit_bricks"] if b["kind"] == "snowpack" and b.get("parent") == cover["name"]]
            ok = len(sps) == 1 and sps[0]["processes"] and sps[0]["processes"][0]["kind"] in _e.MELT_KIND
            transfo = sublim = None
            if ok:
                sp_name = sps[0]["name"]
                melt_out = sps[0]["processes"][0]["outputs"]
                rest = list(sps[0]["processes"][1:])
                if _target(sps[0]["processes"][0]) == sp_name:
                    # SettingsModel::GenerateSnowpacksWithWaterRetention (SettingsModel.cpp:608-634): the melt stays in the
                    # snowpack (instantaneous flux to its own water container), outflow:snow_holding feeds the cover
                    ok = len(melt_out) == 1 and melt_out[0]["instantaneous"] and melt_out[0]["flux_type"] == "water" and \
                        bool(rest) and rest[0]["kind"] == "outflow:snow_holding"
                    if ok:
                        holding = rest.pop(0)
                        out = holding["outputs"]
                        ok = len(out) == 1 and out[0]["target"] == cover["name"] and out[0]["flux_type"] == "water" and \
                            not out[0]["instantaneous"] and not out[0]["static"]
                        refreeze = None
                        if ok and rest and rest[0]["kind"] == "refreeze:degree_day":
                            # SettingsModel::AddSnowpackRefreezing (SettingsModel.cpp:574-590)
                            refreeze = rest.pop(0)
                            out = refreeze["outputs"]
                            ok = len(out) == 1 and out[0]["target"] == sp_name and out[0]["flux_type"] == "snow" and \
                                out[0]["instantaneous"]
                        retention[sp_name] = (holding, refreeze)
                else:
                    ok = _target(sps[0]["processes"][0]) == cover["name"]
                if ok and rest and rest[0]["kind"] in transfo_kinds and cover["kind"] == "glacier":
                    # SettingsModel::AddSnowIceTransformation (SettingsModel.cpp:547-559): snow -> the glacier's ice
                    transfo = rest.pop(0)
                    out = transfo["outputs"]
                    ok = len(out) == 1 and out[0]["target"] == cover["name"] and out[0]["flux_type"] == "ice" and \
                        not out[0]["instantaneous"] and not out[0]["static"]
                if ok and rest and rest[0]["kind"] == "transport:snow_slide":
                    # SettingsModel::AddSnowRedistribution (SettingsModel.cpp:560-572): instantaneous snow fluxes to
                    # the snowpacks of the laterally connected units
                    slide = rest.pop(0)
                    out = slide["outputs"]
                    ok = len(out) == 1 and out[0]["target"] == "lateral" and out[0]["flux_type"] == "snow" and \
                        out[0]["instantaneous"]
                    slide_on[sps[0]["name"]] = slide
                if ok and rest and rest[0]["kind"] in sublim_kinds:
                    # SettingsModel::AddSnowpackSublimation (SettingsModel.cpp:586-598): no target = to the atmosphere
                    sublim = rest.pop(0)
                    ok = not sublim["outputs"]
                ok = ok and not rest
            if not ok:
                raise UnsupportedStructure(f"cover {cover['name']}: expected one snowpack with melt:degree_day"
                                           "|degree_day_aspect|temperature_index [+ outflow:snow_holding [+ refreeze:"
                                           "degree_day]] [+ transform:snow_ice_constant|swat on a glacier] "
                                           "[+ sublimation:constant|pet]")
            return sps[0], transfo, sublim

        def set_retention(sp, whc_slot, cfr_slot, prefix):
            """water_holding_capacity / refreezing_factor of a snowpack with liquid-water retention and its labels."""
            if sp["name"] not in retention:
                return
            holding, refreeze = retention[sp["name"]]
            self._set(whc_slot, sp["name"], "water_holding_capacity", _params(holding)["water_holding_capacity"])
            self.labels[f"{sp['name']}:{holding['name']}:output"] = prefix + "_meltwater"
            if refreeze is not None:
                self._set(cfr_slot, sp["name"], "refreezing_factor", _params(refreeze)["refreezing_factor"])
                self.labels[f"{sp['name']}:{refreeze['name']}:output"] = prefix + "_refreeze"

        def set_sublimation(sp, proc, slot, record):
            kind = sublim_kinds[proc["kind"]]
            if st.sublimation_kind not in (_e.SUBLIMATION_NONE, kind):
                raise UnsupportedStructure("different sublimation processes on the snowpacks")
            st.sublimation_kind = kind
            pname = "sublimation_rate" if kind == _e.SUBLIMATION_CONSTANT else "sublimation_pet_factor"
            self._set(slot, sp["name"], pname, _params(proc)[pname])
            self.labels[f"{sp['name']}:{proc['name']}:output"] = record

        gsp = gsub = None
        snowpacks = []
        if not self.no_snow:
            gsp, _, gsub = snowpack_of(ground)
            if gsub is not None:
                set_sublimation(gsp, gsub, "ground_sublimation", "ground_snowpack_sublimation")
            set_melt(gsp["name"], gsp["processes"][0], "ground_snow_ddf", "ground_snow_tmelt", "ground_snow_melt_b",
                     "ground_snow_melt_c")
            self.labels.update({f"{gsp['name']}:water_content": "ground_snowpack_water",
                                f"{gsp['name']}:snow_content": "ground_snowpack_snow",
                                f"{gsp['name']}:{gsp['processes'][0]['name']}:output": "ground_snowpack_melt"})
            snowpacks = [gsp["name"]]
            set_retention(gsp, "ground_snow_whc", "ground_snow_cfr", "ground_snowpack")

        # --- glacier
        st.has_glacier = 0
        b1 = b2 = None
        for second, glacier_cover in enumerate((glacier, glacier2)):
            if glacier_cover is None:
                continue
            sfx = "2" if second else ""
            slots = {"ice_ddf": f"ice{sfx}_ddf", "ice_tmelt": f"ice{sfx}_tmelt", "ice_melt_b": f"ice{sfx}_melt_b",
                     "ice_melt_c": f"ice{sfx}_melt_c", "glacier_snow_ddf": f"glacier{sfx}_snow_ddf",
                     "glacier_snow_tmelt": f"glacier{sfx}_snow_tmelt", "glacier_snow_melt_b": f"glacier{sfx}_snow_melt_b",
                     "glacier_snow_melt_c": f"glacier{sfx}_snow_melt_c", "glacier_sublimation": f"glacier{sfx}_sublimation",
                     "snow_ice": f"snow_ice{sfx}"}

            def label(name, slot):
                self.labels[name] = slot
                if second:
                    self.twin_labels.add(name)

            st.has_glacier = 1
            gk = _proc_kinds(glacier_cover)
            if len(gk) != 2 or gk[0] != "outflow:direct" or gk[1] not in _e.MELT_KIND:
                raise UnsupportedStructure("glacier cover: expected outflow:direct + melt:degree_day|degree_day_aspect|"
                                           "temperature_index")
            for proc in glacier_cover["processes"]:
                if not proc["outputs"][0]["instantaneous"]:
                    raise UnsupportedStructure("glacier outputs must be instantaneous")
            targets = (_target(glacier_cover["processes"][0]), _target(glacier_cover["processes"][1]))
            gp = _params(glacier_cover)
            infinite = int("infinite_storage" in gp and abs(float(gp["infinite_storage"])) > np.finfo(np.float32).eps)
            no_melt = int("no_melt_when_snow_cover" in gp and float(gp["no_melt_when_snow_cover"]) > 0)
            if second:
                if targets != (b1, b2) or infinite != st.glacier_infinite_storage or \
                        no_melt != st.glacier_no_melt_when_snow_cover:
                    raise UnsupportedStructure("the glacier covers must share their options and their sub-basin stores")
            else:
                b1, b2 = targets
                st.glacier_infinite_storage, st.glacier_no_melt_when_snow_cover = infinite, no_melt
            set_melt(glacier_cover["name"], glacier_cover["processes"][1], slots["ice_ddf"], slots["ice_tmelt"],
                     slots["ice_melt_b"], slots["ice_melt_c"])
            n = glacier_cover["name"]
            label(f"{n}:water_content", "glacier_water")
            label(f"{n}:ice_content", "glacier_ice")
            label(f"{n}:{glacier_cover['processes'][0]['name']}:output", "glacier_outflow")
            label(f"{n}:{glacier_cover['processes'][1]['name']}:output", "glacier_melt")
            if self.no_snow:
                if st.glacier_no_melt_when_snow_cover:  # IceContainer.cpp:10-32 needs the related snowpack
                    raise UnsupportedStructure("No snowpack provided for the glacier melt limitation.")
            else:
                glsp, gltr, glsub = snowpack_of(glacier_cover)
                if (glsub is None) != (gsub is None):
                    raise UnsupportedStructure("sublimation must be on every snowpack or on none")
                if glsub is not None:
                    set_sublimation(glsp, glsub, slots["glacier_sublimation"], "glacier_snowpack_sublimation")
                    if second:
                        self.twin_labels.add(f"{glsp['name']}:{glsub['name']}:output")
                set_melt(glsp["name"], glsp["processes"][0], slots["glacier_snow_ddf"], slots["glacier_snow_tmelt"],
                         slots["glacier_snow_melt_b"], slots["glacier_snow_melt_c"])
                label(f"{glsp['name']}:water_content", "glacier_snowpack_water")
                label(f"{glsp['name']}:snow_content", "glacier_snowpack_snow")
                label(f"{glsp['name']}:{glsp['processes'][0]['name']}:output", "glacier_snowpack_melt")
                snowpacks.append(glsp["name"])
                set_retention(glsp, "glacier_snow_whc", "glacier_snow_cfr", "glacier_snowpack")
                kind = transfo_kinds[gltr["kind"]] if gltr is not None else _e.SNOW_ICE_NONE
                if second and kind != st.snow_ice_kind:
                    raise UnsupportedStructure("the snow -> ice transformation must be the same on every glacier snowpack")
                if gltr is not None:
                    tp = gltr
                    st.snow_ice_kind = kind
                    tv = _params(tp)
                    if st.snow_ice_kind == _e.SNOW_ICE_CONSTANT:
                        pname = "snow_ice_transformation_rate" if "snow_ice_transformation_rate" in tv else "rate"
                    else:
                        pname = "snow_ice_transformation_basal_acc_coeff" \
                            if "snow_ice_transformation_basal_acc_coeff" in tv else "basal_acc_coeff"
                        st.north_hemisphere = int(not (float(tv.get("north_hemisphere", 1)) == 0))
                    self._set(slots["snow_ice"], glsp["name"], pname, tv[pname])
                    label(f"{glsp['name']}:{tp['name']}:output", "glacier_snowpack_transfo")
        if glacier is not None:
            for name, slot in ((b1, "k_snow"), (b2, "k_ice")):
                bb = basin_bricks.get(name)
                if bb is None or bb["kind"] != "storage" or _proc_kinds(bb) != ["outflow:linear"] or \
                        _target(bb["processes"][0]) != "outlet":
                    raise UnsupportedStructure(f"sub-basin store {name}: expected storage with outflow:linear -> outlet")
                self._set(slot, name, "response_factor", _params(bb["processes"][0])["response_factor"])
        if sorted(snow_targets or []) != sorted(snowpacks):
            raise UnsupportedStructure("snow must go to the snowpack of every land cover")
        # liquid water in the snowpacks: the same processes on every snowpack (model_settings.py:248-255)
        st.snow_retention = 0
        if retention:
            kinds = {(True, r is not None) for _, r in retention.values()}
            if len(retention) != len(snowpacks) or len(kinds) != 1:
                raise UnsupportedStructure("outflow:snow_holding / refreeze:degree_day must be on every snowpack or on none")
            st.snow_retention = _e.RETENTION_HOLDING | (_e.RETENTION_REFREEZE if kinds.pop()[1] else 0) | \
                (_e.RETENTION_RAIN_TO_SNOWPACK if rain_to_snowpack else 0)
            if (st.snow_retention & _e.RETENTION_REFREEZE) and st.melt_kind != _e.MELT_DEGREE_DAY:
                # ProcessRefreezeDegreeDay::IsValid (ProcessRefreezeDegreeDay.cpp:19-37)
                raise UnsupportedStructure("The refreeze:degree_day process requires a melt:degree_day process on the same brick.")
            if glacier2 is not None or slide_on:
                raise UnsupportedStructure("liquid water in the snowpacks runs neither with a second glacier cover nor with "
                                           "transport:snow_slide")
        elif rain_to_snowpack:
            raise UnsupportedStructure("rain routed to the snowpacks needs a snow water retention process")
        if slide_on:
            # the snowpacks of the generic cover always slide, those of the glacier unless skipGlaciers; one set of
            # parameters for every sliding snowpack (ProcessLateralSnowSlide.cpp:24-31)
            if gsp is None or gsp["name"] not in slide_on:
                raise UnsupportedStructure("transport:snow_slide must be on the snowpack of the generic land cover")
            st.snow_slide_kind = _e.SNOW_SLIDE_ALL if len(slide_on) == len(snowpacks) and len(snowpacks) > 1 \
                else _e.SNOW_SLIDE_SKIP_GLACIERS
            names = {"coeff": "slide_coeff", "exp": "slide_exp", "min_slope": "slide_min_slope",
                     "max_slope": "slide_max_slope", "min_snow_holding_depth": "slide_min_holding_depth",
                     "max_snow_depth": "slide_max_snow_depth"}
            first = None
            for sp_name, proc in slide_on.items():
                pv = _params(proc)
                if first is None:
                    first = pv
                elif any(pv[k] != first[k] for k in names):
                    raise UnsupportedStructure("transport:snow_slide: different parameters on the snowpacks")
                for pname, slot in names.items():
                    self._set(slot, sp_name, pname, pv[pname])
        self.basin_store_names = (b1, b2)
        # sub-basin bricks that no glacier feeds stay empty forever (helpers.cpp:55-59): ignore them, but
        # only if they are plain linear stores to the outlet.
        for name, bb in basin_bricks.items():
            if name in (b1, b2):
                continue
            if bb["kind"] != "storage" or _proc_kinds(bb) != ["outflow:linear"] or _target(bb["processes"][0]) != "outlet":
                raise UnsupportedStructure(f"sub-basin brick {name}")

        # --- slow reservoir(s)
        slow = bricks.get(slow_name)
        if slow is None or slow["kind"] != "storage" or slow.get("computed_directly"):
            raise UnsupportedStructure("infiltration target must be a storage solved by the ODE solver")
        kinds = _proc_kinds(slow)
        norm = ["overflow" if k in ("overflow", "outflow:overflow") else k for k in kinds]
        if norm == ["et:socont", "outflow:linear", "overflow"]:
            st.n_soil_stores, st.slow_process_order = 1, 0
        elif norm == ["et:socont", "outflow:linear", "overflow", "percolation:constant"]:
            st.n_soil_stores, st.slow_process_order = 2, 0
        elif norm == ["et:socont", "outflow:linear", "percolation:constant", "overflow"]:
            st.n_soil_stores, st.slow_process_order = 2, 1
        else:
            raise UnsupportedStructure(f"slow reservoir processes {kinds}")
        sparams = _params(slow)
        if "capacity" not in sparams:
            raise UnsupportedStructure("slow reservoir needs a capacity")
        self._set("capacity", slow_name, "capacity", sparams["capacity"])
        label_of = {"et:socont": "slow_et", "outflow:linear": "slow_outflow", "overflow": "slow_overflow",
                    "percolation:constant": "slow_percolation"}
        self.labels[f"{slow_name}:water_content"] = "slow_water"
        slow2_name = None
        for proc, k in zip(slow["processes"], norm):
            self.labels[f"{slow_name}:{proc['name']}:output"] = label_of[k]
            if k == "outflow:linear":
                if _target(proc) != "outlet":
                    raise UnsupportedStructure("slow reservoir outflow must go to the outlet")
                self._set("k_slow_1", slow_name, "response_factor", _params(proc)["response_factor"])
            elif k == "overflow" and _target(proc) != "outlet":
                raise UnsupportedStructure("slow reservoir overflow must go to the outlet")
            elif k == "percolation:constant":
                slow2_name = _target(proc)
                pv = _params(proc)
                self._set("percolation", slow_name, "percolation_rate", pv.get("percolation_rate", pv.get("rate", 0)))
        if st.n_soil_stores == 2:
            s2 = bricks.get(slow2_name)
            if s2 is None or s2["kind"] != "storage" or _proc_kinds(s2) != ["outflow:linear"] or \
                    _target(s2["processes"][0]) != "outlet" or _params(s2).get("capacity") is not None:
                raise UnsupportedStructure("second soil store: expected storage with outflow:linear -> outlet")
            self._set("k_slow_2", slow2_name, "response_factor", _params(s2["processes"][0])["response_factor"])
            self.labels[f"{slow2_name}:water_content"] = "slow2_water"
            self.labels[f"{slow2_name}:{s2['processes'][0]['name']}:output"] = "slow2_outflow"

        # --- quick store
        quick = bricks.get(quick_name)
        if quick is None or quick["kind"] != "storage" or len(quick["processes"]) != 1 or \
                _target(quick["processes"][0]) != "outlet" or _params(quick).get("capacity") is not None:
            raise UnsupportedStructure("surface runoff store: expected one process to the outlet")
        qp = quick["processes"][0]
        if qp["kind"] == "runoff:socont":
            st.quick_kind = _e.QUICK_SOCONT
            self._set("quick", quick_name, "beta", _params(qp)["beta"])
        elif qp["kind"] == "outflow:linear":
            st.quick_kind = _e.QUICK_LINEAR
            self._set("quick", quick_name, "response_factor", _params(qp)["response_factor"])
        else:
            raise UnsupportedStructure(f"surface runoff process {qp['kind']}")
        self.labels[f"{quick_name}:water_content"] = "quick_water"
        self.labels[f"{quick_name}:{qp['name']}:output"] = "quick_runoff"

        known = {self.ground_name, slow_name, quick_name} | set(snowpacks)
        for g in (glacier, glacier2):
            if g is not None:
                known.add(g["name"])
        if slow2_name:
            known.add(slow2_name)
        extra = set(bricks) - known
        if extra:
            raise UnsupportedStructure(f"bricks outside the Socont family: {sorted(extra)}")
        for b in full["hydro_unit_bricks"]:
            if b.get("forcing"):
                raise UnsupportedStructure("brick-level forcing")
        self.names = {"slow": slow_name, "slow2": slow2_name, "quick": quick_name}
        if st.melt_kind < 0:
            st.melt_kind = _e.MELT_DEGREE_DAY
        self.hbk = st

    # ------------------------------------------------------------------------------------------------
    def assign_structure_variants(self, unit_land_covers):
        """ModelBuilder::AssignHydroUnitStructures (ModelBuilder.cpp:36-97) -> glacier_variant flag per unit."""
        return [1 if sid == self.glacier_structure_id else 0 for sid in self.assign_structure_variants_ids(unit_land_covers)]

    def assign_structure_variants_ids(self, unit_land_covers):
        """... -> the id of the structure variant of every unit (results: hydro_units_structure_ids)."""
        out = []
        ids = sorted(self.variant_covers)
        for covers in unit_land_covers:
            if len(ids) <= 1:
                sid = ids[0]
            else:
                present = {n for n, f in covers if abs(f) > PRECISION}
                sid = next((i for i in ids if self.variant_covers[i] == present), -1)
                if sid < 0:
                    sup = [(len(self.variant_covers[i]), i) for i in ids if self.variant_covers[i] >= present]
                    if sup:
                        sid = min(sup)[1]
                    else:
                        size = max(len(c) for c in self.variant_covers.values())
                        sid = min(i for i in ids if len(self.variant_covers[i]) == size)
            out.append(sid)
        return out

    def parameter_vector(self, overrides=None):
        """Parameter slots with optional {(component, name): value} overrides (SetParameterValue)."""
        v = self.params.copy()
        for (component, name), value in (overrides or {}).items():
            for comp in [c.strip() for c in component.split(",")]:
                v[self.resolve(comp, name)] = np.float32(value)
        return v

    def resolve(self, component, name):
        if (component, name) in self.param_index:
            return self.param_index[(component, name)]
        raise KeyError(f"parameter {component}:{name} is not part of this structure")

    def resolve_all(self, component, name):
        """SettingsModel::SetParameterValue target resolution incl. 'type:' components and lists."""
        slots = []
        for comp in [c.strip() for c in component.split(",") if c.strip()]:
            if comp.startswith("type:"):
                kind = comp.split(":", 1)[1]
                for st in self.structures:
                    for b in st["hydro_unit_bricks"] + st["sub_basin_bricks"]:
                        if b["kind"] == kind or (kind == "land_cover" and b["kind"] in ("generic_land_cover", "glacier")):
                            if (b["name"], name) in self.param_index:
                                slots.append(self.param_index[(b["name"], name)])
            elif (comp, name) in self.param_index:
                slots.append(self.param_index[(comp, name)])
        return sorted(set(slots))

    def record_slots(self, labels):
        return [self.labels[label] for label in labels]

    def recorded(self, eng, label, set_index=0):
        """ModelHydro::GetHydroUnitValues(label) [T x basin units] from the engine's recorder (the second glacier
        cover's labels live in the twin units' columns)."""
        values = eng.get_recorded(self.labels[label], set_index)
        layout = getattr(self, "layout", None)
        return values if layout is None else layout.gather(values, label in self.twin_labels)


def slope_m_per_m(value: float, unit: str) -> np.float32:
    """HydroUnitProperty::GetValue(..., "m/m") + the float32 cast of HydroUnit::GetPropertyFloat
    (HydroUnitProperty.cpp:19-49, HydroUnit.cpp:146-148)."""
    if unit in ("m/m", "m_per_m"):
        return np.float32(value)
    if unit in ("degrees", "degree", "deg", "°"):
        return np.float32(math.tan(value * 3.1415926535897932384626433832795 / 180.0))
    raise ValueError(f"unsupported slope unit {unit!r}")


def slope_degrees(value: float, unit: str) -> np.float32:
    """HydroUnit::GetPropertyDouble("slope", "degrees") + the float cast of ProcessLateralSnowSlide.cpp:33-35."""
    if unit in ("degrees", "degree", "deg", "°"):
        return np.float32(value)
    # HydroUnitProperty::GetValue (HydroUnitProperty.cpp:19-49) converts degrees to other units, never back
    raise ValueError(f"The unit 'degrees' is not supported for the property 'slope' given in {unit!r}.")


class UnitLayout:
    """Engine units of a basin: the basin's units in order, then one TWIN per unit that carries the glacier bricks when
    the structure has a second glacier cover (hbk_set_unit_twins). twin_of[engine unit] = basin unit or -1; twin_col[basin
    unit] = engine column of its twin or -1."""

    def __init__(self, n_basin, twins_for):
        self.n_basin = n_basin
        self.twin_of = [-1] * n_basin + list(twins_for)
        self.n_engine = len(self.twin_of)
        self.twin_col = [-1] * n_basin
        for k, u in enumerate(twins_for):
            self.twin_col[u] = n_basin + k

    @property
    def has_twins(self):
        return self.n_engine > self.n_basin

    def expand(self, per_unit, axis=-1):
        """A per-basin-unit array -> per-engine-unit (a twin takes its unit's value: same forcing, area, elevation...)."""
        a = np.asarray(per_unit)
        if not self.has_twins:
            return a
        return np.concatenate([a, np.take(a, self.twin_of[self.n_basin:], axis=axis)], axis=axis)

    def gather(self, engine_values, twin):
        """Engine columns [... x n_engine] -> basin columns [... x n_basin] of a recorded series: the units' own columns,
        or for a label of the second glacier cover the twins' columns (NaN where a unit has no twin)."""
        a = np.asarray(engine_values)
        if not self.has_twins:
            return a
        own = a[..., :self.n_basin]
        if not twin:
            return own
        out = np.full_like(own, np.nan)
        idx = [u for u, c in enumerate(self.twin_col) if c >= 0]
        out[..., idx] = a[..., [self.twin_col[u] for u in idx]]
        return out


def build_engine(structures, units, solver, n_steps, device=0, time_step_days=1.0, start_mjd=44605.0):
    """Compile + create an engine + upload the basin. `units` as in tests/golden (basin order).
    start_mjd: modified Julian date of time step 0 (default 1981-01-01), read by day-of-year processes."""
    cs = CompiledStructure(structures, solver, time_step_days, start_mjd)
    covers = [u["land_covers"] for u in units]
    gv = cs.assign_structure_variants(covers)
    layout = UnitLayout(len(units), [u for u, g in enumerate(gv) if g] if cs.glacier2_name else [])
    cs.layout = layout
    eng = _e.Engine(cs.hbk, layout.n_engine, n_steps, device)
    eng.layout = layout
    area = layout.expand([u["area"] for u in units])
    elev = layout.expand([u.get("elevation", 0.0) for u in units])
    slope = None
    if cs.hbk.quick_kind == _e.QUICK_SOCONT:
        slope = layout.expand([slope_m_per_m(*u["slope"]) for u in units])
    fg = [dict(c).get(cs.ground_name, 1.0) for c in covers]
    fgl = [dict(c).get(cs.glacier_name, 0.0) if cs.glacier_name else 0.0 for c in covers]
    if layout.has_twins:
        # a twin: no ground (a null brick), the second cover's fraction as its glacier fraction
        fg = fg + [0.0] * (layout.n_engine - layout.n_basin)
        fgl = fgl + [dict(covers[u]).get(cs.glacier2_name, 0.0) for u in layout.twin_of[layout.n_basin:]]
        gv = gv + [1] * (layout.n_engine - layout.n_basin)
        eng.set_unit_twins(layout.twin_of)
    eng.set_units(area, elev, slope, gv, fg, fgl)
    if cs.hbk.melt_kind == _e.MELT_DEGREE_DAY_ASPECT:
        eng.set_unit_aspect(layout.expand(
            [_e.ASPECT.get(u.get("aspect_class") or u.get("properties", {}).get("aspect_class"), -1) for u in units]))
    if cs.hbk.snow_slide_kind != _e.SNOW_SLIDE_NONE:
        # hydro_units.py set_connectivity -> SettingsBasin::AddLateralConnection: (receiver position, fraction) per unit
        givers, receivers, fractions = [], [], []
        for i, u in enumerate(units):
            for receiver, fraction in u.get("lateral", []):
                givers.append(i)
                receivers.append(int(receiver))
                fractions.append(float(fraction))
        eng.set_lateral_connections(givers, receivers, fractions, [slope_degrees(*u["slope"]) for u in units])
    return cs, eng
