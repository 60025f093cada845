// oracle/ref_binding.cpp — TEST INFRASTRUCTURE. Python access to the UNMODIFIED reference core
// (compiled from /root/reference/core/src by oracle/Makefile into oracle/_ref/).
//
// The reference's own binding (core/bindings/Binding.cpp) needs pybind11/eigen.h and therefore the
// real Eigen, which this image lacks. This file is written for this repo: it converts numpy arrays
// to the shim array types (oracle/ref_shims/Eigen/Dense) by hand and exposes the module surface the
// reference Python layer imports as `hydrobricks._hydrobricks` (names listed in SURVEY.md §8b,
// Binding.cpp:164-413), so that tools/make_goldens.py can drive the real hb.models.Socont end to end
// and record golden vectors. Two additions the reference module does not have, used only to feed
// the CPU restatement (oracle/hb_oracle.cpp) with the exact same structure:
//   SettingsModel.get_structure()  additionally carries "parameters" and "log_items" per element,
//   SettingsModel.get_solver_name(), get_timer().
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <set>

#include "Action.h"
#include "ActionGlacierEvolutionAreaScaling.h"
#include "ActionGlacierEvolutionDeltaH.h"
#include "ActionGlacierSnowToIceTransformation.h"
#include "ActionLandCoverChange.h"
#include "ContentTypes.h"
#include "Includes.h"
#include "ModelHydro.h"
#include "Parameter.h"
#include "ParameterModifier.h"
#include "SettingsBasin.h"
#include "SettingsModel.h"
#include "SubBasin.h"
#include "Utils.h"

namespace py = pybind11;
using namespace pybind11::literals;

using np_f64 = py::array_t<double, py::array::c_style | py::array::forcecast>;
using np_i32 = py::array_t<int, py::array::c_style | py::array::forcecast>;

namespace {

axd ToAxd(const np_f64& a) {
    axd out(a.size());
    const double* p = a.data();
    for (py::ssize_t i = 0; i < a.size(); ++i) out(i) = p[i];
    return out;
}

axi ToAxi(const np_i32& a) {
    axi out(a.size());
    const int* p = a.data();
    for (py::ssize_t i = 0; i < a.size(); ++i) out(i) = p[i];
    return out;
}

// numpy (rows x cols, C order) -> column-major shim array.
axxd ToAxxd(const np_f64& a) {
    if (a.ndim() != 2) throw py::value_error("expected a 2-D array");
    axxd out(a.shape(0), a.shape(1));
    auto r = a.unchecked<2>();
    for (py::ssize_t i = 0; i < a.shape(0); ++i)
        for (py::ssize_t j = 0; j < a.shape(1); ++j) out(i, j) = r(i, j);
    return out;
}

np_f64 FromAxd(const axd& a) {
    np_f64 out(a.size());
    double* p = out.mutable_data();
    for (long i = 0; i < a.size(); ++i) p[i] = a(i);
    return out;
}

np_f64 FromAxxd(const axxd& a) {
    np_f64 out({static_cast<py::ssize_t>(a.rows()), static_cast<py::ssize_t>(a.cols())});
    auto w = out.mutable_unchecked<2>();
    for (long i = 0; i < a.rows(); ++i)
        for (long j = 0; j < a.cols(); ++j) w(i, j) = a(i, j);
    return out;
}

py::list ParamList(const vector<Parameter>& params) {
    py::list l;
    for (const auto& p : params) {
        py::dict d;
        d["name"] = p.GetName();
        d["value"] = p.GetValue();
        l.append(d);
    }
    return l;
}

py::list ForcingList(const vector<VariableType>& forcing) {
    py::list l;
    for (auto f : forcing) {
        switch (f) {
            case VariableType::Precipitation: l.append("precipitation"); break;
            case VariableType::Temperature: l.append("temperature"); break;
            case VariableType::PET: l.append("pet"); break;
            case VariableType::TemperatureMin: l.append("temperature_min"); break;
            case VariableType::TemperatureMax: l.append("temperature_max"); break;
            case VariableType::Radiation: l.append("solar_radiation"); break;  // names of Binding.cpp:30-52
            default: l.append(static_cast<int>(f)); break;
        }
    }
    return l;
}

py::list OutputList(const vector<OutputSettings>& outputs) {
    py::list l;
    for (const auto& o : outputs) {
        py::dict d;
        d["target"] = o.target;
        d["flux_type"] = ContentTypeToString(o.fluxType);
        d["instantaneous"] = o.isInstantaneous;
        d["static"] = o.isStatic;
        l.append(d);
    }
    return l;
}

py::dict BrickDict(const BrickSettings& b, const char* attach, bool landCover, bool surfaceComponent) {
    py::dict d;
    d["name"] = b.name;
    d["kind"] = b.type;
    d["parent"] = b.parent.empty() ? py::object(py::none()) : py::object(py::cast(b.parent));
    d["attach"] = attach;
    d["computed_directly"] = b.computedDirectly;
    d["is_land_cover"] = landCover;
    d["is_surface_component"] = surfaceComponent;
    d["forcing"] = ForcingList(b.forcing);
    d["parameters"] = ParamList(b.parameters);
    d["log_items"] = b.logItems;
    py::list procs;
    for (const auto& p : b.processes) {
        py::dict pd;
        pd["name"] = p.name;
        pd["kind"] = p.type;
        pd["forcing"] = ForcingList(p.forcing);
        pd["outputs"] = OutputList(p.outputs);
        pd["parameters"] = ParamList(p.parameters);
        pd["log_items"] = p.logItems;
        procs.append(pd);
    }
    d["processes"] = procs;
    return d;
}

py::list SplitterList(SettingsModel& s, bool subBasin) {
    py::list l;
    int n = subBasin ? s.GetSubBasinSplitterCount() : s.GetHydroUnitSplitterCount();
    for (int i = 0; i < n; ++i) {
        const SplitterSettings& sp = subBasin ? s.GetSubBasinSplitterSettings(i) : s.GetHydroUnitSplitterSettings(i);
        py::dict d;
        d["name"] = sp.name;
        d["kind"] = sp.type;
        d["attach"] = subBasin ? "sub_basin" : "hydro_unit";
        d["forcing"] = ForcingList(sp.forcing);
        d["outputs"] = OutputList(sp.outputs);
        d["parameters"] = ParamList(sp.parameters);
        d["log_items"] = sp.logItems;
        l.append(d);
    }
    return l;
}

py::list StructureList(SettingsModel& s) {
    py::list out;
    int n = s.GetStructureCount();
    for (int id = 1; id <= n; ++id) {
        s.SelectStructure(id);
        std::set<int> lc, sc;
        for (int i : s.GetLandCoverBricksIndices()) lc.insert(i);
        for (int i : s.GetSurfaceComponentBricksIndices()) sc.insert(i);
        py::list hb, sb;
        for (int i = 0; i < s.GetHydroUnitBrickCount(); ++i)
            hb.append(BrickDict(s.GetHydroUnitBrickSettings(i), "hydro_unit", lc.count(i) > 0, sc.count(i) > 0));
        for (int i = 0; i < s.GetSubBasinBrickCount(); ++i)
            sb.append(BrickDict(s.GetSubBasinBrickSettings(i), "sub_basin", false, false));
        py::dict d;
        d["id"] = id;
        d["hydro_unit_bricks"] = hb;
        d["sub_basin_bricks"] = sb;
        d["hydro_unit_splitters"] = SplitterList(s, false);
        d["sub_basin_splitters"] = SplitterList(s, true);
        out.append(d);
    }
    if (n >= 1) s.SelectStructure(1);
    return out;
}

struct LogNull {
    LogNull() { LogSetLevel(LogLevel::Error); }
    ~LogNull() { LogSetLevel(LogLevel::Message); }
};

}  // namespace

PYBIND11_MODULE(_hydrobricks, m) {
    m.doc() = "reference hydrobricks core (oracle build, numpy binding)";
    m.attr("__hb_oracle_ref__") = true;

    m.def("init", &InitHydrobricksForPython);
    m.def("init_log", &InitLog, "path"_a);
    m.def("close_log", &CloseLog);
    m.def("set_max_log_level", &SetMaxLogLevel);
    m.def("set_debug_log_level", &SetDebugLogLevel);
    m.def("set_message_log_level", &SetMessageLogLevel);

    py::class_<SettingsModel>(m, "SettingsModel")
        .def(py::init<>())
        .def("log_all", &SettingsModel::SetLogAll, "log_all"_a = true)
        .def("record_fractions", &SettingsModel::SetRecordFractions, "record"_a = true)
        .def("add_logging_to", &SettingsModel::AddLoggingToItem, "name"_a)
        .def("add_brick_logging", [](SettingsModel& s, const string& n) { s.AddBrickLogging(n); }, "name"_a)
        .def("select_process", [](SettingsModel& s, const string& n) { s.SelectProcess(n); }, "name"_a)
        .def("add_process_logging", &SettingsModel::AddProcessLogging, "name"_a)
        .def("add_structure", &SettingsModel::AddStructure)
        .def("set_solver", &SettingsModel::SetSolver, "name"_a)
        .def("set_timer", &SettingsModel::SetTimer, "start_date"_a, "end_date"_a, "time_step"_a, "time_step_unit"_a)
        .def("set_spinup_days", &SettingsModel::SetSpinupDays, "days"_a)
        .def("add_land_cover_brick", &SettingsModel::AddLandCoverBrick, "name"_a, "kind"_a)
        .def("add_hydro_unit_brick", &SettingsModel::AddHydroUnitBrick, "name"_a, "kind"_a = "storage")
        .def("add_sub_basin_brick", &SettingsModel::AddSubBasinBrick, "name"_a, "kind"_a = "storage")
        .def("select_hydro_unit_brick", &SettingsModel::SelectHydroUnitBrickByName, "name"_a)
        .def("add_brick_process", &SettingsModel::AddBrickProcess, "name"_a, "kind"_a, "target"_a = "", "log"_a = false)
        .def("add_process_output", [](SettingsModel& s, const string& t) { s.AddProcessOutput(t); }, "target"_a)
        .def("add_process_parameter", &SettingsModel::AddProcessParameter, "name"_a, "value"_a, "kind"_a = "constant")
        .def("add_process_forcing", &SettingsModel::AddProcessForcing, "name"_a)
        .def("add_brick_parameter", &SettingsModel::AddBrickParameter, "name"_a, "value"_a, "kind"_a = "constant")
        // Not in the reference module; used by tools/make_goldens.py to rebuild the C++ test fixtures
        // (core/tests/src/SolverTest.cpp:117-146, helpers.cpp:9-107).
        .def("add_brick_forcing", &SettingsModel::AddBrickForcing, "name"_a)
        .def("set_process_parameter_value", &SettingsModel::SetProcessParameterValue, "name"_a, "value"_a,
             "kind"_a = "constant")
        .def("set_current_brick_computed_directly", &SettingsModel::SetCurrentBrickComputedDirectly)
        .def("set_parameter_value", &SettingsModel::SetParameterValue, "component"_a, "name"_a, "value"_a)
        .def("generate_precipitation_splitters", &SettingsModel::GeneratePrecipitationSplitters, "with_snow"_a = true,
             "splitter_type"_a = "snow_rain:linear")
        .def("generate_snowpacks", &SettingsModel::GenerateSnowpacks, "snow_melt_process"_a)
        .def("generate_snowpacks_with_water_retention", &SettingsModel::GenerateSnowpacksWithWaterRetention,
             "snow_melt_process"_a, "outflow_process"_a, "rain_to_snowpack"_a = false)
        .def("generate_canopy_interception", &SettingsModel::GenerateCanopyInterception, "cover_name"_a,
             "throughfall_target"_a)
        .def("add_snowpack_refreezing", &SettingsModel::AddSnowpackRefreezing,
             "refreezing_process"_a = "refreeze:degree_day")
        .def("add_snowpack_sublimation", &SettingsModel::AddSnowpackSublimation, "sublimation_process"_a)
        .def("add_snow_ice_transformation", &SettingsModel::AddSnowIceTransformation,
             "transformation_process"_a = "transform:snow_ice_swat")
        .def("add_snow_redistribution", &SettingsModel::AddSnowRedistribution,
             "redistribution_process"_a = "transport:snow_slide", "skip_glaciers"_a = false)
        .def("set_process_outputs_as_instantaneous", &SettingsModel::SetProcessOutputsAsInstantaneous)
        .def("set_process_outputs_as_static", &SettingsModel::SetProcessOutputsAsStatic)
        .def("get_structure", [](SettingsModel& s) { return StructureList(s); })
        .def("get_solver_name", [](SettingsModel& s) { return s.GetSolverSettings().name; })
        .def("get_timer", [](SettingsModel& s) {
            const TimerSettings& t = s.GetTimerSettings();
            py::dict d;
            d["start"] = t.start;
            d["end"] = t.end;
            d["time_step"] = t.timeStep;
            d["time_step_unit"] = t.timeStepUnit;
            d["spinup_days"] = t.spinupDays;
            return d;
        });

    py::class_<SettingsBasin>(m, "SettingsBasin")
        .def(py::init<>())
        .def("add_hydro_unit", &SettingsBasin::AddHydroUnit, "id"_a, "area"_a, "elevation"_a = -9999)
        .def("add_land_cover", &SettingsBasin::AddLandCover, "name"_a, "kind"_a, "fraction"_a)
        .def("add_hydro_unit_property_str", &SettingsBasin::AddHydroUnitPropertyString, "name"_a, "value"_a)
        .def("add_hydro_unit_property_double", &SettingsBasin::AddHydroUnitPropertyDouble, "name"_a, "value"_a,
             "unit"_a)
        .def("add_lateral_connection", &SettingsBasin::AddLateralConnection, "giver_hydro_unit_id"_a,
             "receiver_hydro_unit_id"_a, "fraction"_a, "type"_a = "")
        .def("get_lateral_connection_count", &SettingsBasin::GetLateralConnectionCount)
        .def("clear", &SettingsBasin::Clear);

    py::class_<SubBasin>(m, "SubBasin").def(py::init<>()).def(
        "init",
        [](SubBasin& s, SettingsBasin& bs) {
            auto r = s.Initialize(bs);
            if (!r) throw py::value_error(r.error());
        },
        "spatial_structure"_a);

    py::class_<Parameter>(m, "Parameter")
        .def(py::init<const string&, float>())
        .def_property("name", &Parameter::GetName, &Parameter::SetName)
        .def_property("value", &Parameter::GetValue, &Parameter::SetValue)
        .def("get_name", &Parameter::GetName)
        .def("set_name", &Parameter::SetName)
        .def("get_value", &Parameter::GetValue)
        .def("set_value", &Parameter::SetValue)
        .def("set_modifier", &Parameter::SetModifier, "modifier"_a)
        .def("has_modifier", &Parameter::HasModifier)
        .def("update_from_modifier", &Parameter::UpdateFromModifier, "date"_a)
        .def("__repr__", [](const Parameter& a) { return "<_hydrobricks.Parameter named '" + a.GetName() + "'>"; });

    py::enum_<ParameterModifierType>(m, "ParameterModifierType")
        .value("Yearly", ParameterModifierType::Yearly)
        .value("Monthly", ParameterModifierType::Monthly)
        .value("Dates", ParameterModifierType::Dates);

    py::class_<ParameterModifier>(m, "ParameterModifier")
        .def(py::init<>())
        .def(py::init<ParameterModifierType>(), "type"_a)
        .def("set_type", &ParameterModifier::SetType, "type"_a)
        .def("get_type", &ParameterModifier::GetType)
        .def("set_yearly_values", py::overload_cast<int, int, const vecFloat&>(&ParameterModifier::SetYearlyValues),
             "year_start"_a, "year_end"_a, "values"_a)
        .def("set_monthly_values", py::overload_cast<const vecFloat&>(&ParameterModifier::SetMonthlyValues), "values"_a)
        .def("set_dates_and_values", &ParameterModifier::SetDatesAndValues, "dates"_a, "values"_a)
        .def("update_value", &ParameterModifier::UpdateValue, "date"_a)
        .def("updates_on_year_change", &ParameterModifier::UpdatesOnYearChange)
        .def("updates_on_month_change", &ParameterModifier::UpdatesOnMonthChange)
        .def("updates_on_date_change", &ParameterModifier::UpdatesOnDateChange);

    py::class_<TimeSeries>(m, "TimeSeries")
        .def_static(
            "create",
            [](const string& name, const np_f64& time, const np_i32& ids, const np_f64& data) {
                return TimeSeries::Create(name, ToAxd(time), ToAxi(ids), ToAxxd(data));
            },
            "data_name"_a, "time"_a, "ids"_a, "data"_a);

    py::class_<ModelHydro>(m, "ModelHydro")
        .def(py::init<>())
        .def(
            "init_with_basin",
            [](ModelHydro& mh, SettingsModel& ms, SettingsBasin& bs) {
                auto r = mh.InitializeWithBasin(ms, bs);
                if (!r) throw py::value_error(r.error());
            },
            "model_settings"_a, "basin_settings"_a)
        .def("add_action", &ModelHydro::AddAction, "action"_a)
        .def("get_action_count", &ModelHydro::GetActionCount)
        .def("get_sporadic_action_item_count", &ModelHydro::GetSporadicActionItemCount)
        .def(
            "create_time_series",
            [](ModelHydro& mh, const string& name, const np_f64& time, const np_i32& ids, const np_f64& data) {
                return mh.CreateTimeSeries(name, ToAxd(time), ToAxi(ids), ToAxxd(data));
            },
            "data_name"_a, "time"_a, "ids"_a, "data"_a)
        .def("clear_time_series", &ModelHydro::ClearTimeSeries)
        .def("attach_time_series_to_hydro_units", &ModelHydro::AttachTimeSeriesToHydroUnits)
        .def("update_parameters", &ModelHydro::UpdateParameters, "model_settings"_a)
        .def("forcing_loaded", &ModelHydro::ForcingLoaded)
        .def("is_valid", &ModelHydro::IsValid)
        .def(
            "run",
            [](ModelHydro& mh) {
                auto r = mh.Run();
                if (!r) throw py::value_error(r.error());
            },
            py::call_guard<py::gil_scoped_release>())
        .def("reset", &ModelHydro::Reset)
        .def("save_as_initial_state", &ModelHydro::SaveAsInitialState)
        .def("get_outlet_discharge", [](ModelHydro& mh) { return FromAxd(mh.GetOutletDischarge()); })
        .def("get_total_outlet_discharge", &ModelHydro::GetTotalOutletDischarge)
        .def("get_total_et", &ModelHydro::GetTotalET)
        .def("get_total_water_storage_changes", &ModelHydro::GetTotalWaterStorageChanges)
        .def("get_total_snow_storage_changes", &ModelHydro::GetTotalSnowStorageChanges)
        .def("get_total_glacier_storage_changes", &ModelHydro::GetTotalGlacierStorageChanges)
        .def("dump_outputs", &ModelHydro::DumpOutputs, "path"_a)
        .def("get_hydro_unit_values",
             [](ModelHydro& mh, const string& label) { return FromAxxd(mh.GetHydroUnitValues(label)); }, "label"_a)
        .def("get_hydro_unit_fractions",
             [](ModelHydro& mh, const string& label) { return FromAxxd(mh.GetHydroUnitFractions(label)); }, "label"_a)
        .def("get_recorded_hydro_unit_labels", &ModelHydro::GetRecordedHydroUnitLabels)
        .def("get_recorded_hydro_unit_fraction_labels", &ModelHydro::GetRecordedHydroUnitFractionLabels)
        .def("get_hydro_unit_ids", &ModelHydro::GetHydroUnitIds)
        .def("get_hydro_unit_areas", [](ModelHydro& mh) { return FromAxd(mh.GetHydroUnitAreas()); });

    py::class_<Action>(m, "Action").def(py::init<>());

    py::class_<ActionLandCoverChange, Action>(m, "ActionLandCoverChange")
        .def(py::init<>())
        .def("add_change", &ActionLandCoverChange::AddChange, "date"_a, "hydro_unit_id"_a, "land_cover"_a, "area"_a)
        .def("get_change_count", &ActionLandCoverChange::GetChangeCount)
        .def("get_land_cover_count", &ActionLandCoverChange::GetLandCoverCount);

    py::class_<ActionGlacierEvolutionDeltaH, Action>(m, "ActionGlacierEvolutionDeltaH")
        .def(py::init<>())
        .def(
            "add_lookup_tables",
            [](ActionGlacierEvolutionDeltaH& a, int month, const string& cover, const np_i32& ids, const np_f64& areas,
               const np_f64& volumes) { a.AddLookupTables(month, cover, ToAxi(ids), ToAxxd(areas), ToAxxd(volumes)); },
            "month_num"_a, "land_cover"_a, "hu_ids"_a, "areas"_a, "volumes"_a)
        .def("get_land_cover_name", &ActionGlacierEvolutionDeltaH::GetLandCoverName)
        .def("get_lookup_table_area",
             [](ActionGlacierEvolutionDeltaH& a) { return FromAxxd(a.GetLookupTableArea()); })
        .def("get_lookup_table_volume",
             [](ActionGlacierEvolutionDeltaH& a) { return FromAxxd(a.GetLookupTableVolume()); });

    py::class_<ActionGlacierEvolutionAreaScaling, Action>(m, "ActionGlacierEvolutionAreaScaling")
        .def(py::init<>())
        .def(
            "add_lookup_tables",
            [](ActionGlacierEvolutionAreaScaling& a, int month, const string& cover, const np_i32& ids,
               const np_f64& areas,
               const np_f64& volumes) { a.AddLookupTables(month, cover, ToAxi(ids), ToAxxd(areas), ToAxxd(volumes)); },
            "month_num"_a, "land_cover"_a, "hu_ids"_a, "areas"_a, "volumes"_a)
        .def("get_land_cover_name", &ActionGlacierEvolutionAreaScaling::GetLandCoverName);

    py::class_<ActionGlacierSnowToIceTransformation, Action>(m, "ActionGlacierSnowToIceTransformation")
        .def(py::init<int, int, const std::string&>(), "month"_a, "day"_a, "land_cover"_a)
        .def("get_land_cover_name", &ActionGlacierSnowToIceTransformation::GetLandCoverName);

    py::class_<LogNull>(m, "LogNull").def(py::init<>());
}
