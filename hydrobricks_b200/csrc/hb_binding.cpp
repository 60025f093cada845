// hb_binding.cpp — pybind11 module `_hydrobricks` of the B200 engine: the surface of the reference's module
// (core/bindings/Binding.cpp:164-413) for the ModelHydro::Run() path, over hb_host.hpp / the C-ABI (libhbk.so).
// numpy via pybind11/numpy.h (no Eigen). Batched extras: set_parameter_sets, run_batch,
// get_outlet_discharge_batch, set_station_forcing.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "hb_host.hpp"

namespace py = pybind11;
using namespace pybind11::literals;
using f64arr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using f32arr = py::array_t<float, py::array::c_style | py::array::forcecast>;
using i32arr = py::array_t<int, py::array::c_style | py::array::forcecast>;

namespace {

py::list paramList(const std::vector<hb::Param>& v) {
    py::list l;
    for (auto& p : v) l.append(py::dict("name"_a = p.name, "value"_a = (double)p.value));
    return l;
}
py::list outputList(const std::vector<hb::Output>& v) {
    py::list l;
    for (auto& o : v)
        l.append(py::dict("target"_a = o.target, "flux_type"_a = o.fluxType, "instantaneous"_a = o.instantaneous,
                          "static"_a = o.isStatic));
    return l;
}
py::dict brickDict(const hb::Brick& b) {
    py::dict d;
    d["name"] = b.name;
    d["kind"] = b.kind;
    d["parent"] = b.parent.empty() ? py::object(py::none()) : py::object(py::cast(b.parent));
    d["attach"] = b.attach;
    d["computed_directly"] = b.computedDirectly;
    d["is_land_cover"] = b.isLandCover;
    d["is_surface_component"] = b.isSurfaceComponent;
    d["forcing"] = b.forcing;
    d["parameters"] = paramList(b.parameters);
    d["log_items"] = b.logItems;
    py::list procs;
    for (auto& p : b.processes)
        procs.append(py::dict("name"_a = p.name, "kind"_a = p.kind, "forcing"_a = p.forcing,
                              "outputs"_a = outputList(p.outputs), "parameters"_a = paramList(p.parameters),
                              "log_items"_a = p.logItems));
    d["processes"] = procs;
    return d;
}
py::list structureList(const hb::SettingsModel& s) {  // Binding.cpp:119-160 + parameters / log items
    py::list out;
    for (auto& st : s.Structures()) {
        py::list hb_, sb, hs, ss;
        for (auto& b : st.hydroUnitBricks) hb_.append(brickDict(b));
        for (auto& b : st.subBasinBricks) sb.append(brickDict(b));
        auto spl = [](const hb::Splitter& sp) {
            return py::dict("name"_a = sp.name, "kind"_a = sp.kind, "attach"_a = sp.attach, "forcing"_a = sp.forcing,
                            "outputs"_a = outputList(sp.outputs), "parameters"_a = paramList(sp.parameters),
                            "log_items"_a = sp.logItems);
        };
        for (auto& sp : st.hydroUnitSplitters) hs.append(spl(sp));
        for (auto& sp : st.subBasinSplitters) ss.append(spl(sp));
        out.append(py::dict("id"_a = st.id, "log_items"_a = st.logItems, "hydro_unit_bricks"_a = hb_,
                            "sub_basin_bricks"_a = sb, "hydro_unit_splitters"_a = hs, "sub_basin_splitters"_a = ss));
    }
    return out;
}

int varIndex(std::string name) {
    std::transform(name.begin(), name.end(), name.begin(), ::tolower);
    if (name == "precipitation" || name == "p") return HBK_VAR_PRECIPITATION;
    if (name == "temperature" || name == "t") return HBK_VAR_TEMPERATURE;
    if (name == "pet") return HBK_VAR_PET;
    if (name == "solar_radiation" || name == "r_solar" || name == "r") return HBK_VAR_SOLAR_RADIATION;
    throw py::value_error("Unknown forcing variable '" + name + "'");
}

struct LogNull {};

}  // namespace

PYBIND11_MODULE(_hydrobricks, m) {
    m.doc() = "hydrobricks B200 engine: drop-in for the ModelHydro::Run() path of hydrobricks._hydrobricks";
    m.attr("__hb_b200__") = true;
    py::register_exception<hb::Unsupported>(m, "UnsupportedStructure", PyExc_ValueError);

    m.def("init", []() { return true; });
    m.def("init_log", [](const std::string&) {}, "path"_a);
    m.def("close_log", []() {});
    m.def("set_max_log_level", []() {});
    m.def("set_debug_log_level", []() {});
    m.def("set_message_log_level", []() {});

    py::class_<hb::SettingsModel>(m, "SettingsModel")
        .def(py::init<>())
        .def("log_all", &hb::SettingsModel::SetLogAll, "log_all"_a = true)
        .def("record_fractions", &hb::SettingsModel::SetRecordFractions, "record"_a = true)
        .def("add_logging_to", &hb::SettingsModel::AddLoggingToItem, "name"_a)
        .def("add_brick_logging", &hb::SettingsModel::AddBrickLogging, "name"_a)
        .def("select_process", &hb::SettingsModel::SelectProcess, "name"_a)
        .def("add_process_logging", &hb::SettingsModel::AddProcessLogging, "name"_a)
        .def("add_structure", &hb::SettingsModel::AddStructure)
        .def("set_solver", &hb::SettingsModel::SetSolver, "name"_a)
        .def("set_timer", &hb::SettingsModel::SetTimer, "start_date"_a, "end_date"_a, "time_step"_a, "time_step_unit"_a)
        .def("set_spinup_days", &hb::SettingsModel::SetSpinupDays, "days"_a)
        .def("add_land_cover_brick", &hb::SettingsModel::AddLandCoverBrick, "name"_a, "kind"_a)
        .def("add_hydro_unit_brick", &hb::SettingsModel::AddHydroUnitBrick, "name"_a, "kind"_a = "storage")
        .def("add_sub_basin_brick", &hb::SettingsModel::AddSubBasinBrick, "name"_a, "kind"_a = "storage")
        .def("select_hydro_unit_brick", &hb::SettingsModel::SelectHydroUnitBrickByName, "name"_a)
        .def("add_brick_process", &hb::SettingsModel::AddBrickProcess, "name"_a, "kind"_a, "target"_a = "", "log"_a = false)
        .def("add_process_output", [](hb::SettingsModel& s, const std::string& t) { s.AddProcessOutput(t); }, "target"_a)
        .def("add_process_parameter", &hb::SettingsModel::AddProcessParameter, "name"_a, "value"_a, "kind"_a = "constant")
        .def("set_process_parameter_value", &hb::SettingsModel::SetProcessParameterValue, "name"_a, "value"_a,
             "kind"_a = "constant")
        .def("add_process_forcing", &hb::SettingsModel::AddProcessForcing, "name"_a)
        .def("add_brick_parameter", &hb::SettingsModel::AddBrickParameter, "name"_a, "value"_a, "kind"_a = "constant")
        .def("add_brick_forcing", &hb::SettingsModel::AddBrickForcing, "name"_a)
        .def("set_current_brick_computed_directly", &hb::SettingsModel::SetCurrentBrickComputedDirectly)
        .def("set_parameter_value", &hb::SettingsModel::SetParameterValue, "component"_a, "name"_a, "value"_a)
        .def("generate_precipitation_splitters", &hb::SettingsModel::GeneratePrecipitationSplitters, "with_snow"_a = true,
             "splitter_type"_a = "snow_rain:linear")
        .def("generate_snowpacks", &hb::SettingsModel::GenerateSnowpacks, "snow_melt_process"_a)
        .def("generate_snowpacks_with_water_retention", &hb::SettingsModel::GenerateSnowpacksWithWaterRetention,
             "snow_melt_process"_a, "outflow_process"_a, "rain_to_snowpack"_a = false)
        .def("generate_canopy_interception",
             [](hb::SettingsModel& s, py::args, py::kwargs) { s.Unsupported("canopy interception"); })
        .def("add_snowpack_refreezing", &hb::SettingsModel::AddSnowpackRefreezing,
             "refreezing_process"_a = "refreeze:degree_day")
        .def("add_snowpack_sublimation", &hb::SettingsModel::AddSnowpackSublimation, "sublimation_process"_a)
        .def("add_snow_ice_transformation", &hb::SettingsModel::AddSnowIceTransformation,
             "transformation_process"_a = "transform:snow_ice_swat")
        .def("add_snow_redistribution", &hb::SettingsModel::AddSnowRedistribution,
             "redistribution_process"_a = "transport:snow_slide", "skip_glaciers"_a = false)
        .def("set_process_outputs_as_instantaneous", &hb::SettingsModel::SetProcessOutputsAsInstantaneous)
        .def("set_process_outputs_as_static", &hb::SettingsModel::SetProcessOutputsAsStatic)
        .def("get_structure", [](hb::SettingsModel& s) { return structureList(s); })
        .def("get_solver_name", [](hb::SettingsModel& s) { return s.solver; })
        .def("get_hydro_unit_log_labels", &hb::SettingsModel::GetHydroUnitLogLabels)
        .def("get_timer", [](hb::SettingsModel& s) {
            return py::dict("start"_a = s.timerStart, "end"_a = s.timerEnd, "time_step"_a = s.timeStep,
                            "time_step_unit"_a = s.timeStepUnit, "spinup_days"_a = s.spinupDays);
        });

    py::class_<hb::SettingsBasin>(m, "SettingsBasin")
        .def(py::init<>())
        .def("add_hydro_unit", &hb::SettingsBasin::AddHydroUnit, "id"_a, "area"_a, "elevation"_a = -9999)
        .def("add_land_cover", &hb::SettingsBasin::AddLandCover, "name"_a, "kind"_a, "fraction"_a)
        .def("add_hydro_unit_property_str", &hb::SettingsBasin::AddHydroUnitPropertyString, "name"_a, "value"_a)
        .def("add_hydro_unit_property_double", &hb::SettingsBasin::AddHydroUnitPropertyDouble, "name"_a, "value"_a, "unit"_a)
        .def("add_lateral_connection", &hb::SettingsBasin::AddLateralConnection, "giver_hydro_unit_id"_a,
             "receiver_hydro_unit_id"_a, "fraction"_a, "type"_a = "")
        .def("get_lateral_connection_count", &hb::SettingsBasin::GetLateralConnectionCount)
        .def("clear", &hb::SettingsBasin::Clear);

    // SubBasin (Binding.cpp:260-268): the reference builds the spatial structure here; this engine builds it inside
    // ModelHydro.init_with_basin, so the class only validates the basin settings the way SubBasin::Initialize does.
    struct SubBasinShell {};
    py::class_<SubBasinShell>(m, "SubBasin")
        .def(py::init<>())
        .def("init",
             [](SubBasinShell&, hb::SettingsBasin& bs) {
                 if (bs.units.empty()) throw py::value_error("Basin initialization failed: no hydro unit");
             },
             "spatial_structure"_a);

    py::class_<hb::Parameter>(m, "Parameter")
        .def(py::init<const std::string&, float>())
        .def_readwrite("name", &hb::Parameter::name)
        .def_readwrite("value", &hb::Parameter::value)
        .def("get_name", [](hb::Parameter& p) { return p.name; })
        .def("set_name", [](hb::Parameter& p, const std::string& n) { p.name = n; })
        .def("get_value", [](hb::Parameter& p) { return p.value; })
        .def("set_value", [](hb::Parameter& p, float v) { p.value = v; })
        .def("set_modifier", &hb::Parameter::SetModifier, "modifier"_a)
        .def("has_modifier", [](hb::Parameter& p) { return p.hasModifier; })
        .def("update_from_modifier", &hb::Parameter::UpdateFromModifier, "date"_a)
        .def("__repr__", [](const hb::Parameter& p) { return "<_hydrobricks.Parameter named '" + p.name + "'>"; });

    py::enum_<hb::ParameterModifierType>(m, "ParameterModifierType")
        .value("Yearly", hb::ParameterModifierType::Yearly)
        .value("Monthly", hb::ParameterModifierType::Monthly)
        .value("Dates", hb::ParameterModifierType::Dates);

    py::class_<hb::ParameterModifier>(m, "ParameterModifier")
        .def(py::init<>())
        .def(py::init<hb::ParameterModifierType>(), "type"_a)
        .def("set_type", [](hb::ParameterModifier& p, hb::ParameterModifierType t) { p.type = t; }, "type"_a)
        .def("get_type", [](hb::ParameterModifier& p) { return p.type; })
        .def("set_yearly_values", &hb::ParameterModifier::SetYearlyValues, "year_start"_a, "year_end"_a, "values"_a)
        .def("set_monthly_values", &hb::ParameterModifier::SetMonthlyValues, "values"_a)
        .def("set_dates_and_values", &hb::ParameterModifier::SetDatesAndValues, "dates"_a, "values"_a)
        .def("update_value", &hb::ParameterModifier::UpdateValue, "date"_a)
        .def("updates_on_year_change", &hb::ParameterModifier::UpdatesOnYearChange)
        .def("updates_on_month_change", &hb::ParameterModifier::UpdatesOnMonthChange)
        .def("updates_on_date_change", &hb::ParameterModifier::UpdatesOnDateChange);

    py::class_<hb::TimeSeries>(m, "TimeSeries")
        .def_static("create",
                    [](const std::string& name, const f64arr& time, const i32arr& ids, const f64arr& data) {
                        if (data.ndim() != 2 || data.shape(0) != time.size() || data.shape(1) != ids.size())
                            throw py::value_error("TimeSeries.create: data must be [len(time) x len(ids)]");
                        hb::TimeSeries ts;
                        ts.name = name;
                        ts.time.assign(time.data(), time.data() + time.size());
                        ts.ids.assign(ids.data(), ids.data() + ids.size());
                        ts.data.assign(data.data(), data.data() + data.size());
                        return ts;
                    },
                    "data_name"_a, "time"_a, "ids"_a, "data"_a);

    py::class_<hb::Action>(m, "Action").def(py::init<>());
    py::class_<hb::ActionLandCoverChange, hb::Action>(m, "ActionLandCoverChange")
        .def(py::init<>())
        .def("add_change", &hb::ActionLandCoverChange::AddChange, "date"_a, "hydro_unit_id"_a, "land_cover"_a, "area"_a)
        .def("get_change_count", &hb::ActionLandCoverChange::GetChangeCount)
        .def("get_land_cover_count", &hb::ActionLandCoverChange::GetLandCoverCount);
    auto addTables = [](hb::ActionGlacierEvolutionDeltaH& a, int month, const std::string& cover, const i32arr& ids,
                        const f64arr& areas, const f64arr& volumes) {
        if (areas.ndim() != 2 || volumes.ndim() != 2 || areas.shape(1) != ids.size() || volumes.shape(0) != areas.shape(0) ||
            volumes.shape(1) != areas.shape(1))
            throw py::value_error("lookup tables must be [rows x units]");
        a.AddLookupTables(month, cover, std::vector<int>(ids.data(), ids.data() + ids.size()), (int)areas.shape(0),
                          std::vector<double>(areas.data(), areas.data() + areas.size()),
                          std::vector<double>(volumes.data(), volumes.data() + volumes.size()));
    };
    py::class_<hb::ActionGlacierEvolutionDeltaH, hb::Action>(m, "ActionGlacierEvolutionDeltaH")
        .def(py::init<>())
        .def("add_lookup_tables", addTables, "month_num"_a, "land_cover"_a, "hu_ids"_a, "areas"_a, "volumes"_a)
        .def("get_land_cover_name", [](hb::ActionGlacierEvolutionDeltaH& a) { return a.cover; })
        .def("get_hydro_unit_ids", [](hb::ActionGlacierEvolutionDeltaH& a) { return i32arr(a.unitIds.size(), a.unitIds.data()); })
        .def("get_lookup_table_area", [](hb::ActionGlacierEvolutionDeltaH& a) {
            f64arr out({(py::ssize_t)a.rows, (py::ssize_t)a.unitIds.size()});
            std::copy(a.areas.begin(), a.areas.end(), out.mutable_data());
            return out;
        })
        .def("get_lookup_table_volume", [](hb::ActionGlacierEvolutionDeltaH& a) {
            f64arr out({(py::ssize_t)a.rows, (py::ssize_t)a.unitIds.size()});
            std::copy(a.volumes.begin(), a.volumes.end(), out.mutable_data());
            return out;
        });
    py::class_<hb::ActionGlacierEvolutionAreaScaling, hb::ActionGlacierEvolutionDeltaH>(m, "ActionGlacierEvolutionAreaScaling")
        .def(py::init<>());
    py::class_<hb::ActionGlacierSnowToIceTransformation, hb::Action>(m, "ActionGlacierSnowToIceTransformation")
        .def(py::init<int, int, const std::string&>(), "month"_a, "day"_a, "land_cover"_a)
        .def("get_land_cover_name", [](hb::ActionGlacierSnowToIceTransformation& a) { return a.cover; });

    py::class_<hb::ModelHydro>(m, "ModelHydro")
        .def(py::init<>())
        .def("init_with_basin", &hb::ModelHydro::InitializeWithBasin, "model_settings"_a, "basin_settings"_a, "device"_a = 0,
             py::keep_alive<1, 2>())
        .def("is_valid", &hb::ModelHydro::IsValid)
        .def("add_action", &hb::ModelHydro::AddAction, "action"_a, py::keep_alive<1, 2>())
        .def("get_action_count", &hb::ModelHydro::GetActionCount)
        .def("get_sporadic_action_item_count", &hb::ModelHydro::GetSporadicActionItemCount)
        .def("create_time_series",
             [](hb::ModelHydro& mh, const std::string& name, const f64arr& time, const i32arr& ids, const f64arr& data) {
                 if (data.ndim() != 2 || data.shape(0) != time.size() || data.shape(1) != ids.size()) return false;
                 return mh.CreateTimeSeries(name, time.data(), (int)time.size(), ids.data(), (int)ids.size(), data.data());
             },
             "data_name"_a, "time"_a, "ids"_a, "data"_a)
        .def("add_time_series", &hb::ModelHydro::AddTimeSeries, "time_series"_a)
        .def("clear_time_series", &hb::ModelHydro::ClearTimeSeries)
        .def("attach_time_series_to_hydro_units", &hb::ModelHydro::AttachTimeSeriesToHydroUnits)
        .def("forcing_loaded", &hb::ModelHydro::ForcingLoaded)
        .def("set_station_forcing",
             [](hb::ModelHydro& mh, const std::string& var, const f64arr& series, double zref, py::object gradient,
                py::object correction) {
                 f64arr g, c;
                 double gs = 0, cs = 1;
                 const double *gp = nullptr, *cp = nullptr;
                 if (py::isinstance<py::float_>(gradient) || py::isinstance<py::int_>(gradient)) {
                     gs = gradient.cast<double>();
                 } else {
                     g = gradient.cast<f64arr>();
                     gp = g.data();
                 }
                 if (py::isinstance<py::float_>(correction) || py::isinstance<py::int_>(correction)) {
                     cs = correction.cast<double>();
                 } else {
                     c = correction.cast<f64arr>();
                     cp = c.data();
                 }
                 mh.SetStationForcing(varIndex(var), series.data(), (int)series.size(), zref, gs, gp, cs, cp);
             },
             "variable"_a, "series"_a, "ref_elevation"_a = 0.0, "gradient"_a = 0.0, "correction"_a = 1.0)
        .def("update_parameters", &hb::ModelHydro::UpdateParameters, "model_settings"_a)
        .def("set_parameter_sets",
             [](hb::ModelHydro& mh, const f32arr& v) {
                 if (v.ndim() != 2 || v.shape(1) != HBK_N_PARAMS) throw py::value_error("expected [n_sets x HBK_N_PARAMS]");
                 mh.SetParameterSets(v.data(), (int)v.shape(0));
             },
             "values"_a)
        .def("parameter_slot", &hb::ModelHydro::ParameterSlot, "component"_a, "name"_a)
        .def("base_parameters", [](hb::ModelHydro& mh) {
            auto v = mh.BaseParameters();
            return f32arr(v.size(), v.data());
        })
        .def("reset", &hb::ModelHydro::Reset)
        .def("save_as_initial_state", &hb::ModelHydro::SaveAsInitialState)
        .def("run", [](hb::ModelHydro& mh) { mh.Run(true); }, py::call_guard<py::gil_scoped_release>())
        .def("run_batch",
             [](hb::ModelHydro& mh, const f32arr& v, bool record) {
                 if (v.ndim() != 2 || v.shape(1) != HBK_N_PARAMS) throw py::value_error("expected [n_sets x HBK_N_PARAMS]");
                 mh.SetParameterSets(v.data(), (int)v.shape(0));
                 {
                     py::gil_scoped_release nogil;
                     mh.Run(record);
                 }
                 f64arr out({(py::ssize_t)mh.SetCount(), (py::ssize_t)mh.StepCount()});
                 mh.GetOutlet(0, mh.SetCount(), out.mutable_data());
                 return out;
             },
             "parameter_sets"_a, "record"_a = false)
        .def("get_outlet_discharge", [](hb::ModelHydro& mh) {
            f64arr out(mh.StepCount());
            mh.GetOutlet(0, 1, out.mutable_data());
            return out;
        })
        .def("get_outlet_discharge_batch", [](hb::ModelHydro& mh) {
            f64arr out({(py::ssize_t)mh.SetCount(), (py::ssize_t)mh.StepCount()});
            mh.GetOutlet(0, mh.SetCount(), out.mutable_data());
            return out;
        })
        .def("get_total_outlet_discharge", [](hb::ModelHydro& mh) {
            std::vector<double> q(mh.StepCount());
            mh.GetOutlet(0, 1, q.data());
            double s = 0;
            for (double v : q) s += v;
            return s;
        })
        .def("get_recorded_hydro_unit_labels", &hb::ModelHydro::GetRecordedHydroUnitLabels)
        .def("get_recorded_hydro_unit_fraction_labels", &hb::ModelHydro::GetRecordedHydroUnitFractionLabels)
        .def("get_hydro_unit_values", [](hb::ModelHydro& mh, const std::string& label) {
            f64arr out({(py::ssize_t)mh.StepCount(), (py::ssize_t)mh.UnitCount()});
            mh.GetHydroUnitValues(label, 0, out.mutable_data());
            return out;
        }, "label"_a)
        .def("get_hydro_unit_fractions", [](hb::ModelHydro& mh, const std::string& label) {
            f64arr out({(py::ssize_t)mh.StepCount(), (py::ssize_t)mh.UnitCount()});
            mh.GetHydroUnitFractions(label, out.mutable_data());
            return out;
        }, "label"_a)
        .def("get_hydro_unit_ids", &hb::ModelHydro::GetHydroUnitIds)
        .def("get_hydro_unit_areas", [](hb::ModelHydro& mh) {
            auto a = mh.GetHydroUnitAreas();
            return f64arr(a.size(), a.data());
        })
        .def("engine_unit_count", &hb::ModelHydro::EngineUnitCount)
        .def("twin_of", [](hb::ModelHydro& mh) {
            auto v = mh.TwinOf();
            return i32arr(v.size(), v.data());
        }, "Engine unit -> basin unit whose second glacier cover it carries, -1 for the units themselves (empty: no twins).")
        .def("get_state", [](hb::ModelHydro& mh, int slot) {
            f64arr out({(py::ssize_t)mh.SetCount(), (py::ssize_t)mh.EngineUnitCount()});
            mh.GetState(slot, out.mutable_data());
            return out;
        }, "state_slot"_a)
        // water-balance totals (Logger.cpp:168-339): area-weighted ET sum, end-of-run storage contents
        .def("get_total_et", [](hb::ModelHydro& mh) {
            const int T = mh.StepCount(), U = mh.EngineUnitCount();
            std::vector<double> et((size_t)T * U);
            mh.GetRecorded(HBK_R_SLOW_ET, 0, et.data());
            const auto w = mh.EngineUnitWeights();
            double sum = 0;
            for (int t = 0; t < T; ++t)
                for (int u = 0; u < U; ++u) {
                    double v = et[(size_t)t * U + u];
                    if (v == v) sum += v * w[u];
                }
            return sum;
        })
        .def("get_total_water_storage_changes", [](hb::ModelHydro& mh) {
            const int U = mh.EngineUnitCount(), P = mh.SetCount();
            std::vector<double> g, gl, buf((size_t)P * U);
            mh.fractions(g, gl);
            const auto w = mh.EngineUnitWeights();
            double sum = 0;
            const int slots[5] = {HBK_S_GROUND_WATER, HBK_S_GLACIER_WATER, HBK_S_SLOW, HBK_S_SLOW2, HBK_S_QUICK};
            for (int k = 0; k < 5; ++k) {
                mh.GetState(slots[k], buf.data());
                for (int u = 0; u < U; ++u) sum += buf[u] * (k == 0 ? g[u] : (k == 1 ? gl[u] : 1.0)) * w[u];
            }
            if (mh.compiled().st.snow_retention) {  // liquid water kept in the snowpacks: '<cover>_snowpack:water_content'
                const int more[2] = {HBK_S_GROUND_SNOWPACK_WATER, HBK_S_GLACIER_SNOWPACK_WATER};
                for (int k = 0; k < 2; ++k) {
                    mh.GetState(more[k], buf.data());
                    for (int u = 0; u < U; ++u) sum += buf[u] * (k == 0 ? g[u] : gl[u]) * w[u];
                }
            }
            if (mh.compiled().st.has_glacier) {
                std::vector<double> b((size_t)P * 2);
                mh.GetBasinState(b.data());
                sum += b[0] + b[1];
            }
            return sum;
        })
        .def("get_total_snow_storage_changes", [](hb::ModelHydro& mh) {
            const int U = mh.EngineUnitCount(), P = mh.SetCount();
            std::vector<double> g, gl, buf((size_t)P * U);
            mh.fractions(g, gl);
            const auto w = mh.EngineUnitWeights();
            double sum = 0;
            mh.GetState(HBK_S_GROUND_SNOW, buf.data());
            for (int u = 0; u < U; ++u) sum += buf[u] * g[u] * w[u];
            mh.GetState(HBK_S_GLACIER_SNOW, buf.data());
            for (int u = 0; u < U; ++u) sum += buf[u] * gl[u] * w[u];
            return sum;
        })
        .def("get_total_glacier_storage_changes", [](hb::ModelHydro& mh) {
            const int U = mh.EngineUnitCount(), P = mh.SetCount();
            std::vector<double> g, gl, buf((size_t)P * U);
            mh.fractions(g, gl);
            const auto w = mh.EngineUnitWeights();
            double sum = 0;
            mh.GetState(HBK_S_ICE, buf.data());
            for (int u = 0; u < U; ++u) sum += buf[u] * gl[u] * w[u];
            return sum;
        })
        // ModelHydro::DumpOutputs -> ResultWriter::WriteNetCDF (ResultWriter.cpp:8-87): the same variables under the same
        // names in <path>/results.nc (netCDF classic) and <path>/results.npz, written by hydrobricks_b200/results.py from
        // this object's getters
        .def("dump_outputs", [](py::object self, const std::string& path) -> bool {
            return py::module_::import("hydrobricks_b200.results").attr("write_results")(self, path).cast<bool>();
        }, "path"_a)
        .def("get_hydro_unit_structure_ids", [](hb::ModelHydro& mh) {
            auto v = mh.GetHydroUnitStructureIds();
            return i32arr(v.size(), v.data());
        })
        .def("start_mjd", &hb::ModelHydro::StartMjd)
        .def("time_step_days", &hb::ModelHydro::TimeStepDays)
        .def("engine_handle", [](hb::ModelHydro& mh) { return (std::uintptr_t)mh.engine(); },
             "Raw hbk_engine* (for hydrobricks_b200.engine.Engine.from_handle: device pointers, stream, statistics).");

    // The structure compiler alone (no engine, no device): what hb::Compiled derives from a SettingsModel. Lets the CPU
    // test suite compare it with the Python compiler (hydrobricks_b200/structure.py) on every reference structure.
    m.def(
        "compile_structure",
        [](hb::SettingsModel& sm, double dt_days, double start_mjd) {
            std::unique_ptr<hb::Compiled> c;
            try {
                c.reset(new hb::Compiled(sm, dt_days, start_mjd));
            } catch (const hb::Unsupported& err) {
                throw py::value_error(std::string("Model initialization failed: ") + err.what());
            }
            py::dict st;
            st["solver"] = c->st.solver;
            st["n_soil_stores"] = c->st.n_soil_stores;
            st["quick_kind"] = c->st.quick_kind;
            st["slow_process_order"] = c->st.slow_process_order;
            st["splitter_kind"] = c->st.splitter_kind;
            st["has_glacier"] = c->st.has_glacier;
            st["glacier_infinite_storage"] = c->st.glacier_infinite_storage;
            st["glacier_no_melt_when_snow_cover"] = c->st.glacier_no_melt_when_snow_cover;
            st["time_step_days"] = c->st.time_step_days;
            st["snow_ice_kind"] = c->st.snow_ice_kind;
            st["north_hemisphere"] = c->st.north_hemisphere;
            st["start_mjd"] = c->st.start_mjd;
            st["sublimation_kind"] = c->st.sublimation_kind;
            st["melt_kind"] = c->st.melt_kind;
            st["snow_slide_kind"] = c->st.snow_slide_kind;
            py::dict out;
            out["structure"] = st;
            out["params"] = f32arr(HBK_N_PARAMS, c->params);
            py::dict labels, index;
            for (auto& kv : c->labels) labels[py::str(kv.first)] = kv.second;
            for (auto& kv : c->paramIndex) index[py::make_tuple(kv.first.first, kv.first.second)] = kv.second;
            out["labels"] = labels;
            out["param_index"] = index;
            out["ground_name"] = c->groundName;
            out["glacier_name"] = c->glacierName;
            out["glacier2_name"] = c->glacier2Name;
            py::list twins;
            for (auto& l : c->twinLabels) twins.append(l);
            out["twin_labels"] = twins;
            return out;
        },
        "model_settings"_a, "time_step_days"_a = 1.0, "start_mjd"_a = 44605.0);

    py::class_<LogNull>(m, "LogNull").def(py::init<>());
}
