// hb_host.hpp — C++ host side above the C-ABI (include/hbk.h): the classes the reference exposes through its
// pybind11 module for the ModelHydro::Run() path, re-implemented for the B200 engine.
//
//   hb::SettingsModel   base/SettingsModel.{h,cpp}   declarative select-then-add structure builder
//   hb::SettingsBasin   base/SettingsBasin.{h,cpp}
//   hb::Compiled        structure compiler: SettingsModel -> per-structure process table (hbk_structure),
//                       float32 parameter slots, log-label -> recorder-slot map (ModelBuilder.cpp:36-240, 448-727
//                       is the wiring it recognises)
//   hb::ModelHydro      base/ModelHydro.{h,cpp} on top of hbk_engine; actions scheduled like
//                       TimeMachine::IncrementTime -> ActionsManager::DateUpdate (ActionsManager.cpp:80-104)
// Same method names and error behaviour as the reference classes (std::runtime_error where the reference throws,
// std::invalid_argument -> ValueError where its binding raises py::value_error). Header-only; hb_binding.cpp wraps
// it for Python. Nothing here computes hydrology: all arithmetic is in the CUDA engine.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hbk.h"

namespace hb {

using std::string;
using std::vector;

struct Param {
    string name;
    float value;
};
struct Output {
    string target;
    string fluxType = "water";
    bool instantaneous = false;
    bool isStatic = false;
};
struct Process {
    string name, kind;
    vector<string> forcing, logItems;
    vector<Output> outputs;
    vector<Param> parameters;
};
struct Brick {
    string name, kind, parent, attach;
    bool computedDirectly = false, isLandCover = false, isSurfaceComponent = false;
    vector<string> forcing, logItems;
    vector<Param> parameters;
    vector<Process> processes;
};
struct Splitter {
    string name, kind, attach;
    vector<string> forcing, logItems;
    vector<Output> outputs;
    vector<Param> parameters;
};
struct Structure {
    int id = 1;
    vector<string> logItems;
    vector<Brick> hydroUnitBricks, subBasinBricks;
    vector<Splitter> hydroUnitSplitters, subBasinSplitters;
};

inline bool isForcingName(const string& n) {
    static const std::set<string> k = {"precipitation", "pet", "temperature", "temperature_min", "temperature_max",
                                       "solar_radiation", "r_solar"};
    return k.count(n) > 0;
}

// Process::RegisterSettings of the reference (processes/Process*.cpp RegisterProcessSettings): default parameters
// and forcing of the process types the engine runs.
struct ProcessDefaults {
    vector<Param> params;
    vector<string> forcing;
};
inline const std::map<string, ProcessDefaults>& processRegistry() {
    static const std::map<string, ProcessDefaults> r = {
        {"melt:degree_day", {{{"degree_day_factor", 3.0f}, {"melting_temperature", 0.0f}}, {"temperature"}}},
        // ProcessMeltDegreeDayAspect.cpp:13-19, ProcessMeltTemperatureIndex.cpp:14-20
        {"melt:degree_day_aspect",
         {{{"degree_day_factor_n", 3.0f}, {"degree_day_factor_s", 3.0f}, {"degree_day_factor_ew", 3.0f},
           {"melting_temperature", 0.0f}},
          {"temperature"}}},
        {"melt:temperature_index",
         {{{"melt_factor", 3.0f}, {"melting_temperature", 0.0f}, {"radiation_coefficient", 0.0007f}},
          {"temperature", "solar_radiation"}}},
        {"et:socont", {{}, {"pet"}}},
        {"infiltration:socont", {{}, {}}},
        {"runoff:socont", {{{"beta", 500.0f}}, {}}},
        {"outflow:linear", {{{"response_factor", 0.2f}}, {}}},
        {"outflow:direct", {{}, {}}},
        {"outflow:rest", {{}, {}}},
        {"overflow", {{}, {}}},
        {"outflow:overflow", {{}, {}}},
        {"percolation:constant", {{{"percolation_rate", 0.1f}}, {}}},
        // ProcessTransformSnowToIce{Constant,Swat}::RegisterProcessSettings; both spellings (Process.cpp:65-66)
        {"transform:snow_ice_constant", {{{"snow_ice_transformation_rate", 0.5f}}, {}}},
        {"transformation:snow_ice_constant", {{{"snow_ice_transformation_rate", 0.5f}}, {}}},
        {"transform:snow_ice_swat", {{{"snow_ice_transformation_basal_acc_coeff", 0.0014f}, {"north_hemisphere", 1.0f}}, {}}},
        {"transformation:snow_ice_swat",
         {{{"snow_ice_transformation_basal_acc_coeff", 0.0014f}, {"north_hemisphere", 1.0f}}, {}}},
        // ProcessLateralSnowSlide::RegisterProcessSettings (ProcessLateralSnowSlide.cpp:24-31)
        {"transport:snow_slide",
         {{{"coeff", 3178.4f}, {"exp", -1.998f}, {"min_slope", 10.0f}, {"max_slope", 75.0f},
           {"min_snow_holding_depth", 50.0f}, {"max_snow_depth", 20000.0f}},
          {}}},
        // ProcessSublimation{Constant,PET}::RegisterProcessSettings
        {"sublimation:constant", {{{"sublimation_rate", 0.1f}}, {}}},
        {"sublimation:pet", {{{"sublimation_pet_factor", 0.2f}}, {"pet"}}},
        // ProcessOutflowSnowHolding.cpp:15-17, ProcessRefreezeDegreeDay.cpp:14-17
        {"outflow:snow_holding", {{{"water_holding_capacity", 0.1f}}, {}}},
        {"refreeze:degree_day", {{{"refreezing_factor", 0.05f}}, {"temperature"}}},
    };
    return r;
}

class SettingsModel {
  public:
    SettingsModel() {
        _structures.emplace_back();
        _structures.back().id = 1;
        _selStructure = 0;
    }

    // ---- options
    void SetLogAll(bool v = true) { _logAll = v; }
    void SetRecordFractions(bool v = true) { _recordFractions = v; }
    bool LogAll() const { return _logAll; }
    bool RecordsFractions() const { return _recordFractions; }
    void SetSolver(const string& name) { solver = name; }
    void SetTimer(const string& start, const string& end, int step, const string& unit) {
        timerStart = start;
        timerEnd = end;
        timeStep = step;
        timeStepUnit = unit;
    }
    void SetSpinupDays(int days) {
        if (days < 0) throw std::invalid_argument("The spin-up duration cannot be negative.");
        spinupDays = days;
    }

    // ---- structures
    int AddStructure() {
        int id = 0;
        for (auto& s : _structures) id = std::max(id, s.id);
        _structures.emplace_back();
        _structures.back().id = id + 1;
        _selStructure = (int)_structures.size() - 1;
        _selBrick = {-1, -1};
        _selProcess = -1;
        _selSplitter = -1;
        return id + 1;
    }
    int GetStructureCount() const { return (int)_structures.size(); }
    const vector<Structure>& Structures() const { return _structures; }

    // ---- bricks
    void AddHydroUnitBrick(const string& name, const string& kind = "storage") { addBrick(false, name, kind); }
    void AddSubBasinBrick(const string& name, const string& kind = "storage") { addBrick(true, name, kind); }
    void AddLandCoverBrick(const string& name, const string& kind) {
        addBrick(false, name, kind);
        brick().isLandCover = true;
        if (selectSplitterIfFound("rain_splitter")) addSplitterOutput(name, "water");  // SettingsModel.cpp:85-94
    }
    void SelectHydroUnitBrickByName(const string& name) {
        auto& v = sel().hydroUnitBricks;
        for (int i = 0; i < (int)v.size(); ++i) {
            if (v[i].name == name) {
                _selBrick = {0, i};
                _selProcess = -1;
                return;
            }
        }
        throw std::runtime_error("The hydro unit brick '" + name + "' was not found");
    }
    void AddBrickParameter(const string& name, float value, const string& kind = "constant") {
        if (kind != "constant")
            throw std::runtime_error("SettingsModel::AddBrickParameter - Parameter type '" + kind + "' not supported");
        brick().parameters.push_back({name, value});
    }
    void AddBrickForcing(const string& name) {
        if (!isForcingName(name)) throw std::invalid_argument("The provided forcing '" + name + "' is not yet supported.");
        brick().forcing.push_back(name);
    }
    void SetCurrentBrickComputedDirectly() { brick().computedDirectly = true; }
    void AddBrickLogging(const string& name) { addUnique(brick().logItems, name); }
    void AddLoggingToItem(const string& name) { addUnique(sel().logItems, name); }

    // ---- processes
    void AddBrickProcess(const string& name, const string& kind, const string& target = "", bool log = false) {
        auto it = processRegistry().find(kind);
        if (it == processRegistry().end())
            throw std::runtime_error("Process type '" + kind +
                                     "' is not available in the B200 engine (Socont family only).");
        Process p;
        p.name = name;
        p.kind = kind;
        brick().processes.push_back(p);
        _selProcess = (int)brick().processes.size() - 1;
        if (!target.empty()) {
            auto pos = target.find(':');
            if (pos != string::npos) {
                AddProcessOutput(target.substr(0, pos), target.substr(pos + 1));
            } else {
                AddProcessOutput(target);
            }
        }
        if ((log || _logAll) && kind.find("transport:") == string::npos) AddProcessLogging("output");
        for (auto& pr : it->second.params) AddProcessParameter(pr.name, pr.value);
        for (auto& f : it->second.forcing) AddProcessForcing(f);
    }
    void SelectProcess(const string& name) {
        auto& v = brick().processes;
        for (int i = 0; i < (int)v.size(); ++i) {
            if (v[i].name == name) {
                _selProcess = i;
                return;
            }
        }
        throw std::runtime_error("The process '" + name + "' was not found.");
    }
    void AddProcessOutput(const string& target, const string& fluxType = "water") {
        Output o;
        o.target = target;
        o.fluxType = fluxType;
        process().outputs.push_back(o);
    }
    void AddProcessParameter(const string& name, float value, const string& kind = "constant") {
        if (kind != "constant")
            throw std::runtime_error("SettingsModel::AddProcessParameter - Parameter type '" + kind + "' not supported");
        for (auto& p : process().parameters) {
            if (p.name == name) {
                p.value = value;
                return;
            }
        }
        process().parameters.push_back({name, value});
    }
    void SetProcessParameterValue(const string& name, float value, const string& = "constant") {
        for (auto& p : process().parameters) {
            if (p.name == name) {
                p.value = value;
                return;
            }
        }
        throw std::runtime_error("SettingsModel::SetProcessParameterValue - Parameter '" + name + "' not found");
    }
    void AddProcessForcing(const string& name) {
        if (!isForcingName(name)) throw std::invalid_argument("The provided forcing '" + name + "' is not yet supported.");
        process().forcing.push_back(name);
    }
    void AddProcessLogging(const string& name) { addUnique(process().logItems, name); }
    void SetProcessOutputsAsInstantaneous() {
        for (auto& o : process().outputs) o.instantaneous = true;
    }
    void SetProcessOutputsAsStatic() {
        for (auto& o : process().outputs) o.isStatic = true;
    }

    // ---- generators (SettingsModel.cpp:491-545)
    void GeneratePrecipitationSplitters(bool withSnow = true, const string& type = "snow_rain:linear") {
        if (!withSnow) {  // SplitterRain (SplitterRain.cpp:37-39): all precipitation is rain, no snowpacks
            addSplitter("rain", "rain");
            splitter().forcing = {"precipitation"};
            addSplitterOutput("rain_splitter", "water");
            splitter().parameters.push_back({"rain_correction_factor", 1.0f});
            addSplitter("rain_splitter", "multi_fluxes");
            return;
        }
        if (type != "snow_rain:linear" && type != "snow_rain:threshold")
            throw std::runtime_error("Splitter type '" + type + "' is not available in the B200 engine.");
        addSplitter("snow_rain_transition", type);
        splitter().forcing = {"precipitation", "temperature"};
        addSplitterOutput("rain_splitter", "water");
        addSplitterOutput("snow_splitter", "snow");
        if (type == "snow_rain:threshold") {
            splitter().parameters.push_back({"threshold", 0.0f});
        } else {
            splitter().parameters.push_back({"transition_start", 0.0f});
            splitter().parameters.push_back({"transition_end", 2.0f});
        }
        splitter().parameters.push_back({"rain_correction_factor", 1.0f});
        splitter().parameters.push_back({"snow_correction_factor", 1.0f});
        addSplitter("snow_splitter", "multi_fluxes");
        addSplitter("rain_splitter", "multi_fluxes");
    }
    void GenerateSnowpacks(const string& meltProcess) {
        vector<string> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back(b.name);
        for (auto& cover : covers) {
            if (!selectSplitterIfFound("snow_splitter"))
                throw std::runtime_error("The hydro unit splitter 'snow_splitter' was not found");
            addSplitterOutput(cover + "_snowpack", "snow");
            addBrick(false, cover + "_snowpack", "snowpack");
            brick().isSurfaceComponent = true;
            brick().parent = cover;
            AddBrickProcess("melt", meltProcess, cover);
        }
    }
    // SettingsModel::GenerateSnowpacksWithWaterRetention (SettingsModel.cpp:608-634): the melt goes to the snowpack's own
    // water container (instantaneous), outflowProcess (outflow:snow_holding) releases it to the land cover; with
    // rainToSnowpack the rain splitter feeds the snowpack instead of the cover
    void GenerateSnowpacksWithWaterRetention(const string& meltProcess, const string& outflowProcess,
                                             bool rainToSnowpack = false) {
        vector<string> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back(b.name);
        for (auto& cover : covers) {
            if (!selectSplitterIfFound("snow_splitter"))
                throw std::runtime_error("The hydro unit splitter 'snow_splitter' was not found");
            addSplitterOutput(cover + "_snowpack", "snow");
            addBrick(false, cover + "_snowpack", "snowpack");
            brick().isSurfaceComponent = true;
            brick().parent = cover;
            AddBrickProcess("melt", meltProcess);
            Output same;  // OutputProcessToSameBrick (SettingsModel.cpp:295-303)
            same.target = cover + "_snowpack";
            same.instantaneous = true;
            process().outputs.push_back(same);
            AddBrickProcess("meltwater", outflowProcess);
            AddProcessOutput(cover);
            if (rainToSnowpack) {
                if (!selectSplitterIfFound("rain_splitter"))
                    throw std::runtime_error("The hydro unit splitter 'rain_splitter' was not found");
                for (auto& o : splitter().outputs)  // ChangeSplitterOutputTargetIfFound (SettingsModel.cpp:401-412)
                    if (o.target == cover) {
                        o.target = cover + "_snowpack";
                        break;
                    }
            }
        }
    }
    // SettingsModel::AddSnowpackRefreezing (SettingsModel.cpp:574-590): from the snowpack's water container back to its
    // snow container, instantaneous
    void AddSnowpackRefreezing(const string& refreezingProcess = "refreeze:degree_day") {
        vector<string> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back(b.name);
        for (auto& c : covers) {
            SelectHydroUnitBrickByName(c + "_snowpack");
            AddBrickProcess("refreeze", refreezingProcess, c + "_snowpack:snow");
            SetProcessOutputsAsInstantaneous();
        }
    }
    // SettingsModel::AddSnowIceTransformation (SettingsModel.cpp:547-559)
    void AddSnowIceTransformation(const string& transformationProcess = "transform:snow_ice_swat") {
        vector<std::pair<string, string>> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back({b.name, b.kind});
        for (auto& c : covers) {
            SelectHydroUnitBrickByName(c.first + "_snowpack");
            if (c.second == "glacier") AddBrickProcess("snow_ice_transfo", transformationProcess, c.first + ":ice");
        }
    }
    // SettingsModel::AddSnowRedistribution (SettingsModel.cpp:560-572): a lateral process on the snowpack of every land
    // cover (the glacier covers only without skipGlaciers), instantaneous snow fluxes to the connected units' snowpacks
    void AddSnowRedistribution(const string& redistributionProcess = "transport:snow_slide", bool skipGlaciers = false) {  // default as SettingsModel.h:391
        if (redistributionProcess != "transport:snow_slide")
            throw std::runtime_error("Process type '" + redistributionProcess + "' is not available in the B200 engine.");
        vector<std::pair<string, string>> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back({b.name, b.kind});
        for (auto& c : covers) {
            if (skipGlaciers && c.second == "glacier") continue;
            SelectHydroUnitBrickByName(c.first + "_snowpack");
            AddBrickProcess("snow_redistribution", redistributionProcess, "lateral:snow");
            SetProcessOutputsAsInstantaneous();
        }
    }
    // SettingsModel::AddSnowpackSublimation (SettingsModel.cpp:586-598): no target = a flux to the atmosphere
    void AddSnowpackSublimation(const string& sublimationProcess) {
        vector<string> covers;
        for (auto& b : sel().hydroUnitBricks)
            if (b.isLandCover) covers.push_back(b.name);
        for (auto& c : covers) {
            SelectHydroUnitBrickByName(c + "_snowpack");
            AddBrickProcess("sublimation", sublimationProcess);
        }
    }
    [[noreturn]] void Unsupported(const string& what) const {
        throw std::runtime_error(what + " is not available in the B200 engine (SURVEY.md §8f lists it as next).");
    }

    // ---- parameters (SettingsModel.cpp:923-1031)
    bool SetParameterValue(const string& component, const string& name, float value) {
        if (component.find(',') != string::npos) {
            size_t start = 0;
            while (start <= component.size()) {
                size_t end = component.find(',', start);
                if (end == string::npos) end = component.size();
                string tok = component.substr(start, end - start);
                size_t a = tok.find_first_not_of(' '), b = tok.find_last_not_of(' ');
                if (a != string::npos) {
                    if (!SetParameterValue(tok.substr(a, b - a + 1), name, value)) return false;
                }
                start = end + 1;
            }
            return true;
        }
        bool found = false;
        for (auto& st : _structures) {
            Brick* b = findBrick(st, component);
            if (b) {
                setInBrick(*b, name, value);
                found = true;
                continue;
            }
            Splitter* sp = findSplitter(st, component);
            if (sp) {
                bool ok = false;
                for (auto& p : sp->parameters) {
                    if (p.name == name) {
                        p.value = value;
                        ok = true;
                        break;
                    }
                }
                if (!ok) throw std::runtime_error("SettingsModel::SetSplitterParameterValue - Parameter '" + name + "' not found");
                found = true;
                continue;
            }
            if (component.find("type:") != string::npos) {
                string kind = component.substr(component.find(':') + 1);
                auto matches = [&](const Brick& br) {
                    return br.kind == kind ||
                           (kind == "land_cover" && (br.kind == "generic_land_cover" || br.kind == "glacier"));
                };
                for (auto& br : st.hydroUnitBricks)
                    if (matches(br)) setInBrick(br, name, value);
                for (auto& br : st.subBasinBricks)
                    if (matches(br)) setInBrick(br, name, value);
                found = true;  // SetParameterValueInSelectedStructure falls through to `return true`
            }
        }
        return found;
    }

    // SettingsModel::GetHydroUnitLogLabels (SettingsModel.cpp:808-841): union over the variants, in order
    vector<string> GetHydroUnitLogLabels() const {
        vector<string> labels;
        std::set<string> seen;
        auto add = [&](const string& l) {
            if (seen.insert(l).second) labels.push_back(l);
        };
        for (auto& st : _structures) {
            for (auto& b : st.hydroUnitBricks) {
                for (auto& item : b.logItems) add(b.name + ":" + item);
                for (auto& p : b.processes)
                    for (auto& item : p.logItems) add(b.name + ":" + p.name + ":" + item);
            }
        }
        return labels;
    }
    vector<string> GetLandCoverBricksNames() const {
        vector<string> names;
        std::set<string> seen;
        for (auto& st : _structures)
            for (auto& b : st.hydroUnitBricks)
                if (b.isLandCover && seen.insert(b.name).second) names.push_back(b.name);
        return names;
    }

    string solver, timerStart, timerEnd, timeStepUnit;
    int timeStep = 1, spinupDays = 0;

  private:
    bool _logAll = false, _recordFractions = false;
    vector<Structure> _structures;
    int _selStructure = 0;
    std::pair<int, int> _selBrick{-1, -1};  // (0 hydro unit | 1 sub-basin, index)
    int _selProcess = -1, _selSplitter = -1;

    Structure& sel() { return _structures[_selStructure]; }
    Brick& brick() {
        if (_selBrick.second < 0) throw std::runtime_error("No brick selected.");
        return _selBrick.first == 0 ? sel().hydroUnitBricks[_selBrick.second] : sel().subBasinBricks[_selBrick.second];
    }
    Process& process() {
        if (_selProcess < 0) throw std::runtime_error("No process selected.");
        return brick().processes[_selProcess];
    }
    Splitter& splitter() { return sel().hydroUnitSplitters[_selSplitter]; }
    static void addUnique(vector<string>& v, const string& s) {
        if (std::find(v.begin(), v.end(), s) == v.end()) v.push_back(s);
    }
    void addBrick(bool subBasin, const string& name, const string& kind) {
        Brick b;
        b.name = name;
        b.kind = kind;
        b.attach = subBasin ? "sub_basin" : "hydro_unit";
        auto& v = subBasin ? sel().subBasinBricks : sel().hydroUnitBricks;
        v.push_back(b);
        _selBrick = {subBasin ? 1 : 0, (int)v.size() - 1};
        _selProcess = -1;
        if (_logAll) {  // SettingsModel.cpp:55-62
            AddBrickLogging("water_content");
            if (kind == "glacier") AddBrickLogging("ice_content");
            if (kind == "snowpack") AddBrickLogging("snow_content");
        }
    }
    void addSplitter(const string& name, const string& kind) {
        Splitter s;
        s.name = name;
        s.kind = kind;
        s.attach = "hydro_unit";
        sel().hydroUnitSplitters.push_back(s);
        _selSplitter = (int)sel().hydroUnitSplitters.size() - 1;
    }
    bool selectSplitterIfFound(const string& name) {
        auto& v = sel().hydroUnitSplitters;
        for (int i = 0; i < (int)v.size(); ++i) {
            if (v[i].name == name) {
                _selSplitter = i;
                return true;
            }
        }
        return false;
    }
    void addSplitterOutput(const string& target, const string& fluxType) {
        Output o;
        o.target = target;
        o.fluxType = fluxType;
        splitter().outputs.push_back(o);
    }
    static Brick* findBrick(Structure& st, const string& name) {
        for (auto& b : st.hydroUnitBricks)
            if (b.name == name) return &b;
        for (auto& b : st.subBasinBricks)
            if (b.name == name) return &b;
        return nullptr;
    }
    static Splitter* findSplitter(Structure& st, const string& name) {
        for (auto& s : st.hydroUnitSplitters)
            if (s.name == name) return &s;
        for (auto& s : st.subBasinSplitters)
            if (s.name == name) return &s;
        return nullptr;
    }
    static void setInBrick(Brick& b, const string& name, float value) {
        for (auto& p : b.parameters) {
            if (p.name == name) {
                p.value = value;
                return;
            }
        }
        for (auto& pr : b.processes) {
            for (auto& p : pr.parameters) {
                if (p.name == name) {
                    p.value = value;
                    return;
                }
            }
        }
        throw std::runtime_error("The parameter '" + name + "' was not found.");
    }
};

// ---------------------------------------------------------------------------------------------------------
struct UnitSettings {
    int id = 0;
    double area = 0, elevation = -9999;
    bool hasSlope = false;
    double slopeValue = 0;
    string slopeUnit;
    vector<std::pair<string, double>> landCovers;
    string aspectClass;  // property "aspect_class" (melt:degree_day_aspect)
};

class SettingsBasin {
  public:
    void AddHydroUnit(int id, double area, double elevation = -9999) {
        UnitSettings u;
        u.id = id;
        u.area = area;
        u.elevation = elevation;
        units.push_back(u);
    }
    void AddLandCover(const string& name, const string&, double fraction) { last().landCovers.push_back({name, fraction}); }
    void AddHydroUnitPropertyString(const string& name, const string& value) {
        if (name == "aspect_class") last().aspectClass = value;
    }
    void AddHydroUnitPropertyDouble(const string& name, double value, const string& unit) {
        if (name == "slope") {
            last().hasSlope = true;
            last().slopeValue = value;
            last().slopeUnit = unit;
        }
    }
    // SettingsBasin::AddLateralConnection: kept by unit id, resolved to basin positions when the model is built
    struct Connection {
        int giver, receiver;
        double fraction;
    };
    void AddLateralConnection(int giver, int receiver, double fraction, const string& = "") {
        connections.push_back({giver, receiver, fraction});
    }
    int GetLateralConnectionCount() const { return (int)connections.size(); }
    void Clear() {
        units.clear();
        connections.clear();
    }
    vector<UnitSettings> units;
    vector<Connection> connections;

  private:
    UnitSettings& last() {
        if (units.empty()) throw std::runtime_error("No hydro unit defined.");
        return units.back();
    }
};

// ---------------------------------------------------------------------------------------------------------
struct Unsupported : std::invalid_argument {
    using std::invalid_argument::invalid_argument;
};

inline const std::map<string, int>& recordSlots() {
    static const std::map<string, int> m = {
        {"ground_water", HBK_R_GROUND_WATER}, {"ground_infiltration", HBK_R_GROUND_INFILTRATION},
        {"ground_runoff", HBK_R_GROUND_RUNOFF}, {"glacier_water", HBK_R_GLACIER_WATER},
        {"glacier_ice", HBK_R_GLACIER_ICE}, {"glacier_outflow", HBK_R_GLACIER_OUTFLOW},
        {"glacier_melt", HBK_R_GLACIER_MELT}, {"ground_snowpack_water", HBK_R_GROUND_SNOWPACK_WATER},
        {"ground_snowpack_snow", HBK_R_GROUND_SNOWPACK_SNOW}, {"ground_snowpack_melt", HBK_R_GROUND_SNOWPACK_MELT},
        {"glacier_snowpack_water", HBK_R_GLACIER_SNOWPACK_WATER}, {"glacier_snowpack_snow", HBK_R_GLACIER_SNOWPACK_SNOW},
        {"glacier_snowpack_melt", HBK_R_GLACIER_SNOWPACK_MELT}, {"slow_water", HBK_R_SLOW_WATER},
        {"slow_et", HBK_R_SLOW_ET}, {"slow_outflow", HBK_R_SLOW_OUTFLOW}, {"slow_overflow", HBK_R_SLOW_OVERFLOW},
        {"slow_percolation", HBK_R_SLOW_PERCOLATION}, {"slow2_water", HBK_R_SLOW2_WATER},
        {"slow2_outflow", HBK_R_SLOW2_OUTFLOW}, {"quick_water", HBK_R_QUICK_WATER},
        {"quick_runoff", HBK_R_QUICK_RUNOFF}, {"fraction_ground", HBK_R_FRACTION_GROUND},
        {"fraction_glacier", HBK_R_FRACTION_GLACIER}, {"glacier_snowpack_transfo", HBK_R_GLACIER_SNOWPACK_TRANSFO},
        {"ground_snowpack_sublimation", HBK_R_GROUND_SNOWPACK_SUBLIMATION},
        {"glacier_snowpack_sublimation", HBK_R_GLACIER_SNOWPACK_SUBLIMATION},
        {"ground_snowpack_meltwater", HBK_R_GROUND_SNOWPACK_MELTWATER},
        {"glacier_snowpack_meltwater", HBK_R_GLACIER_SNOWPACK_MELTWATER},
        {"ground_snowpack_refreeze", HBK_R_GROUND_SNOWPACK_REFREEZE},
        {"glacier_snowpack_refreeze", HBK_R_GLACIER_SNOWPACK_REFREEZE}};
    return m;
}

// Structure compiler: recognises the Socont family by the ROLE of each brick (python/src/hydrobricks/models/
// socont.py:111-215, modules/glacier.py:87-131, core/tests/src/helpers.cpp:9-107) and refuses anything else.
struct Compiled {
    hbk_structure st{};
    float params[HBK_N_PARAMS] = {};
    std::map<std::pair<string, string>, int> paramIndex;
    std::map<string, int> labels;  // log label -> HBK_R_* slot
    string groundName, glacierName, glacier2Name;
    std::set<string> twinLabels;  // log labels of the second glacier cover: recorded in the twin units' columns
    int glacierStructureId = -1;
    std::map<int, std::set<string>> variantCovers;
    vector<std::pair<string, string>> brickKinds;  // (name, kind) over all variants, for "type:" components

    static int sublimationKindOf(const string& kind) {
        if (kind == "sublimation:constant") return HBK_SUBLIMATION_CONSTANT;
        if (kind == "sublimation:pet") return HBK_SUBLIMATION_PET;
        return HBK_SUBLIMATION_NONE;
    }
    static int meltKindOf(const string& kind) {
        if (kind == "melt:degree_day") return HBK_MELT_DEGREE_DAY;
        if (kind == "melt:degree_day_aspect") return HBK_MELT_DEGREE_DAY_ASPECT;
        if (kind == "melt:temperature_index") return HBK_MELT_TEMPERATURE_INDEX;
        return -1;
    }
    // HydroUnit property "aspect_class" -> HBK_ASPECT_* (ProcessMeltDegreeDayAspect.cpp:38-58)
    static int aspectOf(const string& a) {
        if (a == "north" || a == "N") return HBK_ASPECT_NORTH;
        if (a == "south" || a == "S") return HBK_ASPECT_SOUTH;
        if (a == "east" || a == "west" || a == "E" || a == "W") return HBK_ASPECT_EAST_WEST;
        throw std::invalid_argument("Invalid aspect: " + a);
    }
    bool noSnow = false;
    bool rainToSnowpack = false;  // the rain splitter feeds the snowpacks (GenerateSnowpacksWithWaterRetention)
    // The melt process of a snowpack or of the glacier ice: its kind (one per structure) and its parameters.
    void setMelt(const string& component, const Process& proc, int slotA, int slotTm, int slotB, int slotC) {
        const int kind = meltKindOf(proc.kind);
        if (kind < 0) throw Unsupported(component + ": melt process " + proc.kind);
        if (st.melt_kind != -1 && st.melt_kind != kind)
            throw Unsupported("different melt processes on the snowpacks / the glacier ice");
        st.melt_kind = kind;
        set(slotTm, component, "melting_temperature", paramOf(proc.parameters, "melting_temperature"));
        if (kind == HBK_MELT_DEGREE_DAY) {
            set(slotA, component, "degree_day_factor", paramOf(proc.parameters, "degree_day_factor"));
        } else if (kind == HBK_MELT_DEGREE_DAY_ASPECT) {
            set(slotA, component, "degree_day_factor_n", paramOf(proc.parameters, "degree_day_factor_n"));
            set(slotB, component, "degree_day_factor_s", paramOf(proc.parameters, "degree_day_factor_s"));
            bool hasEw, hasWe;
            const float ew = paramOf(proc.parameters, "degree_day_factor_ew", &hasEw);
            const float we = paramOf(proc.parameters, "degree_day_factor_we", &hasWe);
            if (!hasEw && !hasWe) throw Unsupported("Missing parameter 'degree_day_factor_ew' or 'degree_day_factor_we'");
            set(slotC, component, hasEw ? "degree_day_factor_ew" : "degree_day_factor_we", hasEw ? ew : we);
        } else {
            set(slotA, component, "melt_factor", paramOf(proc.parameters, "melt_factor"));
            set(slotB, component, "radiation_coefficient", paramOf(proc.parameters, "radiation_coefficient"));
        }
    }
    static int snowIceKindOf(const string& kind) {
        if (kind == "transform:snow_ice_constant" || kind == "transformation:snow_ice_constant") return HBK_SNOW_ICE_CONSTANT;
        if (kind == "transform:snow_ice_swat" || kind == "transformation:snow_ice_swat") return HBK_SNOW_ICE_SWAT;
        return HBK_SNOW_ICE_NONE;
    }
    static float paramOf(const vector<Param>& v, const string& name, bool* found = nullptr) {
        for (auto& p : v) {
            if (p.name == name) {
                if (found) *found = true;
                return p.value;
            }
        }
        if (found) {
            *found = false;
            return 0;
        }
        throw Unsupported("parameter '" + name + "' missing");
    }
    static vector<string> kinds(const Brick& b) {
        vector<string> k;
        for (auto& p : b.processes) k.push_back(p.kind == "outflow:overflow" ? "overflow" : p.kind);
        return k;
    }
    static string target(const Process& p) { return p.outputs.empty() ? "" : p.outputs[0].target; }
    void set(int slot, const string& component, const string& name, float value) {
        params[slot] = value;
        paramIndex[{component, name}] = slot;
    }
    void label(const string& l, const string& slot) { labels[l] = recordSlots().at(slot); }

    Compiled(const SettingsModel& sm, double dtDays, double startMjd = 44605.0) {
        const string& solver = sm.solver;
        if (solver == "euler_explicit") {
            st.solver = HBK_SOLVER_EULER;
        } else if (solver == "heun_explicit") {
            st.solver = HBK_SOLVER_HEUN;
        } else if (solver == "rk4" || solver == "runge_kutta") {
            st.solver = HBK_SOLVER_RK4;
        } else if (solver == "crank_nicolson" || solver == "trapezoidal") {  // Solver::Factory, Solver.cpp:42-64
            st.solver = HBK_SOLVER_CRANK_NICOLSON;
        } else if (solver == "implicit_euler" || solver == "euler_implicit") {
            st.solver = HBK_SOLVER_IMPLICIT_EULER;
        } else if (solver == "exponential_euler") {
            st.solver = HBK_SOLVER_EXPONENTIAL_EULER;
        } else if (solver == "analytic_linear" || solver == "analytic") {
            st.solver = HBK_SOLVER_ANALYTIC_LINEAR;
        } else {
            throw Unsupported("Incorrect solver name: " + solver + ". Valid solver names: rk4, runge_kutta, "
                              "euler_explicit, heun_explicit, analytic_linear, analytic, implicit_euler, "
                              "euler_implicit, crank_nicolson, trapezoidal, exponential_euler");
        }
        st.time_step_days = dtDays;
        st.snow_ice_kind = HBK_SNOW_ICE_NONE;
        st.sublimation_kind = HBK_SUBLIMATION_NONE;
        st.north_hemisphere = 1;
        st.start_mjd = startMjd;
        st.melt_kind = -1;
        const auto& structures = sm.Structures();
        if (structures.empty()) throw Unsupported("empty structure");
        if (structures.size() > 2) throw Unsupported("more than two structure variants");
        const Structure* base = nullptr;
        const Structure* glacierSt = nullptr;
        for (auto& s : structures) {
            if (s.id == 1) base = &s;
            for (auto& b : s.hydroUnitBricks) {
                if (b.kind == "glacier") glacierSt = &s;
                brickKinds.push_back({b.name, b.kind});
            }
            for (auto& b : s.subBasinBricks) brickKinds.push_back({b.name, b.kind});
            std::set<string> covers;
            for (auto& b : s.hydroUnitBricks)
                if (b.isLandCover) covers.insert(b.name);
            variantCovers[s.id] = covers;
        }
        if (!base) throw Unsupported("structure 1 missing");
        const Structure& full = glacierSt ? *glacierSt : *base;
        glacierStructureId = glacierSt ? glacierSt->id : -1;
        auto findBrick = [&](const string& n) -> const Brick* {
            for (auto& b : full.hydroUnitBricks)
                if (b.name == n) return &b;
            return nullptr;
        };
        auto findBasinBrick = [&](const string& n) -> const Brick* {
            for (auto& b : base->subBasinBricks)
                if (b.name == n) return &b;
            return nullptr;
        };

        // splitters
        if (!base->subBasinSplitters.empty()) throw Unsupported("sub-basin splitters");
        const Splitter* tr = nullptr;
        const Splitter* rainOnly = nullptr;
        int nMulti = 0;
        for (auto& s : full.hydroUnitSplitters) {
            if (s.kind == "snow_rain:linear" || s.kind == "snow_rain:threshold") tr = &s;
            if (s.kind == "rain") rainOnly = &s;
            if (s.kind == "multi_fluxes") nMulti++;
        }
        vector<string> rainTargets, snowTargets;
        bool f;
        noSnow = rainOnly && !tr;
        if (noSnow) {
            // SettingsModel::GeneratePrecipitationSplitters(withSnow = false) (SettingsModel.cpp:518-529): SplitterRain
            // (SplitterRain.cpp:37-39) sends precipitation x rain_correction_factor to the land covers; no snowpacks. On
            // the device this is the threshold splitter with a threshold no temperature reaches (-FLT_MAX): rain =
            // precipitation x rcf, snow = 0, and the snowpack code only ever moves zeros.
            if (nMulti != 1 || full.hydroUnitSplitters.size() != 2)
                throw Unsupported("expected the rain splitter + the rain multi_fluxes splitter");
            tr = rainOnly;
            st.splitter_kind = HBK_SPLIT_THRESHOLD;
            params[HBK_P_TRANSITION_START] = -3.402823466e+38f;
            params[HBK_P_SNOW_CORRECTION] = 1.0f;
            float rcf = paramOf(tr->parameters, "rain_correction_factor", &f);
            set(HBK_P_RAIN_CORRECTION, tr->name, "rain_correction_factor", f ? rcf : 1.0f);
            const Splitter* tgt = nullptr;
            if (tr->outputs.size() == 1)
                for (auto& s : full.hydroUnitSplitters)
                    if (s.name == tr->outputs[0].target && s.kind == "multi_fluxes") tgt = &s;
            if (!tgt) throw Unsupported("the rain splitter must feed a multi_fluxes splitter");
            for (auto& oo : tgt->outputs) rainTargets.push_back(oo.target);
        } else {
            if (!tr || nMulti != 2 || full.hydroUnitSplitters.size() != 3)
                throw Unsupported("expected snow_rain_transition + rain/snow multi_fluxes splitters");
            if (tr->kind == "snow_rain:linear") {
                st.splitter_kind = HBK_SPLIT_LINEAR;
                set(HBK_P_TRANSITION_START, tr->name, "transition_start", paramOf(tr->parameters, "transition_start"));
                set(HBK_P_TRANSITION_END, tr->name, "transition_end", paramOf(tr->parameters, "transition_end"));
            } else {
                st.splitter_kind = HBK_SPLIT_THRESHOLD;
                set(HBK_P_TRANSITION_START, tr->name, "threshold", paramOf(tr->parameters, "threshold"));
            }
            float rcf = paramOf(tr->parameters, "rain_correction_factor", &f);
            set(HBK_P_RAIN_CORRECTION, tr->name, "rain_correction_factor", f ? rcf : 1.0f);
            float scf = paramOf(tr->parameters, "snow_correction_factor", &f);
            set(HBK_P_SNOW_CORRECTION, tr->name, "snow_correction_factor", f ? scf : 1.0f);
            for (auto& o : tr->outputs) {
                const Splitter* tgt = nullptr;
                for (auto& s : full.hydroUnitSplitters)
                    if (s.name == o.target && s.kind == "multi_fluxes") tgt = &s;
                if (!tgt) throw Unsupported("snow_rain_transition must feed multi_fluxes splitters");
                for (auto& oo : tgt->outputs) (o.fluxType == "snow" ? snowTargets : rainTargets).push_back(oo.target);
            }
        }

        // land covers
        const Brick* ground = nullptr;
        const Brick* glacier = nullptr;
        // a second glacier cover (glacier_ice + glacier_debris, models/socont.py:131-140) is carried by twin units
        // (hbk_set_unit_twins): the same bricks once more with the HBK_P_GLACIER2_* / HBK_P_ICE2_* parameter slots
        const Brick* glacier2 = nullptr;
        int nCovers = 0;
        for (auto& b : full.hydroUnitBricks) {
            if (!b.isLandCover) continue;
            nCovers++;
            if (b.kind == "generic_land_cover") {
                if (ground) throw Unsupported("more than one generic land cover");
                ground = &b;
            } else if (b.kind == "glacier") {
                if (glacier2) throw Unsupported("more than two glacier covers");
                (glacier ? glacier2 : glacier) = &b;
            }
        }
        if (!ground || nCovers != 1 + (glacier ? 1 : 0) + (glacier2 ? 1 : 0))
            throw Unsupported("expected one generic land cover and at most two glacier covers");
        groundName = ground->name;
        glacierName = glacier ? glacier->name : "";
        glacier2Name = glacier2 ? glacier2->name : "";
        {
            vector<string> want = {ground->name};
            if (glacier) want.push_back(glacier->name);
            if (glacier2) want.push_back(glacier2->name);
            std::sort(want.begin(), want.end());
            std::sort(rainTargets.begin(), rainTargets.end());
            // ... or, with rain_to_snowpack, to the snowpack of every land cover (SettingsModel.cpp:624-632)
            vector<string> wantSnowpacks;
            for (auto& w : want) wantSnowpacks.push_back(w + "_snowpack");
            std::sort(wantSnowpacks.begin(), wantSnowpacks.end());
            rainToSnowpack = rainTargets == wantSnowpacks;
            if (!rainToSnowpack && want != rainTargets) throw Unsupported("rain must go to every land cover");
        }
        if (kinds(*ground) != vector<string>{"infiltration:socont", "outflow:rest"})
            throw Unsupported("land cover " + ground->name + ": expected infiltration:socont + outflow:rest");
        const string slowName = target(ground->processes[0]), quickName = target(ground->processes[1]);
        if (!ground->processes[1].outputs[0].isStatic) throw Unsupported("outflow:rest output must be static");
        label(ground->name + ":water_content", "ground_water");
        label(ground->name + ":" + ground->processes[0].name + ":output", "ground_infiltration");
        label(ground->name + ":" + ground->processes[1].name + ":output", "ground_runoff");

        // One snowpack per cover: melt:degree_day [, snow -> ice transformation on a glacier] [, sublimation], in the
        // order model_settings.py:255-263 adds them.
        struct SnowpackInfo {
            const Brick* brick;
            const Process* transfo;
            const Process* sublim;
            const Process* slide = nullptr;
            const Process* holding = nullptr;   // outflow:snow_holding (liquid water kept in the snowpack)
            const Process* refreeze = nullptr;  // refreeze:degree_day
        };
        auto snowpackOf = [&](const Brick& cover) -> SnowpackInfo {
            const Brick* sp = nullptr;
            for (auto& b : full.hydroUnitBricks) {
                if (b.kind == "snowpack" && b.parent == cover.name) {
                    if (sp) throw Unsupported("cover " + cover.name + ": more than one snowpack");
                    sp = &b;
                }
            }
            SnowpackInfo info{sp, nullptr, nullptr};
            bool ok = sp && !sp->processes.empty() && meltKindOf(sp->processes[0].kind) >= 0;
            if (ok) {
                size_t i = 1;
                if (target(sp->processes[0]) == sp->name) {
                    // SettingsModel::GenerateSnowpacksWithWaterRetention (SettingsModel.cpp:608-634): the melt stays in the
                    // snowpack (instantaneous flux to its own water container), outflow:snow_holding feeds the cover
                    const auto& mo = sp->processes[0].outputs;
                    ok = mo.size() == 1 && mo[0].instantaneous && mo[0].fluxType == "water" && i < sp->processes.size() &&
                         sp->processes[i].kind == "outflow:snow_holding";
                    if (ok) {
                        info.holding = &sp->processes[i++];
                        const auto& out = info.holding->outputs;
                        ok = out.size() == 1 && out[0].target == cover.name && out[0].fluxType == "water" &&
                             !out[0].instantaneous && !out[0].isStatic;
                    }
                    if (ok && i < sp->processes.size() && sp->processes[i].kind == "refreeze:degree_day") {
                        // SettingsModel::AddSnowpackRefreezing (SettingsModel.cpp:574-590)
                        info.refreeze = &sp->processes[i++];
                        const auto& out = info.refreeze->outputs;
                        ok = out.size() == 1 && out[0].target == sp->name && out[0].fluxType == "snow" && out[0].instantaneous;
                    }
                } else {
                    ok = target(sp->processes[0]) == cover.name;
                }
                if (ok && i < sp->processes.size() && cover.kind == "glacier" && snowIceKindOf(sp->processes[i].kind)) {
                    // SettingsModel::AddSnowIceTransformation: snow -> the glacier's ice container, a regular flux
                    info.transfo = &sp->processes[i++];
                    const auto& out = info.transfo->outputs;
                    ok = out.size() == 1 && out[0].target == cover.name && out[0].fluxType == "ice" &&
                         !out[0].instantaneous && !out[0].isStatic;
                }
                if (ok && i < sp->processes.size() && sp->processes[i].kind == "transport:snow_slide") {
                    // SettingsModel::AddSnowRedistribution: instantaneous snow fluxes to the connected units' snowpacks
                    info.slide = &sp->processes[i++];
                    const auto& out = info.slide->outputs;
                    ok = out.size() == 1 && out[0].target == "lateral" && out[0].fluxType == "snow" && out[0].instantaneous;
                }
                if (ok && i < sp->processes.size() && sublimationKindOf(sp->processes[i].kind)) {
                    info.sublim = &sp->processes[i++];
                    ok = info.sublim->outputs.empty();
                }
                ok = ok && i == sp->processes.size();
            }
            if (!ok)
                throw Unsupported("cover " + cover.name + ": expected one snowpack with melt:degree_day"
                                  "|degree_day_aspect|temperature_index [+ outflow:snow_holding [+ refreeze:degree_day]] "
                                  "[+ transform:snow_ice_constant|swat on a glacier] [+ sublimation:constant|pet]");
            return info;
        };
        // water_holding_capacity / refreezing_factor of a snowpack with liquid-water retention and its labels
        int nHolding = 0, nRefreeze = 0;
        auto setRetention = [&](const SnowpackInfo& info, int whcSlot, int cfrSlot, const string& prefix) {
            if (!info.holding) return;
            const Brick& sp = *info.brick;
            nHolding++;
            set(whcSlot, sp.name, "water_holding_capacity", paramOf(info.holding->parameters, "water_holding_capacity"));
            label(sp.name + ":" + info.holding->name + ":output", prefix + "_meltwater");
            if (info.refreeze) {
                nRefreeze++;
                set(cfrSlot, sp.name, "refreezing_factor", paramOf(info.refreeze->parameters, "refreezing_factor"));
                label(sp.name + ":" + info.refreeze->name + ":output", prefix + "_refreeze");
            }
        };
        auto setSublimation = [&](const Brick& sp, const Process& proc, int slot, const char* record) {
            const int kind = sublimationKindOf(proc.kind);
            if (st.sublimation_kind != HBK_SUBLIMATION_NONE && st.sublimation_kind != kind)
                throw Unsupported("different sublimation processes on the snowpacks");
            st.sublimation_kind = kind;
            const string pname = kind == HBK_SUBLIMATION_CONSTANT ? "sublimation_rate" : "sublimation_pet_factor";
            set(slot, sp.name, pname, paramOf(proc.parameters, pname));
            label(sp.name + ":" + proc.name + ":output", record);
        };
        vector<string> snowpacks;
        std::set<string> known = {ground->name, slowName, quickName};
        SnowpackInfo gInfo{nullptr, nullptr, nullptr};
        vector<std::pair<string, const Process*>> slideOn;  // snowpack name, its transport:snow_slide process
        if (!noSnow) {
            gInfo = snowpackOf(*ground);
            const Brick& gsp = *gInfo.brick;
            if (gInfo.sublim) setSublimation(gsp, *gInfo.sublim, HBK_P_GROUND_SUBLIMATION, "ground_snowpack_sublimation");
            setMelt(gsp.name, gsp.processes[0], HBK_P_GROUND_SNOW_DDF, HBK_P_GROUND_SNOW_TMELT, HBK_P_GROUND_SNOW_MELT_B,
                    HBK_P_GROUND_SNOW_MELT_C);
            label(gsp.name + ":water_content", "ground_snowpack_water");
            label(gsp.name + ":snow_content", "ground_snowpack_snow");
            label(gsp.name + ":" + gsp.processes[0].name + ":output", "ground_snowpack_melt");
            snowpacks.push_back(gsp.name);
            known.insert(gsp.name);
            if (gInfo.slide) slideOn.push_back({gsp.name, gInfo.slide});
            setRetention(gInfo, HBK_P_GROUND_SNOW_WHC, HBK_P_GROUND_SNOW_CFR, "ground_snowpack");
        }

        string b1, b2;
        const Brick* glacierCovers[2] = {glacier, glacier2};
        for (int second = 0; second < 2; ++second) {
            const Brick* cover = glacierCovers[second];
            if (!cover) continue;
            auto lab = [&](const string& l, const string& slot) {
                label(l, slot);
                if (second) twinLabels.insert(l);
            };
            st.has_glacier = 1;
            {
                const vector<string> gk = kinds(*cover);
                if (gk.size() != 2 || gk[0] != "outflow:direct" || meltKindOf(gk[1]) < 0)
                    throw Unsupported("glacier cover: expected outflow:direct + melt:degree_day|degree_day_aspect|"
                                      "temperature_index");
            }
            for (auto& p : cover->processes)
                if (p.outputs.empty() || !p.outputs[0].instantaneous) throw Unsupported("glacier outputs must be instantaneous");
            bool hasInf, hasNoMelt;
            float inf = paramOf(cover->parameters, "infinite_storage", &hasInf);
            float noMelt = paramOf(cover->parameters, "no_melt_when_snow_cover", &hasNoMelt);
            const int infinite = hasInf && std::fabs(inf) > 1.1920929e-07f ? 1 : 0;  // Glacier.cpp:23-28
            const int noMeltFlag = hasNoMelt && noMelt > 0 ? 1 : 0;
            if (second) {
                if (target(cover->processes[0]) != b1 || target(cover->processes[1]) != b2 ||
                    infinite != st.glacier_infinite_storage || noMeltFlag != st.glacier_no_melt_when_snow_cover)
                    throw Unsupported("the glacier covers must share their options and their sub-basin stores");
            } else {
                b1 = target(cover->processes[0]);
                b2 = target(cover->processes[1]);
                st.glacier_infinite_storage = infinite;
                st.glacier_no_melt_when_snow_cover = noMeltFlag;
            }
            setMelt(cover->name, cover->processes[1], second ? HBK_P_ICE2_DDF : HBK_P_ICE_DDF,
                    second ? HBK_P_ICE2_TMELT : HBK_P_ICE_TMELT, second ? HBK_P_ICE2_MELT_B : HBK_P_ICE_MELT_B,
                    second ? HBK_P_ICE2_MELT_C : HBK_P_ICE_MELT_C);
            const string& n = cover->name;
            lab(n + ":water_content", "glacier_water");
            lab(n + ":ice_content", "glacier_ice");
            lab(n + ":" + cover->processes[0].name + ":output", "glacier_outflow");
            lab(n + ":" + cover->processes[1].name + ":output", "glacier_melt");
            known.insert(cover->name);
            if (noSnow) {
                if (st.glacier_no_melt_when_snow_cover)  // IceContainer.cpp:10-32 needs the related snowpack
                    throw Unsupported("No snowpack provided for the glacier melt limitation.");
                continue;
            }
            const SnowpackInfo glInfo = snowpackOf(*cover);
            const Brick& glsp = *glInfo.brick;
            if ((glInfo.sublim == nullptr) != (gInfo.sublim == nullptr))
                throw Unsupported("sublimation must be on every snowpack or on none");
            if (glInfo.sublim) {
                setSublimation(glsp, *glInfo.sublim, second ? HBK_P_GLACIER2_SUBLIMATION : HBK_P_GLACIER_SUBLIMATION,
                               "glacier_snowpack_sublimation");
                if (second) twinLabels.insert(glsp.name + ":" + glInfo.sublim->name + ":output");
            }
            setMelt(glsp.name, glsp.processes[0], second ? HBK_P_GLACIER2_SNOW_DDF : HBK_P_GLACIER_SNOW_DDF,
                    second ? HBK_P_GLACIER2_SNOW_TMELT : HBK_P_GLACIER_SNOW_TMELT,
                    second ? HBK_P_GLACIER2_SNOW_MELT_B : HBK_P_GLACIER_SNOW_MELT_B,
                    second ? HBK_P_GLACIER2_SNOW_MELT_C : HBK_P_GLACIER_SNOW_MELT_C);
            lab(glsp.name + ":water_content", "glacier_snowpack_water");
            lab(glsp.name + ":snow_content", "glacier_snowpack_snow");
            lab(glsp.name + ":" + glsp.processes[0].name + ":output", "glacier_snowpack_melt");
            snowpacks.push_back(glsp.name);
            known.insert(glsp.name);
            if (glInfo.slide) slideOn.push_back({glsp.name, glInfo.slide});
            setRetention(glInfo, HBK_P_GLACIER_SNOW_WHC, HBK_P_GLACIER_SNOW_CFR, "glacier_snowpack");
            const int kind = glInfo.transfo ? snowIceKindOf(glInfo.transfo->kind) : HBK_SNOW_ICE_NONE;
            if (second && kind != st.snow_ice_kind)
                throw Unsupported("the snow -> ice transformation must be the same on every glacier snowpack");
            if (glInfo.transfo) {
                const Process& tp = *glInfo.transfo;
                st.snow_ice_kind = kind;
                bool has;
                string pname;
                if (st.snow_ice_kind == HBK_SNOW_ICE_CONSTANT) {
                    paramOf(tp.parameters, "snow_ice_transformation_rate", &has);
                    pname = has ? "snow_ice_transformation_rate" : "rate";
                } else {
                    paramOf(tp.parameters, "snow_ice_transformation_basal_acc_coeff", &has);
                    pname = has ? "snow_ice_transformation_basal_acc_coeff" : "basal_acc_coeff";
                    bool hasNh;
                    const float nh = paramOf(tp.parameters, "north_hemisphere", &hasNh);
                    st.north_hemisphere = (hasNh && nh == 0) ? 0 : 1;
                }
                set(second ? HBK_P_SNOW_ICE2 : HBK_P_SNOW_ICE, glsp.name, pname, paramOf(tp.parameters, pname));
                lab(glsp.name + ":" + tp.name + ":output", "glacier_snowpack_transfo");
            }
        }
        if (glacier) {
            known.insert(glacier->name);
            int slot[2] = {HBK_P_K_SNOW, HBK_P_K_ICE};
            const string* names[2] = {&b1, &b2};
            for (int i = 0; i < 2; ++i) {
                const Brick* bb = findBasinBrick(*names[i]);
                if (!bb || bb->kind != "storage" || kinds(*bb) != vector<string>{"outflow:linear"} ||
                    target(bb->processes[0]) != "outlet")
                    throw Unsupported("sub-basin store " + *names[i] + ": expected storage with outflow:linear -> outlet");
                set(slot[i], *names[i], "response_factor", paramOf(bb->processes[0].parameters, "response_factor"));
            }
        }
        {
            vector<string> a = snowTargets, b = snowpacks;
            std::sort(a.begin(), a.end());
            std::sort(b.begin(), b.end());
            if (a != b) throw Unsupported("snow must go to the snowpack of every land cover");
        }
        // liquid water in the snowpacks: the same processes on every snowpack (model_settings.py:248-255)
        st.snow_retention = 0;
        if (nHolding > 0) {
            if (nHolding != (int)snowpacks.size() || (nRefreeze != 0 && nRefreeze != nHolding))
                throw Unsupported("outflow:snow_holding / refreeze:degree_day must be on every snowpack or on none");
            st.snow_retention = HBK_RETENTION_HOLDING | (nRefreeze ? HBK_RETENTION_REFREEZE : 0) |
                                (rainToSnowpack ? HBK_RETENTION_RAIN_TO_SNOWPACK : 0);
            if (nRefreeze && st.melt_kind != HBK_MELT_DEGREE_DAY)  // ProcessRefreezeDegreeDay::IsValid (:19-37)
                throw Unsupported("The refreeze:degree_day process requires a melt:degree_day process on the same brick.");
            if (glacier2 || !slideOn.empty())
                throw Unsupported("liquid water in the snowpacks runs neither with a second glacier cover nor with "
                                  "transport:snow_slide");
        } else if (rainToSnowpack) {
            throw Unsupported("rain routed to the snowpacks needs a snow water retention process");
        }
        st.snow_slide_kind = HBK_SNOW_SLIDE_NONE;
        if (!slideOn.empty()) {
            // the snowpacks of the generic cover always slide, those of the glacier unless skipGlaciers; one set of
            // parameters for every sliding snowpack (ProcessLateralSnowSlide.cpp:24-31)
            if (!gInfo.slide) throw Unsupported("transport:snow_slide must be on the snowpack of the generic land cover");
            st.snow_slide_kind = (slideOn.size() == snowpacks.size() && snowpacks.size() > 1) ? HBK_SNOW_SLIDE_ALL
                                                                                             : HBK_SNOW_SLIDE_SKIP_GLACIERS;
            static const std::pair<const char*, int> names[6] = {
                {"coeff", HBK_P_SLIDE_COEFF}, {"exp", HBK_P_SLIDE_EXP}, {"min_slope", HBK_P_SLIDE_MIN_SLOPE},
                {"max_slope", HBK_P_SLIDE_MAX_SLOPE}, {"min_snow_holding_depth", HBK_P_SLIDE_MIN_HOLDING_DEPTH},
                {"max_snow_depth", HBK_P_SLIDE_MAX_SNOW_DEPTH}};
            for (auto& sp : slideOn) {
                for (auto& nm : names) {
                    const float v = paramOf(sp.second->parameters, nm.first);
                    if (sp.second != slideOn[0].second && v != paramOf(slideOn[0].second->parameters, nm.first))
                        throw Unsupported("transport:snow_slide: different parameters on the snowpacks");
                    set(nm.second, sp.first, nm.first, v);
                }
            }
        }
        for (auto& bb : base->subBasinBricks) {
            if (bb.name == b1 || bb.name == b2) continue;
            if (bb.kind != "storage" || kinds(bb) != vector<string>{"outflow:linear"} || target(bb.processes[0]) != "outlet")
                throw Unsupported("sub-basin brick " + bb.name);
        }

        // slow reservoir(s)
        const Brick* slow = findBrick(slowName);
        if (!slow || slow->kind != "storage" || slow->computedDirectly)
            throw Unsupported("infiltration target must be a storage solved by the ODE solver");
        vector<string> k = kinds(*slow);
        if (k == vector<string>{"et:socont", "outflow:linear", "overflow"}) {
            st.n_soil_stores = 1;
            st.slow_process_order = 0;
        } else if (k == vector<string>{"et:socont", "outflow:linear", "overflow", "percolation:constant"}) {
            st.n_soil_stores = 2;
            st.slow_process_order = 0;
        } else if (k == vector<string>{"et:socont", "outflow:linear", "percolation:constant", "overflow"}) {
            st.n_soil_stores = 2;
            st.slow_process_order = 1;
        } else {
            throw Unsupported("slow reservoir processes are not et:socont, outflow:linear, overflow[, percolation:constant]");
        }
        set(HBK_P_CAPACITY, slowName, "capacity", paramOf(slow->parameters, "capacity"));
        label(slowName + ":water_content", "slow_water");
        string slow2Name;
        for (size_t i = 0; i < slow->processes.size(); ++i) {
            const Process& p = slow->processes[i];
            static const std::map<string, string> lab = {{"et:socont", "slow_et"}, {"outflow:linear", "slow_outflow"},
                                                         {"overflow", "slow_overflow"},
                                                         {"percolation:constant", "slow_percolation"}};
            label(slowName + ":" + p.name + ":output", lab.at(k[i]));
            if (k[i] == "outflow:linear") {
                if (target(p) != "outlet") throw Unsupported("slow reservoir outflow must go to the outlet");
                set(HBK_P_K_SLOW_1, slowName, "response_factor", paramOf(p.parameters, "response_factor"));
            } else if (k[i] == "overflow") {
                if (target(p) != "outlet") throw Unsupported("slow reservoir overflow must go to the outlet");
            } else if (k[i] == "percolation:constant") {
                slow2Name = target(p);
                bool has;
                float v = paramOf(p.parameters, "percolation_rate", &has);
                if (!has) v = paramOf(p.parameters, "rate");
                set(HBK_P_PERCOLATION, slowName, "percolation_rate", v);
            }
        }
        if (st.n_soil_stores == 2) {
            const Brick* s2 = findBrick(slow2Name);
            bool hasCap = false;
            if (s2) paramOf(s2->parameters, "capacity", &hasCap);
            if (!s2 || s2->kind != "storage" || kinds(*s2) != vector<string>{"outflow:linear"} ||
                target(s2->processes[0]) != "outlet" || hasCap)
                throw Unsupported("second soil store: expected storage with outflow:linear -> outlet");
            set(HBK_P_K_SLOW_2, slow2Name, "response_factor", paramOf(s2->processes[0].parameters, "response_factor"));
            label(slow2Name + ":water_content", "slow2_water");
            label(slow2Name + ":" + s2->processes[0].name + ":output", "slow2_outflow");
            known.insert(slow2Name);
        }

        // quick store
        const Brick* quick = findBrick(quickName);
        bool hasCap = false;
        if (quick) paramOf(quick->parameters, "capacity", &hasCap);
        if (!quick || quick->kind != "storage" || quick->processes.size() != 1 || target(quick->processes[0]) != "outlet" || hasCap)
            throw Unsupported("surface runoff store: expected one process to the outlet");
        const Process& qp = quick->processes[0];
        if (qp.kind == "runoff:socont") {
            st.quick_kind = HBK_QUICK_SOCONT;
            set(HBK_P_QUICK, quickName, "beta", paramOf(qp.parameters, "beta"));
        } else if (qp.kind == "outflow:linear") {
            st.quick_kind = HBK_QUICK_LINEAR;
            set(HBK_P_QUICK, quickName, "response_factor", paramOf(qp.parameters, "response_factor"));
        } else {
            throw Unsupported("surface runoff process " + qp.kind);
        }
        label(quickName + ":water_content", "quick_water");
        label(quickName + ":" + qp.name + ":output", "quick_runoff");

        for (auto& b : full.hydroUnitBricks) {
            if (!known.count(b.name)) throw Unsupported("brick outside the Socont family: " + b.name);
            if (!b.forcing.empty()) throw Unsupported("brick-level forcing");
        }
        if (st.melt_kind < 0) st.melt_kind = HBK_MELT_DEGREE_DAY;
    }

    // ModelBuilder::AssignHydroUnitStructures (ModelBuilder.cpp:36-97) -> glacier-variant flag of one unit
    int glacierVariantOf(const vector<std::pair<string, double>>& covers) const {
        return structureIdOf(covers) == glacierStructureId ? 1 : 0;
    }
    // ... -> the id of the structure variant of one unit (results: hydro_units_structure_ids)
    int structureIdOf(const vector<std::pair<string, double>>& covers) const {
        if (variantCovers.size() <= 1) return variantCovers.begin()->first;
        std::set<string> present;
        for (auto& c : covers)
            if (std::fabs(c.second) > 1e-8) present.insert(c.first);
        int sid = -1;
        for (auto& v : variantCovers) {
            if (v.second == present) {
                sid = v.first;
                break;
            }
        }
        if (sid < 0) {
            size_t best = SIZE_MAX, largest = 0;
            int largestId = variantCovers.begin()->first;
            for (auto& v : variantCovers) {
                if (v.second.size() > largest) {
                    largest = v.second.size();
                    largestId = v.first;
                }
                bool superset = std::includes(v.second.begin(), v.second.end(), present.begin(), present.end());
                if (superset && v.second.size() < best) {
                    best = v.second.size();
                    sid = v.first;
                }
            }
            if (sid < 0) sid = largestId;
        }
        return sid;
    }

    // SettingsModel::SetParameterValue target resolution incl. lists and "type:" components -> slots
    vector<int> resolveAll(const string& component, const string& name) const {
        std::set<int> slots;
        size_t start = 0;
        while (start <= component.size()) {
            size_t end = component.find(',', start);
            if (end == string::npos) end = component.size();
            string tok = component.substr(start, end - start);
            size_t a = tok.find_first_not_of(' '), b = tok.find_last_not_of(' ');
            if (a != string::npos) {
                tok = tok.substr(a, b - a + 1);
                if (tok.rfind("type:", 0) == 0) {
                    string kind = tok.substr(5);
                    for (auto& bk : brickKinds) {
                        bool m = bk.second == kind ||
                                 (kind == "land_cover" && (bk.second == "generic_land_cover" || bk.second == "glacier"));
                        auto it = paramIndex.find({bk.first, name});
                        if (m && it != paramIndex.end()) slots.insert(it->second);
                    }
                } else {
                    auto it = paramIndex.find({tok, name});
                    if (it != paramIndex.end()) slots.insert(it->second);
                }
            }
            start = end + 1;
        }
        return vector<int>(slots.begin(), slots.end());
    }
};

// ---------------------------------------------------------------------------------------------------------
// dates (UtilsDateTime): proleptic Gregorian <-> modified Julian date
inline long daysFromCivil(int y, unsigned m, unsigned d) {
    y -= m <= 2;
    const long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long)doe - 719468;  // days since 1970-01-01
}
inline void civilFromDays(long z, int& y, unsigned& m, unsigned& d) {
    z += 719468;
    const long era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    y = (int)yoe + (int)era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    d = doy - (153 * mp + 2) / 5 + 1;
    m = mp < 10 ? mp + 3 : mp - 9;
    y += m <= 2;
}
constexpr long kMjdEpochUnixDays = -40587;  // 1858-11-17
inline double parseDateMjd(const string& text) {
    int y, mo, d;
    if (std::sscanf(text.c_str(), "%d-%d-%d", &y, &mo, &d) != 3) throw std::invalid_argument("bad date '" + text + "'");
    return (double)(daysFromCivil(y, (unsigned)mo, (unsigned)d) - kMjdEpochUnixDays);
}

// Action dates -> step indices (TimeMachine.cpp:49-59, ActionsManager.cpp:80-104, Action.cpp:18-57)
inline vector<int> recursiveSteps(double mjdStart, int nSteps, int month, int day, double dt) {
    static const int dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if (month < 1 || month > 12) month = 1;
    day = std::max(1, std::min(day, dim[month - 1]));
    vector<int> out;
    for (int k = 1; k < nSteps; ++k) {
        int y;
        unsigned m, d;
        civilFromDays((long)std::floor(mjdStart + k * dt) + kMjdEpochUnixDays, y, m, d);
        if ((int)m == month && (int)d == day) out.push_back(k);
    }
    return out;
}
inline int sporadicStep(double mjdStart, double date, double dt) {
    return std::max(1, (int)std::ceil((date - mjdStart) / dt - 1e-12));
}
inline double checkFraction(double f) {  // Action::CheckLandCoverAreaFraction (Action.cpp:71-91)
    if (std::fabs(f) <= 1e-8) return 0.0;
    if (std::fabs(1 - f) <= 1e-8) return 1.0;
    if (f < 0 || f > 1) throw std::invalid_argument("The fraction is not in the range [0 .. 1]");
    return f;
}

// ---------------------------------------------------------------------------------------------------------
// Standalone value classes of the binding surface (Binding.cpp:270-306). The reference exposes them but its Python
// layer never attaches a modifier to a running model (nothing adds parameters to ModelHydro's ParametersUpdater), so
// here too they are plain objects: time-varying parameters inside a run stay out of scope.
enum class ParameterModifierType { Yearly, Monthly, Dates };

// base/ParameterModifier.cpp: value looked up by the year (Yearly), the month (Monthly) or the exact date (Dates)
// of a modified Julian date; NaN when the date has no entry.
struct ParameterModifier {
    ParameterModifierType type = ParameterModifierType::Yearly;
    vector<double> dates;
    vector<float> values;
    ParameterModifier() = default;
    explicit ParameterModifier(ParameterModifierType t) : type(t) {}
    bool SetYearlyValues(int yearStart, int yearEnd, const vector<float>& v) {
        if (type != ParameterModifierType::Yearly) throw std::runtime_error("ParameterModifier::SetYearlyValues - Modifier type is not Yearly.");
        values = v;
        dates.clear();
        for (int y = yearStart; y <= yearEnd; ++y) dates.push_back((double)y);
        return values.size() == dates.size();
    }
    bool SetMonthlyValues(const vector<float>& v) {
        if (type != ParameterModifierType::Monthly) throw std::runtime_error("ParameterModifier::SetMonthlyValues - Modifier type is not Monthly.");
        values = v;
        dates.clear();
        return values.size() == 12;
    }
    bool SetDatesAndValues(const vector<double>& d, const vector<float>& v) {
        if (type != ParameterModifierType::Dates) throw std::runtime_error("ParameterModifier::SetDatesAndValues - Modifier type is not Dates.");
        dates = d;
        values = v;
        return values.size() == dates.size();
    }
    // Utils.cpp FindT with tolerance 0: index of the entry equal to `value`, -1 otherwise
    int find(double value) const {
        for (size_t i = 0; i < dates.size() && i < values.size(); ++i)
            if (dates[i] == value) return (int)i;
        return -1;
    }
    float UpdateValue(double mjd) const {
        const float nan = std::numeric_limits<float>::quiet_NaN();
        int y;
        unsigned m, d;
        civilFromDays((long)std::floor(mjd) + kMjdEpochUnixDays, y, m, d);
        switch (type) {
            case ParameterModifierType::Yearly: {
                const int i = find((double)y);
                return i < 0 ? nan : values[i];
            }
            case ParameterModifierType::Monthly: return values.size() == 12 ? values[m - 1] : nan;
            default: {
                const int i = find(mjd);
                return i < 0 ? nan : values[i];
            }
        }
    }
    bool UpdatesOnYearChange() const { return type == ParameterModifierType::Yearly; }
    bool UpdatesOnMonthChange() const { return type == ParameterModifierType::Monthly; }
    bool UpdatesOnDateChange() const { return type == ParameterModifierType::Dates; }
};

// base/Parameter.h / Parameter.cpp: name + float32 value + optional modifier.
struct Parameter {
    string name;
    float value;
    ParameterModifier modifier;
    bool hasModifier = false;
    Parameter(const string& n, float v) : name(n), value(v) {}
    void SetModifier(const ParameterModifier& m) {
        modifier = m;
        hasModifier = true;
    }
    bool UpdateFromModifier(double mjd) {
        if (!hasModifier) return false;
        const float v = modifier.UpdateValue(mjd);
        if (std::isnan(v)) return false;
        value = v;
        return true;
    }
};

// base/TimeSeries.cpp (TimeSeries::Create): a named [time x ids] block handed to ModelHydro::AddTimeSeries.
struct TimeSeries {
    string name;
    vector<double> time;
    vector<int> ids;
    vector<double> data;  // [time][ids]
};

// ---------------------------------------------------------------------------------------------------------
struct Action {
    virtual ~Action() = default;
};
struct ActionLandCoverChange : Action {
    struct Change {
        double date;
        int unitId;
        string cover;
        double area;
    };
    vector<Change> changes;
    void AddChange(double date, int unitId, const string& cover, double area) { changes.push_back({date, unitId, cover, area}); }
    int GetChangeCount() const { return (int)changes.size(); }
    int GetLandCoverCount() const {
        std::set<string> s;
        for (auto& c : changes) s.insert(c.cover);
        return (int)s.size();
    }
};
struct ActionGlacierEvolutionDeltaH : Action {
    int month = 10;
    string cover;
    vector<int> unitIds;
    int rows = 0;
    vector<double> areas, volumes;  // [rows][units]
    void AddLookupTables(int m, const string& c, const vector<int>& ids, int nRows, const vector<double>& a,
                         const vector<double>& v) {
        month = m;
        cover = c;
        unitIds = ids;
        rows = nRows;
        areas = a;
        volumes = v;
    }
};
struct ActionGlacierEvolutionAreaScaling : ActionGlacierEvolutionDeltaH {};
struct ActionGlacierSnowToIceTransformation : Action {
    int month, day;
    string cover;
    ActionGlacierSnowToIceTransformation(int m, int d, const string& c) : month(m), day(d), cover(c) {}
};

inline float slopeMPerM(double value, const string& unit) {
    // HydroUnitProperty::GetValue(..., "m/m") + float cast of HydroUnit::GetPropertyFloat
    // (HydroUnitProperty.cpp:19-49, HydroUnit.cpp:146-148)
    if (unit == "m/m" || unit == "m_per_m") return (float)value;
    if (unit == "degrees" || unit == "degree" || unit == "deg" || unit == "°")
        return (float)std::tan(value * 3.1415926535897932384626433832795 / 180.0);
    throw std::invalid_argument("unsupported slope unit '" + unit + "'");
}

class ModelHydro {
  public:
    ModelHydro() = default;
    ~ModelHydro() { hbk_engine_destroy(_e); }
    ModelHydro(const ModelHydro&) = delete;
    ModelHydro& operator=(const ModelHydro&) = delete;

    void InitializeWithBasin(SettingsModel& ms, SettingsBasin& bs, int device = 0) {
        if (ms.solver.empty()) throw std::invalid_argument("Model settings are not valid.");
        if (ms.timeStepUnit != "day" && ms.timeStepUnit != "days" && ms.timeStepUnit != "d")
            throw std::invalid_argument("Model initialization failed: only daily time steps (time_step_unit='day')");
        _mjd0 = parseDateMjd(ms.timerStart);
        _dt = (double)ms.timeStep;
        _nSteps = (int)(1 + (parseDateMjd(ms.timerEnd) - _mjd0) / _dt);  // TimeMachine.cpp:61-64
        try {
            _c.reset(new Compiled(ms, _dt, _mjd0));
        } catch (const Unsupported& err) {
            throw std::invalid_argument(string("Model initialization failed: ") + err.what());
        }
        _settings = &ms;
        _units = bs.units;
        if (_units.empty()) throw std::invalid_argument("Basin initialization failed: no hydro unit");
        hbk_engine_destroy(_e);
        _e = nullptr;
        // Engine units: the basin's units in order, then one TWIN per unit that carries the glacier bricks when the
        // structure has a second glacier cover (hbk_set_unit_twins): same area, elevation, slope and forcing, no ground,
        // the second cover's fraction as its glacier fraction.
        const int nBasin = (int)_units.size();
        _twinOf.assign(nBasin, -1);
        _twinCol.assign(nBasin, -1);
        if (!_c->glacier2Name.empty()) {
            for (int u = 0; u < nBasin; ++u) {
                if (_c->glacierVariantOf(_units[u].landCovers)) {
                    _twinCol[u] = (int)_twinOf.size();
                    _twinOf.push_back(u);
                }
            }
        }
        const int U = (int)_twinOf.size();
        if (hbk_engine_create(&_c->st, U, _nSteps, device, &_e) != HBK_OK)
            throw std::invalid_argument("hbk_engine_create failed: no usable CUDA device (the B200 engine has no CPU fallback)");
        vector<double> area(U), elev(U), fg(U, 1.0), fgl(U, 0.0);
        vector<float> slope(U, 0.0f);
        vector<int32_t> gv(U, 0);
        for (int e = 0; e < U; ++e) {
            const bool twin = e >= nBasin;
            const int u = twin ? _twinOf[e] : e;
            area[e] = _units[u].area;
            elev[e] = _units[u].elevation;
            if (_c->st.quick_kind == HBK_QUICK_SOCONT) {
                if (!_units[u].hasSlope) throw std::invalid_argument("No property with the name 'slope' was found.");
                slope[e] = slopeMPerM(_units[u].slopeValue, _units[u].slopeUnit);
            }
            gv[e] = twin ? 1 : _c->glacierVariantOf(_units[u].landCovers);
            if (twin) fg[e] = 0.0;
            for (auto& lc : _units[u].landCovers) {
                if (!twin && lc.first == _c->groundName) fg[e] = lc.second;
                if (!twin && !_c->glacierName.empty() && lc.first == _c->glacierName) fgl[e] = lc.second;
                if (twin && lc.first == _c->glacier2Name) fgl[e] = lc.second;
            }
        }
        if (U > nBasin) check(hbk_set_unit_twins(_e, _twinOf.data()));
        check(hbk_set_units(_e, area.data(), elev.data(), _c->st.quick_kind == HBK_QUICK_SOCONT ? slope.data() : nullptr,
                            gv.data(), fg.data(), fgl.data()));
        if (_c->st.snow_slide_kind != HBK_SNOW_SLIDE_NONE) {
            // ModelBuilder.cpp:476-508: connections by unit id -> basin positions, in the order they were added
            std::map<int, int> pos;
            for (int u = 0; u < nBasin; ++u) pos[_units[u].id] = u;
            vector<int32_t> giver, receiver;
            vector<double> fraction;
            for (auto& c : bs.connections) {
                auto g = pos.find(c.giver), r = pos.find(c.receiver);
                if (g == pos.end() || r == pos.end())
                    throw std::invalid_argument("Model initialization failed: lateral connection to an unknown hydro unit");
                giver.push_back(g->second);
                receiver.push_back(r->second);
                fraction.push_back(c.fraction);
            }
            vector<float> slopeDeg(U, 0.0f);
            for (int e = 0; e < U; ++e) {
                const int u = unitOf(e);
                if (!_units[u].hasSlope) throw std::invalid_argument("No property with the name 'slope' was found.");
                const string& su = _units[u].slopeUnit;
                // HydroUnitProperty::GetValue (HydroUnitProperty.cpp:19-49) converts degrees to other units, never back
                if (su != "degrees" && su != "degree" && su != "deg" && su != "°")
                    throw std::invalid_argument("The unit 'degrees' is not supported for the property 'slope'.");
                slopeDeg[e] = (float)_units[u].slopeValue;
            }
            check(hbk_set_lateral_connections(_e, (int)giver.size(), giver.data(), receiver.data(), fraction.data(),
                                              slopeDeg.data()));
        }
        if (_c->st.melt_kind == HBK_MELT_DEGREE_DAY_ASPECT) {
            vector<int32_t> aspect(U);
            for (int e = 0; e < U; ++e) aspect[e] = Compiled::aspectOf(_units[unitOf(e)].aspectClass);
            check(hbk_set_unit_aspect(_e, aspect.data()));
        }
        check(hbk_set_spinup_steps(_e, (int)(ms.spinupDays / _dt)));
        _labels = ms.GetHydroUnitLogLabels();
        _sets.assign(_c->params, _c->params + HBK_N_PARAMS);
        _nSets = 1;
    }
    bool IsValid() const { return _e != nullptr; }

    // ---- forcing
    bool CreateTimeSeries(const string& name, const double* time, int nTime, const int* ids, int nIds, const double* data) {
        string key = name;
        std::transform(key.begin(), key.end(), key.begin(), ::tolower);
        int var;
        if (key == "precipitation" || key == "p") {
            var = HBK_VAR_PRECIPITATION;
        } else if (key == "temperature" || key == "t") {
            var = HBK_VAR_TEMPERATURE;
        } else if (key == "pet") {
            var = HBK_VAR_PET;
        } else if (key == "solar_radiation" || key == "r_solar" || key == "r") {  // TimeSeries.cpp:134-146
            var = HBK_VAR_SOLAR_RADIATION;
        } else {
            throw std::invalid_argument("Unknown forcing variable '" + name + "'");
        }
        Series s;
        s.time.assign(time, time + nTime);
        s.ids.assign(ids, ids + nIds);
        s.data.assign(data, data + (size_t)nTime * nIds);
        _series[var] = std::move(s);
        return true;
    }
    bool AddTimeSeries(const TimeSeries& ts) {  // ModelHydro::AddTimeSeries (ModelHydro.cpp:296-304)
        return CreateTimeSeries(ts.name, ts.time.data(), (int)ts.time.size(), ts.ids.data(), (int)ts.ids.size(), ts.data.data());
    }
    void ClearTimeSeries() {
        _series.clear();
        _forcingLoaded = false;
        _stationVars.clear();
    }
    bool AttachTimeSeriesToHydroUnits() {
        const int U = EngineUnitCount();  // a twin reads its unit's column
        for (auto& kv : _series) {
            const Series& s = kv.second;
            std::map<int, int> pos;
            for (int i = 0; i < (int)s.ids.size(); ++i) pos[s.ids[i]] = i;
            vector<int> cols(U);
            for (int u = 0; u < U; ++u) {
                auto it = pos.find(_units[unitOf(u)].id);
                if (it == pos.end()) return false;
                cols[u] = it->second;
            }
            const long start = std::lround(_mjd0 - s.time[0]);
            if (start < 0 || start + _nSteps > (long)s.time.size())
                throw std::invalid_argument("The forcing data do not cover the modelling period.");
            vector<double> buf((size_t)_nSteps * U);
            const size_t n = s.ids.size();
            for (int t = 0; t < _nSteps; ++t)
                for (int u = 0; u < U; ++u) buf[(size_t)t * U + u] = s.data[(size_t)(start + t) * n + cols[u]];
            check(hbk_set_forcing(_e, kv.first, buf.data()));
        }
        _forcingLoaded = !_series.empty();
        return true;
    }
    bool ForcingLoaded() const { return _forcingLoaded; }
    void SetStationForcing(int var, const double* series, int n, double zref, double gradient, const double* gradSet,
                           double corr, const double* corrSet) {
        if (n != _nSteps) throw std::invalid_argument("station series length differs from the modelling period");
        if (gradSet || corrSet) check(hbk_set_parameters(_e, _nSets, _sets.data()));
        check(hbk_set_station_forcing(_e, var, series, zref, gradient, gradSet, corr, corrSet));
        _stationVars.insert(var);
        _forcingLoaded = _stationVars.size() >= (_c->st.melt_kind == HBK_MELT_TEMPERATURE_INDEX ? 4u : 3u);
    }

    // ---- parameters
    void UpdateParameters(SettingsModel& ms) {
        Compiled c(ms, _dt);
        std::memcpy(_c->params, c.params, sizeof(c.params));
        _sets.assign(c.params, c.params + HBK_N_PARAMS);
        _nSets = 1;
    }
    void SetParameterSets(const float* values, int nSets) {
        _sets.assign(values, values + (size_t)nSets * HBK_N_PARAMS);
        _nSets = nSets;
    }
    vector<int> ParameterSlot(const string& component, const string& name) const { return _c->resolveAll(component, name); }
    vector<float> BaseParameters() const { return vector<float>(_c->params, _c->params + HBK_N_PARAMS); }

    // ---- actions
    bool AddAction(Action* a) {
        _actions.push_back(a);
        return true;
    }
    int GetActionCount() const { return (int)_actions.size(); }
    int GetSporadicActionItemCount() const {
        int n = 0;
        for (auto* a : _actions)
            if (auto* lc = dynamic_cast<ActionLandCoverChange*>(a)) n += lc->GetChangeCount();
        return n;
    }

    // ---- run
    void Reset() {}
    void SaveAsInitialState() { check(hbk_save_as_initial_state(_e)); }
    void Run(bool record = true) {
        if (!_forcingLoaded) throw std::invalid_argument("Time step processing failed: no forcing attached.");
        scheduleActions();
        check(hbk_set_parameters(_e, _nSets, _sets.data()));
        uint64_t mask = 0;
        if (record) {
            for (auto& l : _labels) {
                auto it = _c->labels.find(l);
                if (it != _c->labels.end()) mask |= 1ull << it->second;
            }
            if (!_labels.empty() && (_settings->LogAll() || _settings->RecordsFractions())) {
                mask |= 1ull << HBK_R_FRACTION_GROUND;
                if (!_c->glacierName.empty()) mask |= 1ull << HBK_R_FRACTION_GLACIER;
            }
        }
        check(hbk_set_record_mask(_e, mask));
        check(hbk_run(_e));
    }

    // ---- results
    int StepCount() const { return _nSteps; }
    int UnitCount() const { return (int)_units.size(); }
    // engine units = the basin's units + the twins that carry a second glacier cover (hbk_set_unit_twins)
    int EngineUnitCount() const { return _twinOf.empty() ? (int)_units.size() : (int)_twinOf.size(); }
    int unitOf(int engineUnit) const {
        return (engineUnit < (int)_units.size() || _twinOf.empty()) ? engineUnit : _twinOf[engineUnit];
    }
    vector<int> TwinOf() const { return _twinOf; }
    // a recorded series [T x engine units] -> [T x basin units]: the units' own columns, or the twins' columns for a label
    // of the second glacier cover (NaN where a unit has no twin)
    void gatherRecorded(int slot, int set, bool twinLabel, double* out) {
        const int U = UnitCount(), UE = EngineUnitCount();
        if (UE == U) {
            check(hbk_get_recorded(_e, slot, set, out));
            return;
        }
        vector<double> buf((size_t)_nSteps * UE);
        check(hbk_get_recorded(_e, slot, set, buf.data()));
        for (int t = 0; t < _nSteps; ++t)
            for (int u = 0; u < U; ++u) {
                const int col = twinLabel ? _twinCol[u] : u;
                out[(size_t)t * U + u] = col < 0 ? std::nan("") : buf[(size_t)t * UE + col];
            }
    }
    int SetCount() const { return _nSets; }
    void GetOutlet(int first, int n, double* out) { check(hbk_get_outlet(_e, first, n, out)); }
    vector<string> GetRecordedHydroUnitLabels() const { return _labels; }
    vector<string> GetRecordedHydroUnitFractionLabels() const {
        if (_settings && (_settings->LogAll() || _settings->RecordsFractions())) return _settings->GetLandCoverBricksNames();
        return {};
    }
    void GetHydroUnitValues(const string& label, int set, double* out) {
        auto it = _c->labels.find(label);
        if (it == _c->labels.end()) throw std::runtime_error("Label '" + label + "' is not recorded.");
        gatherRecorded(it->second, set, _c->twinLabels.count(label) > 0, out);
    }
    void GetHydroUnitFractions(const string& label, double* out) {
        int slot;
        if (label == _c->groundName) {
            slot = HBK_R_FRACTION_GROUND;
        } else if (!_c->glacierName.empty() && label == _c->glacierName) {
            slot = HBK_R_FRACTION_GLACIER;
        } else if (!_c->glacier2Name.empty() && label == _c->glacier2Name) {
            slot = HBK_R_FRACTION_GLACIER;  // of the twin units
        } else {
            throw std::runtime_error("Fraction label '" + label + "' is not recorded.");
        }
        gatherRecorded(slot, 0, !_c->glacier2Name.empty() && label == _c->glacier2Name, out);
    }
    vector<int> GetHydroUnitIds() const {
        vector<int> ids;
        for (auto& u : _units) ids.push_back(u.id);
        return ids;
    }
    vector<int> GetHydroUnitStructureIds() const {
        vector<int> ids;
        for (auto& u : _units) ids.push_back(_c->structureIdOf(u.landCovers));
        return ids;
    }
    double StartMjd() const { return _mjd0; }
    double TimeStepDays() const { return _dt; }
    vector<double> GetHydroUnitAreas() const {
        vector<double> a;
        for (auto& u : _units) a.push_back(u.area);
        return a;
    }
    void GetState(int slot, double* out) { check(hbk_get_state(_e, slot, out)); }
    void GetBasinState(double* out) { check(hbk_get_basin_state(_e, out)); }
    void GetRecorded(int slot, int set, double* out) { check(hbk_get_recorded(_e, slot, set, out)); }
    const Compiled& compiled() const { return *_c; }
    const vector<UnitSettings>& units() const { return _units; }
    hbk_engine* engine() { return _e; }
    // initial land-cover fractions and basin weights (area / basin area) per ENGINE unit: a twin has no ground, the
    // second glacier cover's fraction, and its unit's weight
    void fractions(vector<double>& g, vector<double>& gl) const {
        const int UE = EngineUnitCount(), U = UnitCount();
        g.assign(UE, 1.0);
        gl.assign(UE, 0.0);
        for (int e = 0; e < UE; ++e) {
            const bool twin = e >= U;
            if (twin) g[e] = 0.0;
            for (auto& lc : _units[unitOf(e)].landCovers) {
                if (!twin && lc.first == _c->groundName) g[e] = lc.second;
                if (!twin && !_c->glacierName.empty() && lc.first == _c->glacierName) gl[e] = lc.second;
                if (twin && lc.first == _c->glacier2Name) gl[e] = lc.second;
            }
        }
    }
    vector<double> EngineUnitWeights() const {
        double total = 0;
        for (auto& u : _units) total += u.area;
        vector<double> w(EngineUnitCount());
        for (int e = 0; e < (int)w.size(); ++e) w[e] = _units[unitOf(e)].area / total;
        return w;
    }

  private:
    struct Series {
        vector<double> time, data;
        vector<int> ids;
    };
    hbk_engine* _e = nullptr;
    std::unique_ptr<Compiled> _c;
    SettingsModel* _settings = nullptr;  // non-owning, like Process parameter pointers into the SettingsModel
    vector<UnitSettings> _units;
    vector<int> _twinOf, _twinCol;  // engine unit -> basin unit (-1 for the units themselves); basin unit -> twin column
    std::map<int, Series> _series;
    std::set<int> _stationVars;
    bool _forcingLoaded = false;
    vector<Action*> _actions;  // non-owning (ActionsManager.cpp:24-39)
    vector<string> _labels;
    vector<float> _sets;
    int _nSets = 1, _nSteps = 0;
    double _mjd0 = 0, _dt = 1;

    void check(int rc) {
        if (rc != HBK_OK) throw std::invalid_argument(string("hbk error ") + std::to_string(rc) + ": " + hbk_last_error(_e));
    }

    void scheduleActions() {
        check(hbk_clear_actions(_e));
        if (_actions.empty()) return;
        std::map<int, int> pos;
        for (int i = 0; i < (int)_units.size(); ++i) pos[_units[i].id] = i;
        struct Ev {
            int step, order, kind, unit;
            double value;
        };
        vector<Ev> events;
        struct Sp {
            double date;
            int order, unitId;
            string cover;
            double area;
        };
        vector<Sp> sporadic;
        bool haveTables = false;
        // 0 for the first glacier cover, 1 for the second (carried by the twin units); engine unit of (unit, cover)
        auto coverIndex = [&](const string& name, const char* what) -> int {
            if (!_c->glacierName.empty() && name == _c->glacierName) return 0;
            if (!_c->glacier2Name.empty() && name == _c->glacier2Name) return 1;
            throw std::invalid_argument(string(what) + ": land cover '" + name + "' is not a glacier cover");
        };
        auto engineUnit = [&](int u, int cover) -> int {
            if (cover == 0) return u;
            if (_twinCol[u] < 0)
                throw std::invalid_argument("The land cover '" + _c->glacier2Name + "' was not found in hydro unit " +
                                            std::to_string(_units[u].id));
            return _twinCol[u];
        };
        for (int order = 0; order < (int)_actions.size(); ++order) {
            Action* a = _actions[order];
            if (auto* s2i = dynamic_cast<ActionGlacierSnowToIceTransformation*>(a)) {
                const int c = coverIndex(s2i->cover, "snow-to-ice");
                for (int k : recursiveSteps(_mjd0, _nSteps, s2i->month, s2i->day, _dt)) events.push_back({k, order, 1, c, 0});
            } else if (auto* dh = dynamic_cast<ActionGlacierEvolutionDeltaH*>(a)) {
                // ActionGlacierEvolutionAreaScaling shares the lookup-table interface and Init of the delta-h action
                // (…AreaScaling.cpp:9-68); only its Apply rule differs (event kind 3)
                const bool areaScaling = dynamic_cast<ActionGlacierEvolutionAreaScaling*>(a) != nullptr;
                if (haveTables) throw std::invalid_argument("only one glacier-evolution action per model");
                haveTables = true;
                const int c = coverIndex(dh->cover, "delta-h");
                vector<int32_t> units;
                for (int id : dh->unitIds) {
                    auto it = pos.find(id);
                    if (it == pos.end()) throw std::invalid_argument("The hydro unit " + std::to_string(id) + " was not found");
                    units.push_back(engineUnit(it->second, c));
                }
                check(hbk_set_delta_h_tables(_e, (int)units.size(), units.data(), dh->rows, dh->areas.data(), dh->volumes.data()));
                for (int k : recursiveSteps(_mjd0, _nSteps, dh->month, 1, _dt))
                    events.push_back({k, order, areaScaling ? 3 : 2, -1, 0});
            } else if (auto* lc = dynamic_cast<ActionLandCoverChange*>(a)) {
                for (auto& c : lc->changes) sporadic.push_back({c.date, order, c.unitId, c.cover, c.area});
            } else {
                throw std::invalid_argument("This action is not available in the B200 engine");
            }
        }
        std::stable_sort(sporadic.begin(), sporadic.end(), [](const Sp& a, const Sp& b) { return a.date < b.date; });
        int n = 0;
        for (auto& s : sporadic) {
            auto it = pos.find(s.unitId);
            if (it == pos.end()) throw std::invalid_argument("The hydro unit " + std::to_string(s.unitId) + " was not found");
            if (s.cover == _c->groundName) continue;  // undone by FixLandCoverFractionsTotal (HydroUnit.cpp:459-490)
            if (s.cover != _c->glacierName && (_c->glacier2Name.empty() || s.cover != _c->glacier2Name))
                throw std::invalid_argument("Land cover '" + s.cover + "' was not found.");
            double fraction = checkFraction(s.area / _units[it->second].area);
            events.push_back({sporadicStep(_mjd0, s.date, _dt), 10000 + n++, 0,
                              engineUnit(it->second, coverIndex(s.cover, "land cover change")), fraction});
        }
        std::stable_sort(events.begin(), events.end(), [](const Ev& a, const Ev& b) {
            return a.step != b.step ? a.step < b.step : a.order < b.order;
        });
        for (auto& ev : events) {
            if (ev.kind == 0) {
                check(hbk_add_land_cover_change(_e, ev.step, ev.unit, ev.value));
            } else if (ev.kind == 1) {
                check(hbk_add_snow_to_ice_on(_e, ev.step, ev.unit));
            } else if (ev.kind == 3) {
                check(hbk_add_area_scaling_update(_e, ev.step));
            } else {
                check(hbk_add_delta_h_update(_e, ev.step));
            }
        }
    }
};

}  // namespace hb
