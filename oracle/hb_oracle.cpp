// oracle/hb_oracle.cpp — CPU restatement of hydrobricks' per-time-step solve (ModelHydro::Run()).
//
// TEST INFRASTRUCTURE ONLY (see hb_oracle.h). Plain C++, flat arrays, no virtual dispatch. Every
// function cites the reference file:line (under /root/reference/core/src) it restates. The
// arithmetic keeps the reference's operation order and its float32-parameter promotions, so that on
// the same compiler flags it reproduces the compiled reference (oracle/_ref) bit for bit.
//
// Parity status: PINNED (tests/test_oracle_golden.py: SolverTest / ModelSocontTest known-answer vectors and
// tests/golden/*.npz, vectors generated from the reference by tools/make_goldens.py, bit-exact;
// tests/test_oracle_vs_ref.py: bit-exact against oracle/_ref where that library and /root/reference are present).
#include "hb_oracle.h"

#ifdef HBO_COUNT_FLOPS  // instrumented build (oracle/flop_count.h): every FP64 operation counts itself
hbo_counters g_hbo_counters = {0, 0, 0, 0, 0, 0, 0};
extern "C" void hbo_flop_counts(unsigned long long* out7, int reset) {
    const hbo_counters c = g_hbo_counters;
    out7[0] = c.add; out7[1] = c.mul; out7[2] = c.div; out7[3] = c.cmp; out7[4] = c.abs; out7[5] = c.pow; out7[6] = c.exp;
    if (reset) g_hbo_counters = {0, 0, 0, 0, 0, 0, 0};
}
#endif

#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

namespace {

// base/NumericConstants.h:21-34
const double EPSILON_D = std::numeric_limits<double>::epsilon();
const double EPSILON_F = std::numeric_limits<float>::epsilon();
const double PRECISION = 0.00000001;

inline bool NearlyEqual(double a, double b, double tol) { return std::fabs(a - b) <= tol; }
inline bool NearlyZero(double v, double tol) { return std::fabs(v) <= tol; }
inline bool GreaterThan(double a, double b, double tol) { return a > b + tol; }
inline bool LessThanOrEqual(double a, double b, double tol) { return a <= b + tol; }

struct Flux {  // fluxes/Flux.h:133-153
    int kind;
    double amount = 0;
    double* changeRate = nullptr;
    bool isStatic = false;
    bool needsWeighting = false;
    double fractionUnitArea = 1.0;
    double fractionLandCover = 1.0;
    double fractionTotal = 1.0;
    int targetContainer = -1;
    int forcingVar = -1;
    int unit = -1;
};

struct Container {  // containers/WaterContainer.h:342-353
    int brick;
    int contentKind;
    double content = 0, dyn = 0, stat = 0, initial = 0;
    bool hasCapacity = false;
    float capacity = 0;
    bool infinite = false;
    bool noMeltWhenSnowCover = false;
    bool allowNegative = false;  // WaterContainer::SetAllowNegativeContent (routing:gr6j's bottomless store)
    int relatedSnowpackBrick = -1;
    int overflowProcess = -1;
    std::vector<int> inputs;
};

struct Process {
    int brick, type, container;
    float params[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // lateral processes (ProcessLateral.cpp): per output the receiving snowpack brick and the connection's weight
    std::vector<int> lateralTargets;
    std::vector<double> weights;
    float slopeDeg = 0;
    int targetBrick = -1;
    double unitArea = 0;
    float slope = 0;
    std::vector<int> outputs;
    std::vector<double> rates;
    // capillary:hbv (ProcessCapillaryHBV.h:66-69): target soil stores and, per target, the land covers feeding it
    std::vector<int> capTargets;
    std::vector<std::vector<int>> capWeights;
    // routing:hbv (ProcessRoutingHBV.h): unit-hydrograph ordinates and the delivery schedule
    std::vector<double> uhOrd, stuh;
    double lastMaxbas = 0, previousContent = 0, processStorage = 0;
    // routing:gr4j (ProcessRoutingGR4J.h): the two unit hydrographs, the routing store, the last two discharges
    std::vector<double> uh1Ord, uh2Ord, stuh1, stuh2;
    double gr_r = 0, gr_qr = 0, gr_qd = 0, gr_rexp = 0, gr_qrexp = 0;  // + routing:gr6j's exponential store
};

struct Brick {
    int kind, unit;
    bool needsSolver;
    int parent = -1;
    int genericCover = -1;  // the unit's generic (soil) land cover, for land covers (HydroUnit::GetGenericLandCover)
    int snowpack = -1;      // '<cover>_snowpack' brick of a land cover
    double areaFraction = 1.0;
    double initialAreaFraction = 1.0;
    int water = -1, second = -1;  // second: snow (snowpack) or ice (glacier)
    std::vector<int> processes;
};

struct Splitter {
    int unit, type;
    float params[4] = {0, 0, 1, 1};
    std::vector<int> inputs, outputs;
};

struct Record {
    int kind, id;
};

struct Event {
    int step, kind, brick;
    double value;
};

}  // namespace

struct hbo_model {
    int nUnits, nSteps, solver, spinupSteps;
    double dt;
    double startMjd = 44605.0;
    int cursor = 0;
    std::vector<Flux> fluxes;
    std::vector<Container> containers;
    std::vector<Process> processes;
    std::vector<Brick> bricks;
    std::vector<Splitter> splitters;
    std::vector<std::vector<int>> unitBricks, unitSplitters;
    std::vector<double> unitAreas;  // HydroUnit::GetArea, for the lateral processes
    std::vector<int> basinBricks;
    std::vector<int> outletFluxes;
    std::vector<Record> records;
    std::vector<Event> events;
    std::vector<std::pair<int, int>> snowToIce;  // (glacier brick, snowpack brick)
    std::vector<int> dhBricks;
    std::vector<double> dhUnitAreas, dhAreas, dhVolumes;
    int dhRows = 0, dhLastRow = 0;
    double dhInitialWE = 0;
    const double* forcing[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> forcingStore[4];
    double outletTotal = 0;
    // Processor state (base/Processor.h)
    std::vector<int> iterableBricks;
    std::vector<double*> stateVariableChanges;
    std::vector<double> changeRatesNoSolver, ratesBeforeSweep;
    std::vector<double> k1, k2, k3, k4, combined, state;
    int solvableConnectionCount = 0, directConnectionCount = 0;
    long unconverged = 0;
    std::string error;

    // interception:gr4j publishes the NET evaporative demand through the unit's PET forcing object (Forcing::UpdateValue,
    // ProcessInterceptionGR4J.cpp:29-34) until ModelHydro::UpdateForcing refreshes it at the next step
    std::vector<double> petOverride;
    std::vector<int> petOverrideStep;
    double Forcing(int var, int unit) const {  // base/Forcing.cpp:12-16, TimeSeriesData.cpp:57-60
        if (var == 2 && !petOverrideStep.empty() && petOverrideStep[unit] == cursor) return petOverride[unit];
        return forcing[var][static_cast<size_t>(cursor) * nUnits + unit];
    }
};

namespace {

struct Fail {
    std::string msg;
};

// ---- containers -------------------------------------------------------------------------------

// WaterContainer.h:151-157
double ContentWithChanges(const Container& c) {
    if (c.infinite) return INFINITY;
    return c.content + c.dyn + c.stat;
}
// WaterContainer.h:164-170
double ContentWithDynamicChanges(const Container& c) {
    if (c.infinite) return INFINITY;
    return c.content + c.dyn;
}
// WaterContainer.h:228-231
bool IsNotEmpty(const Container& c) {
    double content = ContentWithChanges(c);
    return GreaterThan(content, 0, EPSILON_F) && GreaterThan(content, 0, PRECISION);
}
// WaterContainer.cpp:245-248
double TargetFillingRatio(const Container& c) {
    return std::max(0.0, std::min(1.0, ContentWithChanges(c) / static_cast<double>(c.capacity)));
}
// Snowpack.cpp:139-141
bool HasSnow(const hbo_model& m, int snowpackBrick) { return IsNotEmpty(m.containers[m.bricks[snowpackBrick].second]); }

// WaterContainer.cpp:237-239, IceContainer.cpp:22-32
bool ContentAccessible(const hbo_model& m, const Container& c) {
    if (c.contentKind == HBO_ICE && c.noMeltWhenSnowCover) {
        if (c.relatedSnowpackBrick < 0) throw Fail{"No snowpack provided for the glacier melt limitation."};
        if (HasSnow(m, c.relatedSnowpackBrick)) return false;
    }
    return ContentWithChanges(c) > 0;
}

// ---- fluxes -----------------------------------------------------------------------------------

// FluxForcing.cpp:11-14, FluxToBrickInstantaneous.cpp:13-15, others return _amount
double FluxGetAmount(hbo_model& m, Flux& f) {
    if (f.kind == HBO_F_FORCING) {
        f.amount = m.Forcing(f.forcingVar, f.unit);
        return f.amount;
    }
    if (f.kind == HBO_F_TO_BRICK_INSTANT) return 0;
    return f.amount;
}

// Flux.cpp:20-26, FluxToBrick.cpp:17-24, FluxToBrickInstantaneous.cpp:21-38, FluxSimple.cpp:14-17
void FluxUpdate(hbo_model& m, Flux& f, double amount) {
    if (f.kind == HBO_F_SIMPLE) {
        f.amount = amount;
        return;
    }
    if (f.kind == HBO_F_TO_BRICK_INSTANT) {
        if (f.fractionTotal != 1.0) {
            f.amount = amount * f.fractionTotal;
        } else {
            f.amount = amount;
        }
        Container& c = m.containers[f.targetContainer];
        if (!c.infinite) c.stat += f.amount;  // WaterContainer.cpp:50-53
        return;
    }
    if (f.fractionTotal < 1.0) {
        f.amount = amount * f.fractionTotal;
    } else {
        f.amount = amount;
    }
}

// WaterContainer.cpp:228-235
double SumIncomingFluxes(hbo_model& m, Container& c) {
    double sum = 0;
    for (int i : c.inputs) sum += FluxGetAmount(m, m.fluxes[i]);
    return sum;
}

// ---- flux-constraint step ---------------------------------------------------------------------

// WaterContainer.cpp:177-194
void SetOutgoingRatesToZero(hbo_model& m, Container& c, int containerId) {
    for (int ip : m.bricks[c.brick].processes) {
        Process& p = m.processes[ip];
        if (p.container != containerId) continue;
        for (int io : p.outputs) {
            double* r = m.fluxes[io].changeRate;
            if (r) *r = 0;
        }
    }
}

// WaterContainer.cpp:55-175 (+ IceContainer.cpp:10-20)
void ApplyConstraints(hbo_model& m, int containerId, double timeStep) {
    Container& c = m.containers[containerId];
    if (c.contentKind == HBO_ICE && c.noMeltWhenSnowCover) {
        if (c.relatedSnowpackBrick < 0) throw Fail{"No snowpack provided for the glacier melt limitation."};
        if (HasSnow(m, c.relatedSnowpackBrick)) SetOutgoingRatesToZero(m, c, containerId);
    }
    if (c.infinite) return;

    std::vector<double*> outgoingRates;
    double outputs = 0;
    for (int ip : m.bricks[c.brick].processes) {
        Process& p = m.processes[ip];
        if (p.container != containerId) continue;
        for (int io : p.outputs) {
            double* changeRate = m.fluxes[io].changeRate;
            if (changeRate == nullptr) continue;
            if (*changeRate < 0) {
                *changeRate = 0;
            } else if (*changeRate > 10000) {
                throw Fail{"Change rate is too high."};
            }
            outgoingRates.push_back(changeRate);
            outputs += *changeRate;
        }
    }

    std::vector<double*> incomingRates;
    double inputs = 0;
    double inputsStatic = 0;
    for (int ii : c.inputs) {
        Flux& in = m.fluxes[ii];
        if (in.kind == HBO_F_TO_BRICK_INSTANT) {
            inputsStatic += in.amount;  // GetRealAmount()
            continue;
        }
        if (in.kind == HBO_F_FORCING || in.isStatic) {
            inputsStatic += FluxGetAmount(m, in);
            continue;
        }
        double* changeRate = in.changeRate;
        if (changeRate == nullptr) continue;
        if (*changeRate < 0) *changeRate = 0;
        incomingRates.push_back(changeRate);
        inputs += *changeRate;
    }

    double change = inputs - outputs;
    double content = ContentWithDynamicChanges(c);

    // Avoid negative content (unless the container is allowed to go negative, WaterContainer.cpp:116-118)
    if (!c.allowNegative && change < 0 && content + inputsStatic + change * timeStep < 0) {
        double diff = (content + inputsStatic + change * timeStep) / timeStep;
        for (double* rate : outgoingRates) {
            if (NearlyZero(*rate, EPSILON_D)) continue;
            if (NearlyEqual(diff, change, PRECISION)) {
                *rate = 0;
                continue;
            }
            *rate += diff * std::abs((*rate) / outputs);
        }
    }

    // Enforce maximum capacity
    if (c.hasCapacity) {
        if (content + inputsStatic + change * timeStep > c.capacity) {
            double diff = (content + inputsStatic + change * timeStep - c.capacity) / timeStep;
            if (c.overflowProcess >= 0) {
                double* r = m.fluxes[m.processes[c.overflowProcess].outputs[0]].changeRate;
                if (r != nullptr) {
                    *r = diff;
                    return;
                }
                throw Fail{"Overflow exists but has no change rate pointer"};
            }
            if (content + inputsStatic > c.capacity) {
                throw Fail{"Forcing is coming directly into a brick with limited capacity and no overflow."};
            }
            for (double* rate : incomingRates) {
                if (NearlyZero(*rate, EPSILON_D)) continue;
                *rate -= diff * std::abs((*rate) / inputs);
            }
        }
    }
}

// WaterContainer.cpp:196-216
void FinalizeContainer(Container& c) {
    if (c.infinite) return;
    c.content += c.dyn + c.stat;
    c.dyn = 0;
    c.stat = 0;
    if (c.allowNegative) return;  // WaterContainer.cpp:201-203
    if (NearlyZero(c.content, PRECISION)) {
        c.content = 0;
        return;
    }
    if (c.content < 0 - PRECISION) c.content = 0;  // logged as an error in the reference
}

// ---- bricks -----------------------------------------------------------------------------------

// LandCover.h:91-93, SurfaceComponent.cpp:31-40, Brick.h:170
bool IsNull(const hbo_model& m, const Brick& b) {
    if (b.kind == HBO_BRICK_GENERIC_LAND_COVER || b.kind == HBO_BRICK_GLACIER) {
        return NearlyZero(b.areaFraction, PRECISION);
    }
    if (b.kind == HBO_BRICK_SNOWPACK) {
        if (NearlyZero(b.areaFraction, PRECISION)) return true;
        if (b.parent >= 0 && IsNull(m, m.bricks[b.parent])) return true;
        return false;
    }
    return false;
}

// Brick.cpp:166-168, Snowpack.cpp:109-112, Glacier.cpp:116-119
void UpdateContentFromInputs(hbo_model& m, Brick& b) {
    if (b.kind == HBO_BRICK_SNOWPACK) {
        Container& snow = m.containers[b.second];
        double s = SumIncomingFluxes(m, snow);
        if (!snow.infinite) snow.stat += s;
    } else if (b.kind == HBO_BRICK_GLACIER) {
        Container& ice = m.containers[b.second];
        double s = SumIncomingFluxes(m, ice);
        if (!ice.infinite) ice.dyn += s;
    }
    Container& w = m.containers[b.water];
    double s = SumIncomingFluxes(m, w);
    if (!w.infinite) w.dyn += s;
}

// Brick.cpp:170-172, Snowpack.cpp:114-117, Glacier.cpp:121-124
void BrickApplyConstraints(hbo_model& m, Brick& b, double dt) {
    if (b.second >= 0) ApplyConstraints(m, b.second, dt);
    ApplyConstraints(m, b.water, dt);
}

// Brick.cpp:127-135, Snowpack.cpp:58-67, Glacier.cpp:71-74
void ProcessFinalize(hbo_model& m, Process& p);
void BrickFinalize(hbo_model& m, Brick& b) {
    if (b.second >= 0) FinalizeContainer(m.containers[b.second]);
    FinalizeContainer(m.containers[b.water]);
    for (int ip : b.processes) ProcessFinalize(m, m.processes[ip]);  // Brick.cpp:130-134
}

// ---- processes --------------------------------------------------------------------------------

// TimeMachine::GetCurrentDayOfYear (TimeMachine.cpp:87-98) over UtilsDateTime::FromMJD (civil date of floor(mjd)).
int DayOfYear(double mjd) {
    if (mjd <= 0) return 1;
    long long z = static_cast<long long>(std::floor(mjd)) + 2400001LL - 2440588LL + 719468LL;  // days since 0000-03-01
    const long long era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = static_cast<unsigned>(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const unsigned doyMarch = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doyMarch + 2) / 153;
    const unsigned day = doyMarch - (153 * mp + 2) / 5 + 1;
    const unsigned month = mp < 10 ? mp + 3 : mp - 9;
    const long long year = static_cast<long long>(yoe) + era * 400 + (month <= 2);
    static const int cumMonthDays[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    const bool leap = (year % 4 == 0 && (year % 100 != 0 || year % 400 == 0));
    int doy = cumMonthDays[month - 1] + static_cast<int>(day);
    if (leap && month > 2) doy += 1;
    return doy;
}

void StoreRates1(Process& p, double v) { p.rates.assign(1, v); }

// ProcessRoutingHBV::_cumulativeWeight / _recomputeUH (ProcessRoutingHBV.cpp:118-151): triangular weighting function
double CumulativeWeightHBV(double t, double maxbas) {
    if (t <= 0.0) return 0.0;
    if (t >= maxbas) return 1.0;
    double ratio = t / maxbas;
    if (ratio <= 0.5) return 2.0 * ratio * ratio;
    return 1.0 - 2.0 * (1.0 - ratio) * (1.0 - ratio);
}
void RecomputeUH(Process& p) {
    double maxbas = p.params[0];
    if (maxbas < 1.0) maxbas = 1.0;  // shorter than the timestep: pass-through
    p.lastMaxbas = p.params[0];
    int n = static_cast<int>(std::ceil(maxbas));
    p.uhOrd.resize(n);
    for (int j = 1; j <= n; ++j)
        p.uhOrd[j - 1] = CumulativeWeightHBV(static_cast<double>(j), maxbas) - CumulativeWeightHBV(static_cast<double>(j - 1), maxbas);
    p.stuh.resize(n, 0.0);
}
// helpers/GR4JFormulas.h (Perrin et al., 2003)
double Gr4jInfiltration(double S, double Pn, double X1) {
    if (Pn <= 0.0 || X1 <= 0.0) return 0.0;
    double Sr = S / X1;
    double tws = std::tanh(std::min(Pn / X1, 13.0));
    return X1 * (1.0 - Sr * Sr) * tws / (1.0 + Sr * tws);
}
double Gr4jEvaporation(double S, double En, double X1) {
    if (En <= 0.0 || X1 <= 0.0) return 0.0;
    double Sr = S / X1;
    double tws = std::tanh(std::min(En / X1, 13.0));
    return S * (2.0 - Sr) * tws / (1.0 + (1.0 - Sr) * tws);
}
double Gr4jPercolation(double S, double X1) {
    if (X1 <= 0.0) return 0.0;
    double ratio = S / X1;
    return S * (1.0 - std::pow(1.0 + (ratio * ratio * ratio * ratio) / 25.62890625, -0.25));
}
// ProcessRoutingGR4J::_sh1 / _sh2 / _recomputeUH (ProcessRoutingGR4J.cpp:176-216)
double Gr4jSh1(double t, double x4) {
    if (t <= 0.0) return 0.0;
    if (t >= x4) return 1.0;
    return std::pow(t / x4, 2.5);
}
double Gr4jSh2(double t, double x4) {
    if (t <= 0.0) return 0.0;
    if (t >= 2.0 * x4) return 1.0;
    if (t < x4) return 0.5 * std::pow(t / x4, 2.5);
    return 1.0 - 0.5 * std::pow(2.0 - t / x4, 2.5);
}
// helpers/GR6JFormulas.h
double Gr6jExponentialStoreOutflow(double Rexp, double X6) {
    if (X6 <= 0.0) return std::max(0.0, Rexp);
    double ar = std::max(-33.0, std::min(33.0, Rexp / X6));
    if (ar > 7.0) return Rexp + X6 / std::exp(ar);
    if (ar < -7.0) return X6 * std::exp(ar);
    return X6 * std::log(std::exp(ar) + 1.0);
}
void RecomputeUHGr4j(Process& p) {
    double X4 = p.params[2];
    if (X4 <= 0) X4 = 0.5;
    int n1 = static_cast<int>(std::ceil(X4));
    int n2 = static_cast<int>(std::ceil(2.0 * X4));
    p.uh1Ord.resize(n1);
    for (int j = 1; j <= n1; ++j) p.uh1Ord[j - 1] = Gr4jSh1(static_cast<double>(j), X4) - Gr4jSh1(static_cast<double>(j - 1), X4);
    p.uh2Ord.resize(n2);
    for (int j = 1; j <= n2; ++j) p.uh2Ord[j - 1] = Gr4jSh2(static_cast<double>(j), X4) - Gr4jSh2(static_cast<double>(j - 1), X4);
    p.stuh1.resize(n1, 0.0);
    p.stuh2.resize(n2, 0.0);
}
// ProcessRoutingHBV::Finalize (ProcessRoutingHBV.cpp:85-116), after the container has been finalized
void ProcessFinalize(hbo_model& m, Process& p) {
    if (p.type == HBO_P_ROUTING_GR4J) {  // ProcessRoutingGR4J::Finalize (ProcessRoutingGR4J.cpp:130-174)
        double X2 = p.params[0], X3 = p.params[1];
        double PR = SumIncomingFluxes(m, m.containers[p.container]);
        double prhu1 = 0.9 * PR, prhu2 = 0.1 * PR;
        for (int j = 0; j < static_cast<int>(p.stuh1.size()) - 1; ++j) p.stuh1[j] = p.stuh1[j + 1] + p.uh1Ord[j] * prhu1;
        p.stuh1.back() = p.uh1Ord.back() * prhu1;
        for (int j = 0; j < static_cast<int>(p.stuh2.size()) - 1; ++j) p.stuh2[j] = p.stuh2[j + 1] + p.uh2Ord[j] * prhu2;
        p.stuh2.back() = p.uh2Ord.back() * prhu2;
        double F = 0.0;
        if (X3 > 0) {
            double rr = p.gr_r / X3;
            F = X2 * rr * rr * rr * std::sqrt(rr);
        }
        p.gr_r = std::max(0.0, p.gr_r + p.stuh1[0] + F);
        double rr4 = 0.0;
        if (X3 > 0) {
            double ratio = p.gr_r / X3;
            rr4 = ratio * ratio * ratio * ratio;
        }
        p.gr_qr = p.gr_r * (1.0 - std::pow(1.0 + rr4, -0.25));
        p.gr_r -= p.gr_qr;
        p.gr_qd = std::max(0.0, p.stuh2[0] + F);
        return;
    }
    if (p.type == HBO_P_ROUTING_GR6J) {  // ProcessRoutingGR6J::Finalize (ProcessRoutingGR6J.cpp:152-212)
        double X2 = p.params[0], X3 = p.params[1], X5 = p.params[3], X6 = p.params[4];
        double PR = SumIncomingFluxes(m, m.containers[p.container]);
        double prhu1 = 0.9 * PR, prhu2 = 0.1 * PR;
        for (int j = 0; j < static_cast<int>(p.stuh1.size()) - 1; ++j) p.stuh1[j] = p.stuh1[j + 1] + p.uh1Ord[j] * prhu1;
        if (!p.stuh1.empty()) p.stuh1.back() = p.uh1Ord.back() * prhu1;
        for (int j = 0; j < static_cast<int>(p.stuh2.size()) - 1; ++j) p.stuh2[j] = p.stuh2[j + 1] + p.uh2Ord[j] * prhu2;
        if (!p.stuh2.empty()) p.stuh2.back() = p.uh2Ord.back() * prhu2;
        double F = 0.0;
        if (X3 > 0) F = X2 * (p.gr_r / X3 - X5);
        double uh1_out = p.stuh1.empty() ? 0.0 : p.stuh1[0];
        double uh2_out = p.stuh2.empty() ? 0.0 : p.stuh2[0];
        p.gr_r = std::max(0.0, p.gr_r + 0.6 * uh1_out + F);
        double rr4 = 0.0;
        if (X3 > 0) {
            double ratio = p.gr_r / X3;
            rr4 = ratio * ratio * ratio * ratio;
        }
        p.gr_qr = p.gr_r * (1.0 - std::pow(1.0 + rr4, -0.25));
        p.gr_r -= p.gr_qr;
        p.gr_rexp = p.gr_rexp + 0.4 * uh1_out + F;
        p.gr_qrexp = Gr6jExponentialStoreOutflow(p.gr_rexp, X6);
        p.gr_rexp -= p.gr_qrexp;
        p.gr_qd = std::max(0.0, uh2_out + F);
        return;
    }
    if (p.type != HBO_P_ROUTING_HBV) return;
    Container& c = m.containers[p.container];
    double delivered = m.fluxes[p.outputs[0]].amount;
    double content = c.content;
    double inflow = content - p.previousContent + delivered;
    p.previousContent = content;
    double carryOver = p.stuh[0] + p.uhOrd[0] * inflow - delivered;
    for (int j = 0; j < static_cast<int>(p.stuh.size()) - 1; ++j) p.stuh[j] = p.stuh[j + 1] + p.uhOrd[j + 1] * inflow;
    p.stuh.back() = 0.0;
    p.stuh[0] += carryOver;
    p.processStorage = std::accumulate(p.stuh.begin(), p.stuh.end(), 0.0);
}

void GetRates(hbo_model& m, Process& p) {
    Container& c = m.containers[p.container];
    const int unit = m.bricks[p.brick].unit;
    switch (p.type) {
        case HBO_P_MELT_DEGREE_DAY: {  // ProcessMeltDegreeDay.cpp:49-60
            if (!ContentAccessible(m, c)) {
                StoreRates1(p, 0);
                return;
            }
            double melt = 0;
            const float ddf = p.params[0], tm = p.params[1];
            double t = m.Forcing(HBO_V_TEMPERATURE, unit);
            if (t >= tm) melt = (t - tm) * ddf;
            StoreRates1(p, melt);
            return;
        }
        case HBO_P_MELT_TEMPERATURE_INDEX: {  // ProcessMeltTemperatureIndex.cpp:62-74
            if (!ContentAccessible(m, c)) {
                StoreRates1(p, 0);
                return;
            }
            double melt = 0;
            const float mf = p.params[0], tm = p.params[1], rc = p.params[2];
            double t = m.Forcing(HBO_V_TEMPERATURE, unit);
            if (t >= tm) melt = (t - tm) * (mf + rc * m.Forcing(HBO_V_SOLAR_RADIATION, unit));
            StoreRates1(p, melt);
            return;
        }
        case HBO_P_LATERAL_SNOW_SLIDE: {  // ProcessLateralSnowSlide.cpp:48-118
            if (p.outputs.empty()) {
                p.rates.clear();
                return;
            }
            const float coeff = p.params[0], expo = p.params[1], minSlope = p.params[2], maxSlope = p.params[3],
                        minSnowHoldingDepth = p.params[4], maxSnowDepth = p.params[5];
            const float sweToDepthFactor = 1000.0 / 250.0;  // constants::waterDensity / constants::snowDensity
            double swe = ContentWithChanges(c);
            double snowDepth = swe * sweToDepthFactor;
            float slope = std::max(p.slopeDeg, minSlope);
            double snowHoldingThresholdMeters = coeff * pow(slope, expo);
            double snowHoldingThreshold = snowHoldingThresholdMeters * 1000.0;
            if (p.slopeDeg > maxSlope) snowHoldingThreshold = minSnowHoldingDepth;
            snowHoldingThreshold = std::max<double>(snowHoldingThreshold, minSnowHoldingDepth);
            if (maxSnowDepth > 0.0f) snowHoldingThreshold = std::min<double>(snowHoldingThreshold, maxSnowDepth);
            double excessSwe = 0.0;
            if (snowDepth > snowHoldingThreshold) {
                double excessSnowDepth = snowDepth - snowHoldingThreshold;
                excessSwe = excessSnowDepth / sweToDepthFactor;
            }
            p.rates.assign(p.outputs.size(), 0.0);
            if (excessSwe <= 0.0) return;
            if (excessSwe > 1000.0) excessSwe = 1000.0;
            const Brick& origin = m.bricks[p.brick];
            const double originFraction = m.bricks[origin.parent].areaFraction;  // SurfaceComponent::GetParentAreaFraction
            for (size_t i = 0; i < p.outputs.size(); ++i) {
                const Brick& target = m.bricks[p.lateralTargets[i]];
                double targetFraction = m.bricks[target.parent].areaFraction;
                if (NearlyZero(targetFraction, PRECISION)) {
                    p.rates[i] = 0.0;
                    continue;
                }
                p.rates[i] = excessSwe * p.weights[i] * targetFraction;
                // AvoidUnrealisticAccumulation (ProcessLateralSnowSlide.cpp:120-134): Snowpack::GetContent(Snow)
                double targetSwe = m.containers[target.second].content;
                if (maxSnowDepth > 0.0f && targetSwe > 1.5 * maxSnowDepth / sweToDepthFactor) p.rates[i] = 0;
                // ProcessLateral::ComputeFractionAreas (ProcessLateral.cpp:67-76) -> Flux::SetFractionUnitArea
                double fractionAreas = (m.unitAreas[origin.unit] * originFraction) / (m.unitAreas[target.unit] * targetFraction);
                Flux& f = m.fluxes[p.outputs[i]];
                f.fractionUnitArea = fractionAreas;
                f.fractionTotal = f.fractionUnitArea * f.fractionLandCover;
            }
            return;
        }
        case HBO_P_ET_SOCONT: {  // ProcessETSocont.cpp:35-38 (exponent is a float member = 0.5)
            const float exponent = 0.5f;
            StoreRates1(p, m.Forcing(HBO_V_PET, unit) * pow(TargetFillingRatio(c), exponent));
            return;
        }
        case HBO_P_INFILTRATION_SOCONT: {  // ProcessInfiltrationSocont.cpp:17-23, ProcessInfiltration.cpp:35-45
            Container& target = m.containers[m.bricks[p.targetBrick].water];
            if (static_cast<double>(target.capacity) <= 0) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, ContentWithChanges(c) * (1 - pow(TargetFillingRatio(target), 2)));
            return;
        }
        case HBO_P_RUNOFF_SOCONT: {  // ProcessRunoffSocont.cpp:41-56
            double storageShape = 2.0;
            const double dt = 86400;
            const double exponent = 5.0 / 3.0;
            const float beta = p.params[0];
            double area = p.unitArea;  // surface_runoff is a storage: no area fraction (ProcessRunoffSocont.cpp:33-39)
            double h = ContentWithChanges(c) * storageShape / 1000;
            double qQuick = beta * pow(p.slope, 0.5) * pow(h, exponent) / area;
            double dh = qQuick * storageShape * dt;
            double runoff = (dh / storageShape) * 1000;
            StoreRates1(p, std::min(runoff, ContentWithChanges(c)));
            return;
        }
        case HBO_P_OUTFLOW_LINEAR: {  // ProcessOutflowLinear.cpp:19-21
            const float k = p.params[0];
            StoreRates1(p, k * ContentWithChanges(c));
            return;
        }
        case HBO_P_OUTFLOW_DIRECT:  // ProcessOutflowDirect.cpp:42-44
            StoreRates1(p, std::max(ContentWithChanges(c), 0.0));
            return;
        case HBO_P_OUTFLOW_REST: {  // ProcessOutflowRest.cpp:13-41
            double rest = ContentWithChanges(c);
            Brick& b = m.bricks[p.brick];
            if (b.needsSolver) {
                for (int ip : b.processes) {
                    Process& other = m.processes[ip];
                    if (&other == &p) continue;
                    for (int io : other.outputs) {
                        double* rate = m.fluxes[io].changeRate;
                        if (rate != nullptr) rest -= *rate;
                    }
                }
            }
            StoreRates1(p, std::max(rest, 0.0));
            return;
        }
        case HBO_P_OVERFLOW:  // ProcessOutflowOverflow.cpp:17-19
            StoreRates1(p, 0);
            return;
        case HBO_P_PERCOLATION_CONSTANT: {  // ProcessPercolationConstant.cpp:23-25
            const float r = p.params[0];
            StoreRates1(p, r);
            return;
        }
        case HBO_P_TRANSFORM_SNOW_ICE_CONSTANT: {  // ProcessTransformSnowToIceConstant.cpp:27-29
            const float r = p.params[0];
            StoreRates1(p, r);
            return;
        }
        case HBO_P_TRANSFORM_SNOW_ICE_SWAT: {  // ProcessTransformSnowToIceSwat.cpp:31-39
            const int doy = DayOfYear(m.startMjd + m.cursor * m.dt);
            const bool northHemisphere = !(p.params[1] == 0);
            const int daysRef = northHemisphere ? 81 : 264;
            const double pi = 3.1415926535897932384626433832795;  // GlobVars.h:6
            auto coeff = static_cast<float>(p.params[0] * (1 + std::sin(2.0f * pi * (doy - daysRef) / 365.0f)));
            StoreRates1(p, coeff * ContentWithChanges(c));
            return;
        }
        case HBO_P_SUBLIMATION_CONSTANT: {  // ProcessSublimationConstant.cpp:31-36
            if (!ContentAccessible(m, c)) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, static_cast<double>(p.params[0]));
            return;
        }
        case HBO_P_OUTFLOW_SNOW_HOLDING: {  // ProcessOutflowSnowHolding.cpp:37-57
            const Brick& b = m.bricks[p.brick];
            if (b.kind != HBO_BRICK_SNOWPACK) {
                StoreRates1(p, 0);
                return;
            }
            const float whc = p.params[0];
            double swe = ContentWithChanges(m.containers[b.second]);
            double excess = ContentWithChanges(c) - whc * swe;
            if (excess <= 0) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, excess / m.dt);  // the time machine is wired to every process (ModelBuilder.cpp:126, 229)
            return;
        }
        case HBO_P_REFREEZE_DEGREE_DAY: {  // ProcessRefreezeDegreeDay.cpp:72-87
            const Process* melt = nullptr;  // FindSiblingMeltProcess (:52-64): the first melt:degree_day of the brick
            for (int ip : m.bricks[p.brick].processes)
                if (m.processes[ip].type == HBO_P_MELT_DEGREE_DAY) {
                    melt = &m.processes[ip];
                    break;
                }
            if (melt == nullptr) {
                StoreRates1(p, 0);
                return;
            }
            const float cfr = p.params[0];
            const double ddf = melt->params[0], meltingTemperature = melt->params[1];
            const double t = m.Forcing(HBO_V_TEMPERATURE, unit);
            if (t >= meltingTemperature) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, cfr * ddf * (meltingTemperature - t));
            return;
        }
        case HBO_P_INFILTRATION_HBV: {  // ProcessInfiltrationHBV.cpp:36-46
            Container& target = m.containers[m.bricks[p.targetBrick].water];
            double in = ContentWithChanges(c);
            double fc = static_cast<double>(target.capacity);
            if (in <= 0 || !(fc > 0)) {
                StoreRates1(p, 0);
                return;
            }
            double ratio = std::clamp(TargetFillingRatio(target), 0.0, 1.0);
            StoreRates1(p, in * (1.0 - std::pow(ratio, static_cast<double>(p.params[0]))));
            return;
        }
        case HBO_P_ET_HBV: {  // ProcessETHBV.cpp:47-58; params lp, et_correction_factor
            double pet = static_cast<double>(p.params[1]) * m.Forcing(HBO_V_PET, unit);
            double threshold = static_cast<double>(p.params[0]) * static_cast<double>(c.capacity);
            if (threshold <= 0) {
                StoreRates1(p, pet);
                return;
            }
            double ratio = std::min(1.0, ContentWithChanges(c) / threshold);
            StoreRates1(p, pet * ratio);
            return;
        }
        case HBO_P_CAPILLARY_HBV: {  // ProcessCapillaryHBV.cpp:66-96
            p.rates.assign(p.capTargets.size(), 0.0);
            for (size_t i = 0; i < p.capTargets.size(); ++i) {
                Container& target = m.containers[m.bricks[p.capTargets[i]].water];
                if (!(static_cast<double>(target.capacity) > 0)) continue;
                double weight = 0.0;
                for (int ib : p.capWeights[i]) weight += m.bricks[ib].areaFraction;
                if (p.capWeights[i].empty()) weight = 1.0;
                double deficit = 1.0 - std::clamp(TargetFillingRatio(target), 0.0, 1.0);
                p.rates[i] = p.params[0] * weight * deficit;
            }
            return;
        }
        case HBO_P_RUNOFF_HBV: {  // ProcessRunoffHBV.cpp:40-47; params response_factor, alpha
            double uz = ContentWithChanges(c);
            if (uz <= 0) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, p.params[0] * std::pow(uz, 1.0 + static_cast<double>(p.params[1])));
            return;
        }
        case HBO_P_ROUTING_HBV: {  // ProcessRoutingHBV.cpp:69-83
            if (p.params[0] != p.lastMaxbas) RecomputeUH(p);
            double in = SumIncomingFluxes(m, c);
            StoreRates1(p, std::max(0.0, p.stuh[0]) + p.uhOrd[0] * in);
            return;
        }
        case HBO_P_INTERCEPTION_GR4J: {  // ProcessInterceptionGR4J.cpp:29-34
            double P = ContentWithChanges(c);
            double E = m.Forcing(HBO_V_PET, unit);
            m.petOverride[unit] = std::max(0.0, E - P);
            m.petOverrideStep[unit] = m.cursor;
            StoreRates1(p, std::min(P, E));
            return;
        }
        case HBO_P_PRODUCTION_GR4J: {  // ProcessProductionGR4J.cpp:29-63
            double X1 = static_cast<double>(c.capacity);
            if (!(X1 > 0)) {
                StoreRates1(p, 0);
                return;
            }
            double S0 = c.content;
            double Pn = ContentWithChanges(c) - ContentWithDynamicChanges(c);
            double En = m.Forcing(HBO_V_PET, unit);
            double Ps = Gr4jInfiltration(S0, Pn, X1);
            double S1 = S0 + Ps;
            double Es = Gr4jEvaporation(S1, En, X1);
            double S2 = S1 - Es;
            double Perc = Gr4jPercolation(S2, X1);
            StoreRates1(p, (Pn - Ps) + Perc);
            return;
        }
        case HBO_P_ET_GR4J: {  // ProcessETGR4J.cpp:30-55
            double En = m.Forcing(HBO_V_PET, unit);
            double X1 = static_cast<double>(c.capacity);
            if (En <= 0 || !(X1 > 0)) {
                StoreRates1(p, 0);
                return;
            }
            double S0 = c.content;
            double Pn = ContentWithChanges(c) - ContentWithDynamicChanges(c);
            double S1 = S0 + Gr4jInfiltration(S0, Pn, X1);
            StoreRates1(p, Gr4jEvaporation(S1, En, X1));
            return;
        }
        case HBO_P_ROUTING_GR4J: {  // ProcessRoutingGR4J.cpp:61-128
            double X2 = p.params[0], X3 = p.params[1];
            RecomputeUHGr4j(p);  // (the reference's size test always fires: harmless, the buffers keep their state)
            double PR = SumIncomingFluxes(m, c);
            double prhu1 = 0.9 * PR, prhu2 = 0.1 * PR;
            double uh1_out = (p.stuh1.size() > 1 ? p.stuh1[1] : 0.0) + p.uh1Ord[0] * prhu1;
            double uh2_out = (p.stuh2.size() > 1 ? p.stuh2[1] : 0.0) + p.uh2Ord[0] * prhu2;
            double F = 0.0;
            if (X3 > 0) {
                double rr = p.gr_r / X3;
                F = X2 * rr * rr * rr * std::sqrt(rr);
            }
            double r_pred = std::max(0.0, p.gr_r + uh1_out + F);
            double rr4 = 0.0;
            if (X3 > 0) {
                double ratio = r_pred / X3;
                rr4 = ratio * ratio * ratio * ratio;
            }
            p.gr_qr = r_pred * (1.0 - std::pow(1.0 + rr4, -0.25));
            p.gr_qd = std::max(0.0, uh2_out + F);
            double exchangeToStore = r_pred - (p.gr_r + uh1_out);
            double exchangeToDirect = p.gr_qd - uh2_out;
            double netExchange = exchangeToStore + exchangeToDirect;
            if (netExchange != 0.0) c.stat += netExchange * m.dt;  // AddAmountToStaticContentChange
            StoreRates1(p, p.gr_qr + p.gr_qd);
            return;
        }
        case HBO_P_ROUTING_GR6J: {  // ProcessRoutingGR6J.cpp:98-150
            double X2 = p.params[0], X3 = p.params[1], X5 = p.params[3], X6 = p.params[4];
            double clampedX4 = p.params[2] <= 0.0 ? 0.5 : static_cast<double>(p.params[2]);
            if (static_cast<int>(std::ceil(clampedX4)) != static_cast<int>(p.uh1Ord.size())) RecomputeUHGr4j(p);
            double PR = SumIncomingFluxes(m, c);
            double prhu1 = 0.9 * PR, prhu2 = 0.1 * PR;
            double uh1_out = (p.stuh1.size() > 1 ? p.stuh1[1] : 0.0) + (p.uh1Ord.empty() ? 0.0 : p.uh1Ord[0]) * prhu1;
            double uh2_out = (p.stuh2.size() > 1 ? p.stuh2[1] : 0.0) + (p.uh2Ord.empty() ? 0.0 : p.uh2Ord[0]) * prhu2;
            double F = 0.0;
            if (X3 > 0) F = X2 * (p.gr_r / X3 - X5);
            double r_pred = std::max(0.0, p.gr_r + 0.6 * uh1_out + F);
            double rr4 = 0.0;
            if (X3 > 0) {
                double ratio = r_pred / X3;
                rr4 = ratio * ratio * ratio * ratio;
            }
            p.gr_qr = r_pred * (1.0 - std::pow(1.0 + rr4, -0.25));
            double rexp_pred = p.gr_rexp + 0.4 * uh1_out + F;
            p.gr_qrexp = Gr6jExponentialStoreOutflow(rexp_pred, X6);
            p.gr_qd = std::max(0.0, uh2_out + F);
            StoreRates1(p, std::max(0.0, p.gr_qr + p.gr_qrexp + p.gr_qd));
            return;
        }
        case HBO_P_SUBLIMATION_PET: {  // ProcessSublimationPET.cpp:46-51
            if (!ContentAccessible(m, c)) {
                StoreRates1(p, 0);
                return;
            }
            StoreRates1(p, m.Forcing(HBO_V_PET, unit) * p.params[0]);
            return;
        }
        default:
            throw Fail{"unknown process type"};
    }
}

// Process.cpp:480-487
const std::vector<double>& GetChangeRates(hbo_model& m, Process& p) {
    // routing:hbv and routing:gr6j bypass the empty-container shortcut (ProcessRoutingHBV.h:62-72, ProcessRoutingGR6J.cpp:91-96)
    if (p.type != HBO_P_ROUTING_HBV && p.type != HBO_P_ROUTING_GR6J && LessThanOrEqual(ContentWithChanges(m.containers[p.container]), 0, PRECISION)) {
        p.rates.assign(p.outputs.size(), 0.0);
        return p.rates;
    }
    GetRates(m, p);
    return p.rates;
}

// Process.cpp:489-493, ProcessOutflowOverflow.cpp:21-26
void StoreInOutgoingFlux(hbo_model& m, Process& p, double* rate, int index) {
    if (p.type == HBO_P_OVERFLOW) *rate = 0;
    m.fluxes[p.outputs[index]].changeRate = rate;
}

// Process.cpp:495-508
void ApplyChange(hbo_model& m, Process& p, int connectionIndex, double rate, double dt) {
    if (rate < 0) rate = 0;
    Flux& f = m.fluxes[p.outputs[connectionIndex]];
    if (GreaterThan(rate, 0, PRECISION)) {
        FluxUpdate(m, f, rate * dt);
        Container& c = m.containers[p.container];
        if (!c.infinite) c.dyn -= rate * dt;  // WaterContainer.cpp:40-43
    } else {
        FluxUpdate(m, f, 0);
    }
}

// ---- splitters --------------------------------------------------------------------------------

void SplitterCompute(hbo_model& m, Splitter& s) {
    switch (s.type) {
        case HBO_S_SNOW_RAIN_LINEAR: {  // SplitterSnowRainLinear.cpp:49-64
            const float t0 = s.params[0], t1 = s.params[1];
            double precip = m.Forcing(HBO_V_PRECIPITATION, s.unit);
            double temp = m.Forcing(HBO_V_TEMPERATURE, s.unit);
            double rcf = s.params[2];
            double scf = s.params[3];
            if (temp <= t0) {
                FluxUpdate(m, m.fluxes[s.outputs[0]], 0);
                FluxUpdate(m, m.fluxes[s.outputs[1]], precip * scf);
            } else if (temp >= t1) {
                FluxUpdate(m, m.fluxes[s.outputs[0]], precip * rcf);
                FluxUpdate(m, m.fluxes[s.outputs[1]], 0);
            } else {
                double rainFraction = (temp - t0) / (t1 - t0);  // (t1 - t0) in float32
                FluxUpdate(m, m.fluxes[s.outputs[0]], precip * rainFraction * rcf);
                FluxUpdate(m, m.fluxes[s.outputs[1]], precip * (1 - rainFraction) * scf);
            }
            return;
        }
        case HBO_S_SNOW_RAIN_THRESHOLD: {  // SplitterSnowRainThreshold.cpp:45-54
            const float thr = s.params[0];
            double precip = m.Forcing(HBO_V_PRECIPITATION, s.unit);
            double temp = m.Forcing(HBO_V_TEMPERATURE, s.unit);
            double rcf = s.params[2];
            double scf = s.params[3];
            if (temp <= thr) {
                FluxUpdate(m, m.fluxes[s.outputs[0]], 0);
                FluxUpdate(m, m.fluxes[s.outputs[1]], precip * scf);
            } else {
                FluxUpdate(m, m.fluxes[s.outputs[0]], precip * rcf);
                FluxUpdate(m, m.fluxes[s.outputs[1]], 0);
            }
            return;
        }
        case HBO_S_RAIN: {  // SplitterRain.cpp:37-39
            double rcf = s.params[2];
            FluxUpdate(m, m.fluxes[s.outputs[0]], m.Forcing(HBO_V_PRECIPITATION, s.unit) * rcf);
            return;
        }
        case HBO_S_MULTI_FLUXES: {  // SplitterMultiFluxes.cpp:35-40
            for (int io : s.outputs) FluxUpdate(m, m.fluxes[io], FluxGetAmount(m, m.fluxes[s.inputs[0]]));
            return;
        }
        default:
            throw Fail{"unknown splitter type"};
    }
}

// ---- processor --------------------------------------------------------------------------------

int ConnectionCount(const hbo_model& m, const Brick& b) {
    int n = 0;
    for (int ip : b.processes) n += static_cast<int>(m.processes[ip].outputs.size());
    return n;
}

// Processor.cpp:58-112
void ConnectToElementsToSolve(hbo_model& m) {
    m.iterableBricks.clear();
    m.stateVariableChanges.clear();
    m.solvableConnectionCount = m.directConnectionCount = 0;
    for (int u = 0; u < m.nUnits; ++u) {
        for (int ib : m.unitBricks[u]) {
            Brick& b = m.bricks[ib];
            if (b.needsSolver) {
                m.iterableBricks.push_back(ib);
                // Brick.cpp:178-180 / Snowpack.cpp:119-129 / Glacier.cpp:126-136: water first, then second
                m.stateVariableChanges.push_back(&m.containers[b.water].dyn);
                if (b.second >= 0) m.stateVariableChanges.push_back(&m.containers[b.second].dyn);
                m.solvableConnectionCount += ConnectionCount(m, b);
            } else {
                m.directConnectionCount += ConnectionCount(m, b);
            }
        }
    }
    for (int ib : m.basinBricks) {
        Brick& b = m.bricks[ib];
        m.iterableBricks.push_back(ib);
        m.stateVariableChanges.push_back(&m.containers[b.water].dyn);
        if (b.second >= 0) m.stateVariableChanges.push_back(&m.containers[b.second].dyn);
        m.solvableConnectionCount += ConnectionCount(m, b);
    }
}

// Processor.cpp:191-209
void EnforceConstraints(hbo_model& m, std::vector<double>& rates, double dt) {
    const int maxSweeps = 10;
    for (int sweep = 0; sweep < maxSweeps; ++sweep) {
        m.ratesBeforeSweep = rates;
        for (int ib : m.iterableBricks) BrickApplyConstraints(m, m.bricks[ib], dt);
        bool stable = true;
        for (size_t i = 0; i < rates.size(); ++i) {
            if (!(std::fabs(rates[i] - m.ratesBeforeSweep[i]) <= PRECISION)) {
                stable = false;
                break;
            }
        }
        if (stable) return;
    }
    m.unconverged++;
}

// Processor.cpp:146-172
void EvaluateRates(hbo_model& m, std::vector<double>& rates, double dt, bool applyConstraints = true) {
    int iRate = 0;
    for (int ib : m.iterableBricks) {
        for (int ip : m.bricks[ib].processes) {
            Process& p = m.processes[ip];
            const std::vector<double>& processRates = GetChangeRates(m, p);
            for (size_t j = 0; j < processRates.size(); ++j) {
                rates[iRate] = processRates[j];
                if (applyConstraints) StoreInOutgoingFlux(m, p, &rates[iRate], static_cast<int>(j));
                iRate++;
            }
        }
    }
    if (applyConstraints) EnforceConstraints(m, rates, dt);
}

// Processor.cpp:174-189
void ConstrainRates(hbo_model& m, std::vector<double>& rates, double dt) {
    int iRate = 0;
    for (int ib : m.iterableBricks) {
        for (int ip : m.bricks[ib].processes) {
            Process& p = m.processes[ip];
            for (size_t j = 0; j < p.outputs.size(); ++j) {
                StoreInOutgoingFlux(m, p, &rates[iRate], static_cast<int>(j));
                iRate++;
            }
        }
    }
    EnforceConstraints(m, rates, dt);
}

// Processor.cpp:211-226
void ApplyRates(hbo_model& m, const std::vector<double>& rates, double dt) {
    int iRate = 0;
    for (int ib : m.iterableBricks) {
        Brick& b = m.bricks[ib];
        if (IsNull(m, b)) continue;
        UpdateContentFromInputs(m, b);
        for (int ip : b.processes) {
            Process& p = m.processes[ip];
            for (size_t ic = 0; ic < p.outputs.size(); ++ic) {
                ApplyChange(m, p, static_cast<int>(ic), rates[iRate], dt);
                iRate++;
            }
        }
    }
}

// Processor.cpp:228-235
void FinalizeTimeStep(hbo_model& m) {
    for (int ib : m.iterableBricks) {
        Brick& b = m.bricks[ib];
        if (IsNull(m, b)) continue;
        BrickFinalize(m, b);
    }
}

void ResetState(hbo_model& m) {  // Processor.cpp:140-144
    for (double* v : m.stateVariableChanges) *v = 0;
}
void HalveState(hbo_model& m) {  // SolverRK4.cpp:26-29 (Gather, *= 0.5, Scatter)
    for (size_t i = 0; i < m.stateVariableChanges.size(); ++i) m.state[i] = *m.stateVariableChanges[i];
    for (size_t i = 0; i < m.state.size(); ++i) m.state[i] *= 0.5;
    for (size_t i = 0; i < m.stateVariableChanges.size(); ++i) *m.stateVariableChanges[i] = m.state[i];
}

// SolverEulerExplicit.cpp:13-22
void SolveEuler(hbo_model& m, double dt) {
    EvaluateRates(m, m.k1, dt);
    ApplyRates(m, m.k1, dt);
    FinalizeTimeStep(m);
}

// SolverHeunExplicit.cpp:16-38
void SolveHeun(hbo_model& m, double dt) {
    EvaluateRates(m, m.k1, dt);
    ApplyRates(m, m.k1, dt);
    EvaluateRates(m, m.k2, dt, false);
    ResetState(m);
    for (size_t i = 0; i < m.combined.size(); ++i) m.combined[i] = (m.k1[i] + m.k2[i]) / 2;
    ConstrainRates(m, m.combined, dt);
    ApplyRates(m, m.combined, dt);
    FinalizeTimeStep(m);
}

// SolverRK4.cpp:19-63
void SolveRK4(hbo_model& m, double dt) {
    EvaluateRates(m, m.k1, dt);
    ApplyRates(m, m.k1, dt);
    HalveState(m);

    EvaluateRates(m, m.k2, dt, false);
    ResetState(m);
    ConstrainRates(m, m.k2, dt);
    ApplyRates(m, m.k2, dt);
    HalveState(m);

    EvaluateRates(m, m.k3, dt, false);
    ResetState(m);
    ConstrainRates(m, m.k3, dt);
    ApplyRates(m, m.k3, dt);

    EvaluateRates(m, m.k4, dt, false);
    ResetState(m);
    for (size_t i = 0; i < m.combined.size(); ++i) {
        m.combined[i] = (m.k1[i] + 2 * m.k2[i] + 2 * m.k3[i] + m.k4[i]) / 6;
    }
    ConstrainRates(m, m.combined, dt);
    ApplyRates(m, m.combined, dt);
    FinalizeTimeStep(m);
}

// ---- sequential solvers (base/SolverSequential.cpp and its four ComputeBrickRates implementations) ----

// SolverSequential.cpp:22-32
double TotalRateAt(hbo_model& m, Brick& b, double* contentDelta, double offset) {
    *contentDelta = offset;
    double total = 0;
    for (int ip : b.processes) {
        const std::vector<double>& processRates = GetChangeRates(m, m.processes[ip]);
        for (double v : processRates) total += v;
    }
    return total;
}

// SolverSequential.cpp:34-40
void StoreRatesAtCurrentContent(hbo_model& m, Brick& b, std::vector<double>& rates) {
    rates.clear();
    for (int ip : b.processes) {
        const std::vector<double>& processRates = GetChangeRates(m, m.processes[ip]);
        rates.insert(rates.end(), processRates.begin(), processRates.end());
    }
}

// SolverCrankNicolson.cpp:12-63 (crank) / SolverImplicitEuler.cpp:12-61: end-of-step content by bisection.
void ComputeBrickRatesBisection(hbo_model& m, Brick& b, double content, double inflow, double h, int iRateStart,
                                bool crank) {
    Container& c = m.containers[b.water];
    double* contentDelta = &c.dyn;
    std::vector<double> startRates, endRates;
    double startTotal = 0;
    if (crank) {
        startTotal = TotalRateAt(m, b, contentDelta, 0);
        StoreRatesAtCurrentContent(m, b, startRates);
    }
    double hi = content + h * std::max(inflow, 0.0);
    double maxOutflow = TotalRateAt(m, b, contentDelta, hi - content);
    double lo = crank ? content + h * (inflow - (startTotal + maxOutflow) / 2) : content + h * (inflow - maxOutflow);
    lo = std::max(lo, 0.0);  // AllowsNegativeContent() is false for every container of this family
    double endContent = hi;
    if (hi > lo) {
        for (int iter = 0; iter < 100; ++iter) {
            endContent = (lo + hi) / 2;
            double endTotal = TotalRateAt(m, b, contentDelta, endContent - content);
            double g = crank ? endContent - content - h * (inflow - (startTotal + endTotal) / 2)
                             : endContent - content - h * (inflow - endTotal);
            if (g > 0) {
                hi = endContent;
            } else {
                lo = endContent;
            }
            if (hi - lo < 1e-12) break;
        }
        endContent = (lo + hi) / 2;
    } else {
        endContent = lo;
    }
    TotalRateAt(m, b, contentDelta, endContent - content);
    StoreRatesAtCurrentContent(m, b, endRates);
    for (size_t i = 0; i < endRates.size(); ++i) {
        m.k1[iRateStart + i] = crank ? (startRates[i] + endRates[i]) / 2 : endRates[i];
    }
    *contentDelta = 0;
}

// SolverExponentialEuler.cpp:11-63
void ComputeBrickRatesExponential(hbo_model& m, Brick& b, double content, double inflow, double h, int iRateStart) {
    Container& c = m.containers[b.water];
    double* contentDelta = &c.dyn;
    std::vector<double> startRates, endRates;
    double startTotal = TotalRateAt(m, b, contentDelta, 0);
    StoreRatesAtCurrentContent(m, b, startRates);
    double scale = std::max(content, content + h * inflow);
    double eps = std::max(0.001, 0.001 * scale);
    double slope = (TotalRateAt(m, b, contentDelta, eps) - startTotal) / eps;
    if (slope <= 1e-10) {
        for (size_t i = 0; i < startRates.size(); ++i) m.k1[iRateStart + i] = startRates[i];
        *contentDelta = 0;
        return;
    }
    double netInflow = inflow - startTotal;
    double endContent = content + netInflow / slope * (1.0 - std::exp(-slope * h));
    endContent = std::max(endContent, 0.0);
    double totalRate = std::max(inflow - (endContent - content) / h, 0.0);
    TotalRateAt(m, b, contentDelta, endContent - content);
    StoreRatesAtCurrentContent(m, b, endRates);
    double sumWeights = 0;
    for (size_t i = 0; i < startRates.size(); ++i) sumWeights += startRates[i] + endRates[i];
    for (size_t i = 0; i < startRates.size(); ++i) {
        double weight = sumWeights > 0 ? (startRates[i] + endRates[i]) / sumWeights : 0.0;
        m.k1[iRateStart + i] = totalRate * weight;
    }
    *contentDelta = 0;
}

// SolverAnalyticLinear.cpp:11-59; HasLinearResponse / GetLinearResponseRate: ProcessOutflowLinear.h:36-46
void ComputeBrickRatesAnalytic(hbo_model& m, Brick& b, double content, double inflow, double h, int iRateStart) {
    int iRate = iRateStart;
    double totalLinearCoefficient = 0;
    double totalFrozenRate = 0;
    for (int ip : b.processes) {
        Process& p = m.processes[ip];
        if (p.type == HBO_P_OUTFLOW_LINEAR && p.outputs.size() == 1) {
            totalLinearCoefficient += p.params[0];
            m.k1[iRate] = 0;
            iRate++;
        } else {
            const std::vector<double>& processRates = GetChangeRates(m, p);
            for (double v : processRates) {
                m.k1[iRate] = v;
                totalFrozenRate += v;
                iRate++;
            }
        }
    }
    if (totalLinearCoefficient <= 0) return;
    double netInflow = inflow - totalFrozenRate;
    double decay = std::exp(-totalLinearCoefficient * h);
    double newContent = content * decay + netInflow / totalLinearCoefficient * (1.0 - decay);
    double linearOutflowRate = (netInflow * h - (newContent - content)) / h;
    linearOutflowRate = std::max(linearOutflowRate, 0.0);
    int iRateFill = iRateStart;
    for (int ip : b.processes) {
        Process& p = m.processes[ip];
        if (p.type == HBO_P_OUTFLOW_LINEAR && p.outputs.size() == 1) {
            m.k1[iRateFill] = linearOutflowRate * p.params[0] / totalLinearCoefficient;
            iRateFill++;
        } else {
            iRateFill += static_cast<int>(p.outputs.size());
        }
    }
}

// SolverSequential.cpp:42-88
void SolveSequential(hbo_model& m, double dt) {
    int iRate = 0;
    for (int ib : m.iterableBricks) {
        Brick& b = m.bricks[ib];
        int iRateStart = iRate;
        for (int ip : b.processes) {
            Process& p = m.processes[ip];
            for (size_t j = 0; j < p.outputs.size(); ++j) {
                m.k1[iRate] = 0;
                StoreInOutgoingFlux(m, p, &m.k1[iRate], static_cast<int>(j));
                iRate++;
            }
        }
        if (IsNull(m, b)) continue;
        Container& c = m.containers[b.water];
        double content = ContentWithChanges(c);
        double inflow = SumIncomingFluxes(m, c) / dt;
        switch (m.solver) {
            case HBO_SOLVER_CRANK_NICOLSON: ComputeBrickRatesBisection(m, b, content, inflow, dt, iRateStart, true); break;
            case HBO_SOLVER_IMPLICIT_EULER: ComputeBrickRatesBisection(m, b, content, inflow, dt, iRateStart, false); break;
            case HBO_SOLVER_EXPONENTIAL_EULER: ComputeBrickRatesExponential(m, b, content, inflow, dt, iRateStart); break;
            default: ComputeBrickRatesAnalytic(m, b, content, inflow, dt, iRateStart); break;
        }
        BrickApplyConstraints(m, b, dt);
        UpdateContentFromInputs(m, b);
        int iRateApply = iRateStart;
        for (int ip : b.processes) {
            Process& p = m.processes[ip];
            for (size_t j = 0; j < p.outputs.size(); ++j) {
                ApplyChange(m, p, static_cast<int>(j), m.k1[iRateApply], dt);
                iRateApply++;
            }
        }
    }
    FinalizeTimeStep(m);
}

// Processor.cpp:278-323
void ApplyDirectChanges(hbo_model& m, Brick& b, int& ptIndex, double dt) {
    UpdateContentFromInputs(m, b);

    int iRate = ptIndex;
    for (int ip : b.processes) {
        Process& p = m.processes[ip];
        for (size_t j = 0; j < p.outputs.size(); ++j) {
            m.changeRatesNoSolver[iRate] = 0;
            StoreInOutgoingFlux(m, p, &m.changeRatesNoSolver[iRate], static_cast<int>(j));
            iRate++;
        }
    }

    iRate = ptIndex;
    for (int ip : b.processes) {
        Process& p = m.processes[ip];
        const std::vector<double>& rates = GetChangeRates(m, p);
        int iRateCopy = iRate;
        for (double rate : rates) {
            m.changeRatesNoSolver[iRateCopy] = rate;
            iRateCopy++;
        }
        ApplyConstraints(m, p.container, dt);
        for (size_t j = 0; j < rates.size(); ++j) {
            ApplyChange(m, p, static_cast<int>(j), m.changeRatesNoSolver[iRate], dt);
            m.changeRatesNoSolver[iRate] = 0;
            iRate++;
            ptIndex++;
        }
    }
    BrickFinalize(m, b);
}

// Processor.cpp:237-276, SubBasin.cpp:322-329
void ProcessTimeStep(hbo_model& m, double dt) {
    int ptIndex = 0;
    for (int u = 0; u < m.nUnits; ++u) {
        for (int is : m.unitSplitters[u]) SplitterCompute(m, m.splitters[is]);
        for (int ib : m.unitBricks[u]) {
            Brick& b = m.bricks[ib];
            if (b.needsSolver) continue;
            if (IsNull(m, b)) continue;
            ApplyDirectChanges(m, b, ptIndex, dt);
        }
    }
    switch (m.solver) {
        case HBO_SOLVER_EULER: SolveEuler(m, dt); break;
        case HBO_SOLVER_HEUN: SolveHeun(m, dt); break;
        case HBO_SOLVER_RK4: SolveRK4(m, dt); break;
        default: SolveSequential(m, dt); break;
    }
    m.outletTotal = 0;
    for (int io : m.outletFluxes) m.outletTotal += FluxGetAmount(m, m.fluxes[io]);
}

// ---- actions ------------------------------------------------------------------------------------

// LandCover::SetAreaFraction (LandCover.cpp:27-37): push the fraction into the weighted out-fluxes.
void SetAreaFraction(hbo_model& m, Brick& b, double value) {
    b.areaFraction = value;
    for (int ip : b.processes) {
        for (int io : m.processes[ip].outputs) {
            Flux& f = m.fluxes[io];
            if (f.needsWeighting) {
                f.fractionLandCover = value;
                f.fractionTotal = f.fractionUnitArea * f.fractionLandCover;
            }
        }
    }
}

struct Slot {  // HydroUnit.cpp:12-52 (GetContentSlots): role 0 SelfWater, 1 SelfIce, 2 Snow, 3 SnowpackWater
    int container, role;
};

std::vector<Slot> ContentSlots(hbo_model& m, const Brick& cover) {
    std::vector<Slot> slots;
    auto add = [&](int c, int role) {
        if (c >= 0 && !m.containers[c].infinite) slots.push_back({c, role});
    };
    add(cover.water, 0);
    if (cover.kind == HBO_BRICK_GLACIER) add(cover.second, 1);
    if (cover.snowpack >= 0) {
        add(m.bricks[cover.snowpack].second, 2);
        add(m.bricks[cover.snowpack].water, 3);
    }
    return slots;
}

// HydroUnit::TransferLandCoverContent (HydroUnit.cpp:384-451)
void TransferLandCoverContent(hbo_model& m, Brick& changed, double oldFraction, double newFraction, Brick& generic,
                              double genericOldFraction, double genericNewFraction) {
    double delta = newFraction - oldFraction;
    if (NearlyZero(delta, EPSILON_D)) return;
    Brick *donor, *receiver;
    double fReceiverOld, fReceiverNew, fDonorNew, strip;
    if (delta > 0) {
        donor = &generic;
        receiver = &changed;
        fReceiverOld = oldFraction;
        fReceiverNew = newFraction;
        fDonorNew = genericNewFraction;
        strip = delta;
    } else {
        donor = &changed;
        receiver = &generic;
        fReceiverOld = genericOldFraction;
        fReceiverNew = genericNewFraction;
        fDonorNew = newFraction;
        strip = -delta;
    }
    if (NearlyZero(fReceiverNew, EPSILON_D)) return;
    std::vector<Slot> receiverSlots = ContentSlots(m, *receiver);
    std::vector<Slot> donorSlots = ContentSlots(m, *donor);
    auto hasRole = [](const std::vector<Slot>& v, int role) {
        for (const Slot& s : v)
            if (s.role == role) return true;
        return false;
    };
    auto depth = [&](const std::vector<Slot>& v, int role) -> double {
        for (const Slot& s : v)
            if (s.role == role) return m.containers[s.container].content;
        return 0.0;
    };
    double unmatchedDonorWater = 0;
    for (const Slot& s : donorSlots) {
        if ((s.role == 0 || s.role == 3) && !hasRole(receiverSlots, s.role)) {
            unmatchedDonorWater += m.containers[s.container].content;
        }
    }
    for (const Slot& s : receiverSlots) {
        double donorDepth = depth(donorSlots, s.role);
        if (s.role == 0) donorDepth += unmatchedDonorWater;
        double receiverDepth = m.containers[s.container].content;
        double newDepth = (receiverDepth * fReceiverOld + donorDepth * strip) / fReceiverNew;
        m.containers[s.container].content = newDepth;
    }
    if (NearlyZero(fDonorNew, EPSILON_D)) {
        for (const Slot& s : donorSlots) m.containers[s.container].content = 0;
    }
}

// HydroUnit::ChangeLandCoverAreaFraction + FixLandCoverFractionsTotal (HydroUnit.cpp:332-374, 459-490)
void ChangeLandCoverAreaFraction(hbo_model& m, int brickId, double fraction) {
    if (fraction < 0 || fraction > 1) throw Fail{"land cover fraction not in [0 .. 1]"};
    Brick& changed = m.bricks[brickId];
    if (changed.genericCover >= 0 && changed.genericCover != brickId) {
        Brick& generic = m.bricks[changed.genericCover];
        double oldFraction = changed.areaFraction;
        double delta = fraction - oldFraction;
        double genericOldFraction = generic.areaFraction;
        double genericNewFraction = genericOldFraction - delta;
        if (genericNewFraction < 0 - EPSILON_D) throw Fail{"generic land cover not large enough for the area change"};
        TransferLandCoverContent(m, changed, oldFraction, fraction, generic, genericOldFraction, genericNewFraction);
    }
    SetAreaFraction(m, changed, fraction);
    // FixLandCoverFractionsTotal: the unit's land covers in brick order
    int genericId = changed.genericCover >= 0 ? changed.genericCover : brickId;
    double total = 0;
    for (int ib : m.unitBricks[changed.unit]) {
        const Brick& b = m.bricks[ib];
        if (b.kind == HBO_BRICK_GENERIC_LAND_COVER || b.kind == HBO_BRICK_GLACIER) total += b.areaFraction;
    }
    if (!NearlyEqual(total, 1.0, EPSILON_D)) {
        double diff = total - 1.0;
        Brick& ground = m.bricks[genericId];
        if (ground.areaFraction < diff - EPSILON_D) throw Fail{"generic land cover not large enough"};
        if (ground.areaFraction - diff > 0) {
            SetAreaFraction(m, ground, ground.areaFraction - diff);
        } else {
            SetAreaFraction(m, ground, 0);
        }
    }
}

// Action::CheckLandCoverAreaFraction (Action.cpp:71-91)
double CheckFraction(double fraction) {
    if (NearlyZero(fraction, PRECISION)) return 0.0;
    if (NearlyEqual(1, fraction, PRECISION)) return 1.0;
    if (fraction < 0 || fraction > 1) throw Fail{"The fraction is not in the range [0 .. 1]"};
    return fraction;
}

// ActionGlacierEvolutionDeltaH::Apply (ActionGlacierEvolutionDeltaH.cpp:79-142)
void ApplyDeltaH(hbo_model& m) {
    const int n = static_cast<int>(m.dhBricks.size());
    double glacierWE = 0.0;
    for (int i = 0; i < n; ++i) {
        Brick& b = m.bricks[m.dhBricks[i]];
        if (NearlyZero(b.areaFraction, PRECISION)) continue;
        double area = b.areaFraction * m.dhUnitAreas[i];
        double brickGlacierWE = m.containers[b.second].content;
        glacierWE += area * brickGlacierWE;
    }
    double glacierRetreatPc = (m.dhInitialWE - glacierWE) / m.dhInitialWE;
    glacierRetreatPc = std::max(0.0, glacierRetreatPc);
    int nRows = m.dhRows;
    double rowPcIncrement = 1.0 / (nRows - 1);
    int row = static_cast<int>(glacierRetreatPc / rowPcIncrement);
    if (row != m.dhLastRow) {
        for (int i = 0; i < n; ++i) {
            double fraction = m.dhAreas[static_cast<size_t>(row) * n + i] / m.dhUnitAreas[i];
            fraction = CheckFraction(fraction);
            ChangeLandCoverAreaFraction(m, m.dhBricks[i], fraction);
        }
        m.dhLastRow = row;
    }
    double rowVolumeSum = 0;
    for (int i = 0; i < n; ++i) rowVolumeSum += m.dhVolumes[static_cast<size_t>(row) * n + i];
    for (int i = 0; i < n; ++i) {
        Container& ice = m.containers[m.bricks[m.dhBricks[i]].second];
        double areaGlacier = m.dhAreas[static_cast<size_t>(row) * n + i];
        if (NearlyZero(areaGlacier, PRECISION)) {
            ice.content = 0;
        } else {
            double volumeRatio = m.dhVolumes[static_cast<size_t>(row) * n + i] / rowVolumeSum;
            double iceWE = glacierWE * volumeRatio / areaGlacier;
            ice.content = iceWE;
        }
    }
}

// ActionGlacierEvolutionAreaScaling::Apply (ActionGlacierEvolutionAreaScaling.cpp:75-123): every unit of the lookup
// table follows its own column (retreat of ITS ice volume -> row -> new area; the ice w.e. is rescaled to the new
// area). Tables and Init are those of the delta-h action (…AreaScaling.cpp:9-68 = …DeltaH.cpp:9-70).
void ApplyAreaScaling(hbo_model& m) {
    const int n = static_cast<int>(m.dhBricks.size());
    int nRows = m.dhRows;
    double rowPcIncrement = 1.0 / (nRows - 1);
    for (int i = 0; i < n; ++i) {
        Brick& b = m.bricks[m.dhBricks[i]];
        if (NearlyZero(b.areaFraction, PRECISION)) continue;
        Container& ice = m.containers[b.second];
        double area = b.areaFraction * m.dhUnitAreas[i];
        double brickIceWE = ice.content;
        double iceVolume = area * brickIceWE;
        if (NearlyZero(iceVolume, PRECISION)) {
            ChangeLandCoverAreaFraction(m, m.dhBricks[i], 0);
            continue;
        }
        double initialWE = m.dhVolumes[i] * 900.0;  // _tableVolume.row(0) * constants::iceDensity
        double glacierRetreatPc = (initialWE - iceVolume) / initialWE;
        glacierRetreatPc = std::max(0.0, glacierRetreatPc);
        int row = static_cast<int>(glacierRetreatPc / rowPcIncrement);
        double newArea = m.dhAreas[static_cast<size_t>(row) * n + i];
        double newFraction = newArea / m.dhUnitAreas[i];
        newFraction = CheckFraction(newFraction);
        ChangeLandCoverAreaFraction(m, m.dhBricks[i], newFraction);
        double iceWE = iceVolume / newArea;  // +inf when the table's area is 0 (as the reference)
        ice.content = iceWE;
    }
}

// ActionGlacierSnowToIceTransformation::Apply (ActionGlacierSnowToIceTransformation.cpp:41-74)
void ApplySnowToIce(hbo_model& m) {
    for (auto& pr : m.snowToIce) {
        Brick& glacier = m.bricks[pr.first];
        if (NearlyZero(glacier.areaFraction, PRECISION)) continue;
        Container& snow = m.containers[m.bricks[pr.second].second];
        double snowWE = snow.content;
        if (snowWE <= 0) continue;
        Container& ice = m.containers[glacier.second];
        if (ice.infinite) throw Fail{"Trying to set the content of an infinite storage."};
        ice.content = ice.content + snowWE;
        snow.content = 0;
    }
}

void ApplyEvents(hbo_model& m, int step) {
    for (const Event& e : m.events) {
        if (e.step != step) continue;
        switch (e.kind) {
            case HBO_ACT_LAND_COVER_CHANGE: ChangeLandCoverAreaFraction(m, e.brick, e.value); break;
            case HBO_ACT_SNOW_TO_ICE: ApplySnowToIce(m); break;
            case HBO_ACT_DELTA_H: ApplyDeltaH(m); break;
            case HBO_ACT_AREA_SCALING: ApplyAreaScaling(m); break;
        }
    }
}

double RecordValue(const hbo_model& m, const Record& r) {
    switch (r.kind) {
        case HBO_REC_CONTAINER: return m.containers[r.id].content;
        case HBO_REC_FLUX: return m.fluxes[r.id].amount;
        case HBO_REC_FRACTION: return m.bricks[r.id].areaFraction;
        default: return m.outletTotal;
    }
}

}  // namespace

extern "C" {

hbo_model* hbo_create(int n_units, int n_steps, double dt_days, int solver, int spinup_steps) {
    auto* m = new hbo_model();
    m->nUnits = n_units;
    m->nSteps = n_steps;
    m->dt = dt_days;
    m->solver = solver;
    m->spinupSteps = std::min(spinup_steps, n_steps);  // ModelHydro.cpp:52-53
    m->unitBricks.resize(n_units);
    m->unitSplitters.resize(n_units);
    return m;
}

void hbo_destroy(hbo_model* m) { delete m; }

int hbo_add_brick(hbo_model* m, int kind, int unit, int needs_solver, int parent_brick) {
    Brick b;
    b.kind = kind;
    b.unit = unit;
    b.needsSolver = needs_solver != 0;
    b.parent = parent_brick;
    m->bricks.push_back(b);
    int id = static_cast<int>(m->bricks.size()) - 1;
    if (unit >= 0) {
        m->unitBricks[unit].push_back(id);
    } else {
        m->basinBricks.push_back(id);
    }
    return id;
}

void hbo_set_area_fraction(hbo_model* m, int brick, double fraction) {
    m->bricks[brick].areaFraction = fraction;
    m->bricks[brick].initialAreaFraction = fraction;
}
void hbo_add_event(hbo_model* m, int step, int kind, int brick, double value) {
    m->events.push_back(Event{step, kind, brick, value});
}
void hbo_add_snow_to_ice_unit(hbo_model* m, int glacier_brick, int snowpack_brick) {
    m->snowToIce.push_back({glacier_brick, snowpack_brick});
}
void hbo_set_delta_h(hbo_model* m, int n_units, const int* glacier_bricks, const double* unit_areas, int n_rows,
                     const double* areas, const double* volumes) {
    m->dhBricks.assign(glacier_bricks, glacier_bricks + n_units);
    m->dhUnitAreas.assign(unit_areas, unit_areas + n_units);
    m->dhRows = n_rows;
    m->dhAreas.assign(areas, areas + static_cast<size_t>(n_rows) * n_units);
    m->dhVolumes.assign(volumes, volumes + static_cast<size_t>(n_rows) * n_units);
    double initialVolume = 0;  // ActionGlacierEvolutionDeltaH.cpp:17-18
    for (int i = 0; i < n_units; ++i) initialVolume += volumes[i];
    m->dhInitialWE = initialVolume * 900.0;
}
void hbo_set_cover_links(hbo_model* m, int cover_brick, int generic_cover_brick, int snowpack_brick) {
    m->bricks[cover_brick].genericCover = generic_cover_brick;
    m->bricks[cover_brick].snowpack = snowpack_brick;
}
void hbo_set_parent(hbo_model* m, int brick, int parent_brick) { m->bricks[brick].parent = parent_brick; }
void hbo_set_target_brick(hbo_model* m, int process, int target_brick) { m->processes[process].targetBrick = target_brick; }
void hbo_add_capillary_target(hbo_model* m, int process, int target_brick, int n_weights, const int* weight_bricks) {
    m->processes[process].capTargets.push_back(target_brick);
    m->processes[process].capWeights.emplace_back(weight_bricks, weight_bricks + n_weights);
}
// ProcessLateralSnowSlide::SetHydroUnitProperties (slope in degrees, float) + ProcessLateral::AttachFluxOutWithWeight:
// call once per output, in output order, after hbo_process_add_output.
void hbo_set_lateral_slope(hbo_model* m, int process, float slope_deg) { m->processes[process].slopeDeg = slope_deg; }
void hbo_add_lateral_output(hbo_model* m, int process, int target_snowpack_brick, double weight) {
    m->processes[process].lateralTargets.push_back(target_snowpack_brick);
    m->processes[process].weights.push_back(weight);
}
void hbo_set_unit_area(hbo_model* m, int unit, double area) {
    if (static_cast<int>(m->unitAreas.size()) <= unit) m->unitAreas.resize(unit + 1, 0.0);
    m->unitAreas[unit] = area;
}

int hbo_add_container(hbo_model* m, int brick, int content_kind, float capacity, int infinite_storage,
                      int no_melt_when_snow_cover, double initial_state) {
    Container c;
    c.brick = brick;
    c.contentKind = content_kind;
    c.hasCapacity = !std::isnan(capacity);
    c.capacity = c.hasCapacity ? capacity : 0.0f;
    c.infinite = infinite_storage != 0;
    c.noMeltWhenSnowCover = no_melt_when_snow_cover != 0;
    c.initial = initial_state;
    m->containers.push_back(c);
    int id = static_cast<int>(m->containers.size()) - 1;
    if (content_kind == HBO_WATER) {
        m->bricks[brick].water = id;
    } else {
        m->bricks[brick].second = id;
    }
    return id;
}

void hbo_set_related_snowpack(hbo_model* m, int ice_container, int snowpack_brick) {
    m->containers[ice_container].relatedSnowpackBrick = snowpack_brick;
}

int hbo_add_process(hbo_model* m, int brick, int type, int container, const float* params, int n_params,
                    int target_brick, double unit_area, float slope) {
    Process p;
    p.brick = brick;
    p.type = type;
    p.container = container;
    for (int i = 0; i < n_params && i < 8; ++i) p.params[i] = params[i];
    p.targetBrick = target_brick;
    p.unitArea = unit_area;
    p.slope = slope;
    m->processes.push_back(p);
    int id = static_cast<int>(m->processes.size()) - 1;
    m->bricks[brick].processes.push_back(id);
    if (type == HBO_P_OVERFLOW) m->containers[m->bricks[brick].water].overflowProcess = id;  // ModelBuilder.cpp:236-238
    return id;
}

int hbo_add_flux(hbo_model* m, int kind, int is_static, int needs_weighting, double fraction_unit_area,
                 int target_container, int forcing_var, int unit) {
    Flux f;
    f.kind = kind;
    f.isStatic = is_static != 0;
    f.needsWeighting = needs_weighting != 0;
    f.fractionUnitArea = fraction_unit_area;
    f.fractionTotal = f.fractionUnitArea * f.fractionLandCover;
    f.targetContainer = target_container;
    f.forcingVar = forcing_var;
    f.unit = unit;
    m->fluxes.push_back(f);
    return static_cast<int>(m->fluxes.size()) - 1;
}

void hbo_process_add_output(hbo_model* m, int process, int flux) { m->processes[process].outputs.push_back(flux); }
void hbo_container_add_input(hbo_model* m, int container, int flux) { m->containers[container].inputs.push_back(flux); }

int hbo_add_splitter(hbo_model* m, int unit, int type, const float* params, int n_params) {
    Splitter s;
    s.unit = unit;
    s.type = type;
    for (int i = 0; i < n_params && i < 4; ++i) s.params[i] = params[i];
    m->splitters.push_back(s);
    int id = static_cast<int>(m->splitters.size()) - 1;
    m->unitSplitters[unit].push_back(id);
    return id;
}

void hbo_splitter_add_output(hbo_model* m, int splitter, int flux) { m->splitters[splitter].outputs.push_back(flux); }
void hbo_splitter_add_input(hbo_model* m, int splitter, int flux) { m->splitters[splitter].inputs.push_back(flux); }
void hbo_add_outlet_flux(hbo_model* m, int flux) { m->outletFluxes.push_back(flux); }

void hbo_set_forcing(hbo_model* m, int var, const double* data) {
    m->forcingStore[var].assign(data, data + static_cast<size_t>(m->nSteps) * m->nUnits);
    m->forcing[var] = m->forcingStore[var].data();
}

int hbo_add_record(hbo_model* m, int kind, int id) {
    m->records.push_back(Record{kind, id});
    return static_cast<int>(m->records.size()) - 1;
}

// LandCover.cpp:27-37, SurfaceComponent.cpp:12-24
void hbo_finalize_graph(hbo_model* m) {
    for (Brick& b : m->bricks) {
        if (b.kind == HBO_BRICK_STORAGE) continue;
        // SurfaceComponent::SetAreaFraction is only reached for surface components listed in the basin settings
        // (SubBasin.cpp:84-94); the basins built here list land covers only, so a snowpack's weighted out-fluxes (the
        // lateral ones) keep the default land-cover fraction of 1.
        if (b.kind == HBO_BRICK_SNOWPACK) continue;
        double value = b.areaFraction;
        for (int ip : b.processes) {
            for (int io : m->processes[ip].outputs) {
                Flux& f = m->fluxes[io];
                if (!f.needsWeighting) continue;
                if (b.kind == HBO_BRICK_SNOWPACK) {
                    value *= m->bricks[b.parent].areaFraction;
                }
                f.fractionLandCover = value;
                f.fractionTotal = f.fractionUnitArea * f.fractionLandCover;
            }
        }
    }
}

int hbo_run(hbo_model* m, double* outlet, double* records) {
    try {
        ConnectToElementsToSolve(*m);
        m->changeRatesNoSolver.assign(m->directConnectionCount, 0.0);
        const int n = m->solvableConnectionCount;
        m->k1.assign(n, 0.0);
        m->k2.assign(n, 0.0);
        m->k3.assign(n, 0.0);
        m->k4.assign(n, 0.0);
        m->combined.assign(n, 0.0);
        m->state.assign(m->stateVariableChanges.size(), 0.0);
        // ModelHydro::Reset (ModelHydro.cpp:183-193): containers to their initial state, fluxes to 0.
        for (Container& c : m->containers) {
            c.content = c.initial;
            c.dyn = 0;
            c.stat = 0;
        }
        for (Flux& f : m->fluxes) {
            f.amount = 0;
            f.changeRate = nullptr;
        }
        m->petOverride.assign(m->nUnits, 0.0);
        m->petOverrideStep.assign(m->nUnits, -1);
        for (Process& p : m->processes) {  // ProcessRoutingGR4J::Reset (ProcessRoutingGR4J.cpp:33-42)
            if (p.type != HBO_P_ROUTING_GR4J && p.type != HBO_P_ROUTING_GR6J) continue;
            RecomputeUHGr4j(p);
            std::fill(p.stuh1.begin(), p.stuh1.end(), 0.0);
            std::fill(p.stuh2.begin(), p.stuh2.end(), 0.0);
            p.gr_r = p.gr_qr = p.gr_qd = p.gr_rexp = p.gr_qrexp = 0.0;
            if (p.type == HBO_P_ROUTING_GR6J) m->containers[p.container].allowNegative = true;  // ProcessRoutingGR6J.cpp:28
        }
        for (Process& p : m->processes) {  // ProcessRoutingHBV::Reset (ProcessRoutingHBV.cpp:43-49)
            if (p.type != HBO_P_ROUTING_HBV) continue;
            RecomputeUH(p);
            std::fill(p.stuh.begin(), p.stuh.end(), 0.0);
            p.previousContent = 0.0;
            p.processStorage = 0.0;
        }
        // LandCover::Reset (LandCover.cpp:13-19): extents back to the initial ones; ActionsManager::Reset
        for (Brick& b : m->bricks) {
            if (b.kind == HBO_BRICK_GENERIC_LAND_COVER || b.kind == HBO_BRICK_GLACIER) {
                SetAreaFraction(*m, b, b.initialAreaFraction);
            }
        }
        m->dhLastRow = 0;
        for (int v = 0; v < 4; ++v) {
            if (!m->forcing[v]) {
                m->forcingStore[v].assign(static_cast<size_t>(m->nSteps) * m->nUnits, 0.0);
                m->forcing[v] = m->forcingStore[v].data();
            }
        }
        // RunSpinup (ModelHydro.cpp:135-181): replay the first steps unlogged, then rewind the clock.
        m->cursor = 0;
        for (int i = 0; i < m->spinupSteps; ++i) {
            ProcessTimeStep(*m, m->dt);
            m->cursor++;
        }
        m->cursor = 0;
        const size_t nrec = m->records.size();
        // RewindAfterSpinup (ModelHydro.cpp:167-181): the extents go back to the declared ones
        if (m->spinupSteps > 0) {
            for (Brick& b : m->bricks) {
                if (b.kind == HBO_BRICK_GENERIC_LAND_COVER || b.kind == HBO_BRICK_GLACIER) {
                    SetAreaFraction(*m, b, b.initialAreaFraction);
                }
            }
        }
        for (int t = 0; t < m->nSteps; ++t) {
            if (t > 0) ApplyEvents(*m, t);  // TimeMachine::IncrementTime -> ActionsManager::DateUpdate
            ProcessTimeStep(*m, m->dt);
            if (outlet) outlet[t] = m->outletTotal;
            for (size_t r = 0; r < nrec; ++r) records[r * m->nSteps + t] = RecordValue(*m, m->records[r]);
            m->cursor++;
        }
    } catch (const Fail& f) {
        m->error = f.msg;
        return -1;
    }
    return 0;
}

void hbo_set_start_mjd(hbo_model* m, double mjd) { m->startMjd = mjd; }

const char* hbo_last_error(hbo_model* m) { return m->error.c_str(); }
long hbo_unconverged_sweeps(hbo_model* m) { return m->unconverged; }

}  // extern "C"
