#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: Socont RK4 calibration ensemble (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            this repo's engine (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation of the path
                                                             (oracle/_ref = unmodified core compiled here, else the
                                                             oracle port) on all host cores, rank 0 only

A "step" = one pass of the hot path over one batch: ModelHydro::Run() for P parameter sets x U hydro units x T
daily steps (engine: one hbk_run). metric = HRU*timesteps/s (whole job, all ranks); evals_per_s = parameter
sets/s. Workload: P=100000 random parameter sets in the reference ranges, U=50 elevation
bands with slopes, T=10957 (1981-2010), Socont 1 soil store + runoff:socont, station forcing with elevation
gradients. Scaling is weak: every rank integrates its own P sets (parameter axis sharded, no communication on
the hot path; one NCCL all_gather of per-set scores per step).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

U_DEFAULT, T_DEFAULT, P_DEFAULT = 50, 10957, 100000
START, END = "1981-01-01", "2010-12-31"
# ALGORITHMIC FP64 work per HRU*timestep: counted in an instrumented build of the CPU restatement (tools/count_flops.py,
# oracle/flop_count.h) — additions + multiplications + divisions + pow/sqrt calls (one operation each, as
# the reference's source writes them) per solver and structure; comparisons are listed in the file, not counted.
def algorithmic_flops(solver, glacier, soil):
    try:
        table = json.load(open(ROOT / "profiles" / "flops_per_hru_step.json"))["per_hru_step"]
        name = "gsm_socont_2store_glacier" if (glacier or soil == 2) else "socont_1store"
        return table[solver][name]["flops"], table[solver][name]
    except Exception:
        return None, None


FP64_FLOPS_SOURCE = ("profiles/flops_per_hru_step.json: FP64 add + mul + div + pow/sqrt operations per HRU*timestep of "
                     "the reference's algorithm, counted in the instrumented CPU restatement (tools/count_flops.py); the "
                     "reference forbids FMA contraction, so its adds and multiplies cannot pair into the DFMAs the peak "
                     "is measured with")


def kernel_profile():
    """Counters of the current hot kernel from its committed ncu capture (tools/ncu_profile_json.py)."""
    try:
        return json.load(open(ROOT / "profiles" / "kernel_profile.json"))
    except Exception:
        return None

# ---------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------
def station_series(T, seed=42):
    rng = np.random.Generator(np.random.MT19937(seed))
    t = np.arange(T, dtype=np.float64)
    doy = np.fmod(t, 365.25) / 365.25
    u1, u2, u3 = rng.random(T), rng.random(T), rng.random(T)
    temp = 5 - 12 * np.cos(2 * np.pi * doy) + 4 * (u1 - 0.5)
    precip = np.where(u2 < 0.4, -8 * np.log(1 - u3), 0.0)
    pet = np.maximum(0.0, 2 - 2 * np.cos(2 * np.pi * doy))
    return precip, temp, pet


def hydro_units(U, seed=43, area=1.0e6, glacier=False):
    rng = np.random.Generator(np.random.MT19937(seed))
    step = 2400.0 / U
    elev = 800 + step * (np.arange(U) + 0.5)
    slope = rng.uniform(5, 35, U)
    units = []
    for i in range(U):
        covers = [("ground", 1.0)]
        if glacier:  # 0 below 2400 m, ramp to 0.8 at 3200 m
            g = float(np.clip((elev[i] - 2400) / 800 * 0.8, 0, 0.8))
            covers = [("ground", 1.0 - g), ("glacier", g)]
        units.append({"id": i + 1, "area": area, "elevation": float(elev[i]), "slope": (float(slope[i]), "deg"),
                      "land_covers": covers})
    return units


WORKLOADS = {
    # BASELINE.json configs[1]: the headline
    "c2": dict(sets=100000, units=50, timesteps=10957, glacier=False, soil=1, area=1.0e6,
               name="C2 Socont-RK4 calibration ensemble"),
    # configs[2]: 1M HRUs (100 m grid) x 40 years, snow + glacier + soil + groundwater bricks, one parameter set
    "c3": dict(sets=1, units=1000000, timesteps=14610, glacier=True, soil=2, area=1.0e4,
               name="C3 synthetic distributed basin, 1M HRUs x 40 y, GSM-Socont"),
    # configs[3]: glacierised Socont with delta-h glacier evolution and land-cover-change actions over 50 years: the
    # run is cut into kernel launches at the action dates (49 October updates + 50 dated land-cover edits)
    "c4": dict(sets=20000, units=50, timesteps=18262, glacier=True, soil=2, area=1.0e6, finite_ice=True, actions=True,
               name="C4 glacierised Socont, delta-h evolution + land-cover changes, 50 y"),
    # configs[4]: solver sweep, 10k sets x 1k HRUs x 10 years
    "c5": dict(sets=10000, units=1000, timesteps=3653, glacier=False, soil=1, area=1.0e6,
               name="C5 solver sweep, 10k sets x 1k HRUs x 10 y"),
}


def block_shape_text(shape):
    """The launch shape of the timed runs in words, from hbk_last_run_shape."""
    sp, ub = shape["sets_per_block"], shape["units_per_block"]
    lanes = "parameter sets (lanes)" if sp == 32 else "parameter set(s)"
    text = f"{sp} {lanes} x {ub} units per block, {shape['unit_tiles']} unit tiles, {shape['tiles_per_block']} tiles " \
           f"time-multiplexed per block, {shape['steps_per_chunk']} steps per chunk"
    if shape["tile_groups"] > 1:
        text += f"; {shape['tile_groups']} tile groups write partial outlet series, summed in group order (hbkReduceTiles)"
    if shape["launches_per_segment"] > 1:
        text += f"; {shape['launches_per_segment']} unit-kernel launches per run segment: the {shape['plain_variant_tiles']} " \
                "tiles without glacier units run the plain kernel variant"
    if sp == 32:
        text += "; parameter sets ordered inside the engine so that warps hold similar sets"
    return text


def parameter_sets(P, seed=44):
    """Uniform in the reference ranges (parameters.py:66-174, 736-738), transition_start < transition_end."""
    rng = np.random.Generator(np.random.MT19937(seed))
    t0 = rng.uniform(-2, 2, P)
    t1 = rng.uniform(0, 4, P)
    bad = t0 >= t1
    while bad.any():
        t0[bad], t1[bad] = rng.uniform(-2, 2, bad.sum()), rng.uniform(0, 4, bad.sum())
        bad = t0 >= t1
    return {"a_snow": rng.uniform(2, 20, P), "A": rng.uniform(0, 3000, P), "k_slow_1": rng.uniform(1e-4, 1, P),
            "beta": rng.uniform(100, 30000, P), "prec_t_start": t0, "prec_t_end": t1}


SPATIAL = {"zref": 1250.0, "grad_t": -0.6, "grad_p": 0.05}


def mjd_axis(T):
    return np.arange(T, dtype=np.float64) + 44605.0  # 1981-01-01


def c4_tables(units, rows=11):
    """Synthetic delta-h lookup tables over the glacier units of `units` (basin order): 11 rows,
    the lower units vanish first, 40..160 m of ice. Returns (unit positions, areas[rows x n] m2, volumes m3)."""
    gl = [i for i, u in enumerate(units) if dict(u["land_covers"]).get("glacier", 0.0) > 0]
    a0 = np.array([units[i]["area"] * dict(units[i]["land_covers"])["glacier"] for i in gl])
    z = np.array([units[i]["elevation"] for i in gl])
    rank = (z - z.min()) / max(z.max() - z.min(), 1.0)
    thick0 = 40.0 + 120.0 * rank
    areas = np.zeros((rows, len(gl)))
    volumes = np.zeros((rows, len(gl)))
    for r in range(rows):
        x = r / (rows - 1)
        keep = np.clip((1 + 0.6 * rank) * (1 - x) - 0.45 * (1 - rank) * x, 0, 1)
        if r == rows - 1:
            keep[:] = 0
        areas[r] = a0 * keep
        volumes[r] = areas[r] * thick0 * (1 - 0.7 * x)
    areas[0] = a0
    volumes[0] = a0 * thick0
    return gl, areas, volumes


def c4_changes(units, gl, areas, T):
    """One dated land-cover edit per year (15 January): the glacier of one unit is set to a share of its initial area."""
    out = []
    for year in range(1, T // 365 + 1):
        day = 365 * year + 14
        if day >= T:
            break
        j = year % len(gl)
        out.append((mjd_axis(1)[0] + day, units[gl[j]]["id"], float(areas[0][j] * max(0.2, 1.0 - 0.012 * year))))
    return out


def add_c4_actions(m, T):
    """Registers the C4 actions on a set-up model through the reference's action classes; returns them (the model
    holds non-owning pointers, ModelHydro::AddAction)."""
    hb = m.backend
    gl, areas, volumes = c4_tables(m.units)
    dh = hb.ActionGlacierEvolutionDeltaH()
    dh.add_lookup_tables(10, "glacier", np.array([m.units[i]["id"] for i in gl], dtype=np.int32), areas, volumes)
    assert m.model.add_action(dh)
    lc = hb.ActionLandCoverChange()
    for date, unit_id, area in c4_changes(m.units, gl, areas, T):
        lc.add_change(date, unit_id, "glacier", area)
    assert m.model.add_action(lc)
    return [dh, lc]


def spatialise(units, precip, temp, pet):
    """forcing.py:1061-1113 in numpy: the pre-spatialised [T x U] arrays the reference core is given."""
    z = np.array([u["elevation"] for u in units])
    t2 = temp[:, None] + SPATIAL["grad_t"] * (z - SPATIAL["zref"]) / 100
    p2 = precip[:, None] * (1 + SPATIAL["grad_p"] * (z - SPATIAL["zref"]) / 100)
    p2[p2 < 0] = 0
    e2 = np.repeat(pet[:, None], len(z), axis=1)
    return p2, t2, e2




# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on all host cores
# ---------------------------------------------------------------------------------------------------------
_W = {}
# parameter sets per host core for the cpu_baseline sample (>= 64 per core for the ensembles; C5 has 20 x
# the units per set, C3 is one parameter set per core on a 1 000-unit basin)
REF_SETS_PER_CORE = {"c2": 64, "c4": 32, "c5": 4, "c3": 1}
REF_UNITS = {"c3": 1000}  # the reference is timed on U = 1 000 (and once on 10 000) for the basin config


def _load_ref_backend():
    cands = sorted((ROOT / "oracle" / "_ref").glob("_hydrobricks*.so"))
    if not cands:
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location("_hydrobricks", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_parameter_sets(wl, P, seed=44, extra_seed=0):
    ps = parameter_sets(P, seed=seed)
    if wl["glacier"]:
        rng = np.random.Generator(np.random.MT19937(91 + extra_seed))
        ps.update({"a_ice": ps["a_snow"] + rng.uniform(0.5, 6, P), "k_snow": rng.uniform(0.05, 0.6, P),
                   "k_ice": rng.uniform(0.05, 0.6, P)})
    if wl["soil"] == 2:
        rng = np.random.Generator(np.random.MT19937(92 + extra_seed))
        ps.update({"percol": rng.uniform(0, 10, P), "k_slow_2": ps["k_slow_1"] * rng.uniform(0.05, 0.9, P)})
    return ps


def _ref_worker_init(workload, solver, U, T, log_path):
    from hydrobricks_b200 import models

    # the reference logs "did not stabilize after 10 sweeps" on stderr: keep it in a file per worker and count it
    fd = os.open(f"{log_path}.{os.getpid()}", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    os.dup2(fd, 2)
    wl = WORKLOADS[workload]
    ref = _load_ref_backend()
    units = hydro_units(U, area=wl["area"], glacier=wl["glacier"])
    precip, temp, pet = station_series(T)
    covers = ["ground", "glacier"] if wl["glacier"] else ["ground"]
    kw = dict(solver=solver, soil_storage_nb=wl["soil"], surface_runoff="socont_runoff", land_cover_names=covers,
              land_cover_types=covers, glacier_infinite_storage=not wl.get("finite_ice", False))
    end = END if T == T_DEFAULT else str(np.datetime64(START) + np.timedelta64(T - 1, "D"))
    if ref is not None:
        ref.init()
        m = models.Socont(backend=ref, **kw)
        m.setup(units, START, end)
        p2, t2, e2 = spatialise(m.units, precip, temp, pet)
        m.set_forcing(mjd_axis(T), p2, t2, e2)
        _W["actions"] = add_c4_actions(m, T) if wl.get("actions") else []
        _W["kind"], _W["model"] = "reference", m
    else:  # oracle port (the restatement) when the compiled reference did not travel
        from hydrobricks_b200 import _hydrobricks as mirror
        from oracle import hb_oracle

        m = models.Socont(backend=mirror, **kw)
        m.units = sorted(units, key=lambda u: -u["elevation"])
        _W["kind"], _W["model"], _W["oracle"] = "port", m, hb_oracle
        _W["forcing"] = spatialise(m.units, precip, temp, pet)
    _W["T"], _W["U"], _W["solver"] = T, U, solver


def _ref_worker_run(sets):
    m, T = _W["model"], _W["T"]
    t_run = 0.0
    q = np.zeros(1)
    for values in sets:
        if _W["kind"] == "reference":
            m.set_parameters(values)
            m.model.reset()
            t0 = time.perf_counter()
            m.model.run()
            t_run += time.perf_counter() - t0
            q = m.model.get_outlet_discharge()
        else:
            for alias, v in values.items():
                comp, name = m.aliases[alias]
                m.settings.set_parameter_value(comp, name, v)
            om = _W["oracle"].OracleModel(m.settings.get_structure(), m.units, _W["solver"], T)
            for k, v in zip(("precipitation", "temperature", "pet"), _W["forcing"]):
                om.set_forcing(k, v)
            t0 = time.perf_counter()
            q = om.run()
            t_run += time.perf_counter() - t0
    return t_run, float(np.sum(q))


def run_reference_arm(args, quiet=False):
    import glob
    import math
    import multiprocessing as mp
    import tempfile

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    wl = WORKLOADS[args.workload]
    T = args.timesteps
    U = REF_UNITS.get(args.workload, args.units)
    # >= the workload's sets per core over the timed steps (64 for C2), spread over the steps so that the arm ends
    # within minutes whatever --steps the driver asks for
    target = REF_SETS_PER_CORE.get(args.workload, 64)
    n_per_core = args.ref_sets_per_core or max(1, math.ceil(target / max(args.steps, 1)))
    n_steps = args.steps + args.warmup
    ps = workload_parameter_sets(wl, cores * n_per_core * n_steps + 1)
    chunks = []
    k = 0
    for _ in range(n_steps):
        step = []
        for _c in range(cores):
            step.append([{a: float(ps[a][k + j]) for a in ps} for j in range(n_per_core)])
            k += n_per_core
        chunks.append(step)
    log_path = os.path.join(tempfile.mkdtemp(prefix="hbref_"), "stderr")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(args.workload, args.solver, U, T, log_path)) as pool:
        pool.map(_ref_worker_run, [[] for _ in range(cores)])  # make sure every worker is initialised
        times = []
        kind = "reference" if _load_ref_backend() is not None else "port"
        for i, step in enumerate(chunks):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_run, step, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append((dt, max(r[0] for r in res)))
    warnings = 0
    for f in glob.glob(log_path + ".*"):
        try:
            warnings += open(f, errors="replace").read().count("did not stabilize")
        except OSError:
            pass
    wall = sum(t[0] for t in times)
    n_sets = cores * n_per_core * len(times)
    value = n_sets * U * T / wall
    line = {
        "metric": "hru_timesteps_per_s", "value": value, "unit": "HRU*timesteps/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "evals_per_s": n_sets / wall,
        "config": {"workload": wl["name"] + " (bounded CPU sample)", "sets_per_step": cores * n_per_core,
                   "hydro_units": U, "timesteps": T, "solver": args.solver,
                   "structure": f"socont {wl['soil']} soil store(s), runoff:socont" + (", glacier (GSM)" if wl["glacier"] else ""),
                   "forcing": "pre-spatialised per-unit series (station + elevation gradients)"},
        "cpu_baseline": {"value": value, "unit": "HRU*timesteps/s", "cores": cores, "kind": kind,
                         "sample": f"{n_sets} parameter sets over {len(times)} timed step(s) ({n_per_core} per core and "
                                   f"step, {n_per_core * len(times)} per core in total) x {U} units x {T} steps, one model "
                                   f"instance per core, wall time incl. parameter update",
                         "did_not_stabilize_warnings": warnings,
                         "note": "the reference's own 'did not stabilize after 10 sweeps' log lines over the whole sample "
                                 "(incl. warm-up steps)"},
        "e2e": {"value": value, "unit": "HRU*timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not quiet:
        print(json.dumps(line), flush=True)
    return line


# ---------------------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class DeviceArray:
    """Zero-copy view of engine-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, shape, strides):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": strides}


def run_engine_arm(args):
    import torch
    import torch.distributed as dist

    from hydrobricks_b200 import engine as hbk_engine
    from hydrobricks_b200 import models, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (engine arm) needs a CUDA device; there is no CPU fallback")

    # cpu_baseline: rank 0, N=1 only, in a child process BEFORE this process touches CUDA: the reference's own CPU
    # implementation on all host cores, 64 parameter sets per core of the same workload
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
               "--workload", args.workload, "--solver", args.solver, "--timesteps", str(args.timesteps),
               "--ref-sets-per-core", str(args.ref_sets_per_core or REF_SETS_PER_CORE.get(args.workload, 64))]
        if args.workload in ("c2", "c4", "c5"):
            cmd += ["--units", str(args.units)]
        out = subprocess.run(cmd, capture_output=True, text=True)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                cpu_baseline = json.loads(ln)["cpu_baseline"]
        if cpu_baseline is None:
            cpu_baseline = {"value": None, "unit": "HRU*timesteps/s", "cores": 0, "kind": "port",
                            "sample": "failed: " + out.stderr[-300:]}

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    U, T = args.units, args.timesteps
    wl = WORKLOADS[args.workload]
    # C3 on several GPUs: ONE basin, every rank holds a block of its hydro units; per step one
    # all-reduce of the partial outlet and of the two deposit series, then the sub-basin stores.
    #   weak: a basin of world x U units (U per rank)      strong: the basin of U units split over the ranks
    unit_sharded = args.workload == "c3" and world > 1
    strong = args.scaling == "strong"
    if unit_sharded:
        P = args.sets
        U_total = U if strong else world * U
        # blocks of equal integration cost, not of equal unit count: the units with ice (the upper third, at one end of
        # the elevation-sorted basin) run the glacier kernel variant (parallel.GLACIER_UNIT_COST)
        all_units = hydro_units(U_total, area=wl["area"], glacier=wl["glacier"])
        u0, u1 = parallel.weighted_shard_bounds(parallel.unit_costs(all_units), rank, world)
        units = all_units[u0:u1]
        del all_units
        U = u1 - u0
        set_seed = 0  # unit shards integrate the SAME parameter sets
        ps = parameter_sets(P, seed=44)
        p0 = 0
    else:
        # ensembles: the parameter axis is sharded, no communication on the path.
        #   weak: every rank integrates args.sets sets of its own     strong: args.sets sets in total, split over the ranks
        units = hydro_units(U, area=wl["area"], glacier=wl["glacier"])
        U_total = U
        if strong:
            p0, p1 = parallel.shard_bounds(args.sets, rank, world)
            P = p1 - p0
            set_seed = 0
            ps = {k: v[p0:p1] for k, v in parameter_sets(args.sets, seed=44).items()}
        else:
            P, p0, set_seed = args.sets, 0, rank
            ps = parameter_sets(P, seed=44 + set_seed)
    P_total = args.sets if (strong or unit_sharded) else world * args.sets
    precip, temp, pet = station_series(T)
    n_draw = args.sets if strong and not unit_sharded else P
    if wl["glacier"]:
        rng = np.random.Generator(np.random.MT19937(91 + set_seed))
        extra = {"a_ice": rng.uniform(0.5, 6, n_draw), "k_snow": rng.uniform(0.05, 0.6, n_draw),
                 "k_ice": rng.uniform(0.05, 0.6, n_draw)}
        sl = slice(p0, p0 + P) if (strong and not unit_sharded) else slice(None)
        ps.update({"a_ice": ps["a_snow"] + extra["a_ice"][sl], "k_snow": extra["k_snow"][sl], "k_ice": extra["k_ice"][sl]})
    if wl["soil"] == 2:
        rng = np.random.Generator(np.random.MT19937(92 + set_seed))
        percol, ratio = rng.uniform(0, 10, n_draw), rng.uniform(0.05, 0.9, n_draw)
        sl = slice(p0, p0 + P) if (strong and not unit_sharded) else slice(None)
        ps.update({"percol": percol[sl], "k_slow_2": ps["k_slow_1"] * ratio[sl]})

    covers = ["ground", "glacier"] if wl["glacier"] else ["ground"]
    m = models.Socont(solver=args.solver, soil_storage_nb=wl["soil"], surface_runoff="socont_runoff",
                      land_cover_names=covers, land_cover_types=covers,
                      glacier_infinite_storage=not wl.get("finite_ice", False))
    end = END if T == T_DEFAULT else str(np.datetime64(START) + np.timedelta64(T - 1, "D"))
    m.setup(units, START, end, device=local)
    matrix = m.parameter_matrix(ps, P)
    actions = add_c4_actions(m, T) if wl.get("actions") else []
    eng = m.engine(P)
    if unit_sharded:
        eng.set_unit_shard(parallel.basin_area(np.full(U_total, wl["area"])))

    def run_engine(e):
        e.run()
        if unit_sharded:
            parallel.reduce_partials(parallel.shard_partials(e))
            torch.cuda.current_stream().synchronize()
            e.finish_sharded_run()

    # pinned host staging: the step's inputs and its result
    pin_params = torch.from_numpy(matrix).pin_memory()
    pin_forcing = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (precip, temp, pet)]
    host_out_pinned = True
    try:
        pin_out = torch.empty((P, T), dtype=torch.float64).pin_memory()
    except RuntimeError:  # page-locking P x T x 8 bytes per rank refused (8 ranks x 8.8 GB on one host): pageable
        pin_out = torch.empty((P, T), dtype=torch.float64)  # buffer, the e2e figure then includes a staged copy
        host_out_pinned = False
    out_ptr = ctypes.cast(pin_out.data_ptr(), ctypes.POINTER(ctypes.c_double))

    def upload(eng=eng):
        eng.set_parameters(pin_params.numpy())
        eng.set_station_forcing("precipitation", pin_forcing[0].numpy(), SPATIAL["zref"], SPATIAL["grad_p"], 1.0)
        eng.set_station_forcing("temperature", pin_forcing[1].numpy(), SPATIAL["zref"], SPATIAL["grad_t"], 1.0)
        eng.set_station_forcing("pet", pin_forcing[2].numpy())

    upload()
    if actions:
        # one run through the host's ModelHydro maps the action dates to time steps and registers them with the engine
        # as kernel boundaries (ActionsManager::DateUpdate); the timed eng.run() calls below reuse that schedule
        m.model.set_parameter_sets(matrix)
        m.model.set_station_forcing("precipitation", precip, SPATIAL["zref"], SPATIAL["grad_p"], 1.0)
        m.model.set_station_forcing("temperature", temp, SPATIAL["zref"], SPATIAL["grad_t"], 1.0)
        m.model.set_station_forcing("pet", pet)
        m.model.run()
    stream = torch.cuda.ExternalStream(eng.stream())
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    gather_scores = world > 1 and not unit_sharded
    if gather_scores:  # ranks may hold different numbers of sets (strong scaling): gather padded to the largest share
        p_max = max(parallel.shard_bounds(args.sets, r, world)[1] - parallel.shard_bounds(args.sets, r, world)[0]
                    for r in range(world)) if strong else P
        scores_pad = torch.zeros(p_max, dtype=torch.float64, device="cuda")
        scores_all = torch.empty(world * p_max, dtype=torch.float64, device="cuda")

    # synthetic observed discharge for the objective function (NSE of every parameter set, computed on the device)
    rng = np.random.default_rng(5)
    pin_obs = torch.from_numpy(np.abs(precip * 0.6 + rng.normal(0, 0.3, T)) + 0.1).pin_memory()
    pin_scores = torch.empty(P, dtype=torch.float64).pin_memory()
    score_warmup = min(365, T // 4)

    def scores():
        eng.compute_scores("nse", pin_obs.numpy(), score_warmup, to_host=False)
        if gather_scores:  # the only collective of the job: final gather of the per-set scores
            sc = torch.as_tensor(DeviceArray(eng.scores_device_pointer(), (P,), (8,)), device="cuda")
            scores_pad[:P].copy_(sc)
            dist.all_gather_into_tensor(scores_all, scores_pad)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        run_engine(eng)
        scores()

    peak_fp64 = hbk_engine.measure_fp64_peak(local)

    # ---- resident (value) ----
    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    kernel_ms, launches = [], 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (256 MiB write > 126 MB L2), outside the step's events
        torch.cuda.synchronize()
        ev[i][0].record(stream)
        resident_step()
        ev[i][1].record(stream)
        st = eng.last_run_stats()
        kernel_ms.append(st["kernel_ms"])
        launches += st["launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    unconverged_sweeps = eng.last_run_stats()["unconverged_sweeps"]
    shape = eng.last_run_shape()
    unconverged_sets = int((eng.unconverged_per_set() > 0).sum())
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, wall * 1e3, sum(kernel_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, kern_ms = t.tolist()
    units_done = P_total * U_total * T * args.steps
    value = units_done / (dev_ms * 1e-3)

    # ---- end to end: pinned host inputs -> device -> run -> full outlet back to pinned host, through the C-ABI. The
    # outlet of step i is fetched asynchronously (hbk_get_outlet_async, one strided copy on the engine's copy stream)
    # while step i+1 uploads its inputs and integrates into the engine's second outlet buffer; every step's outlet has
    # landed in host memory when the timed region ends (hbk_wait_outlet). ----
    def e2e_step():
        upload()
        run_engine(eng)
        eng._check(eng.L.hbk_wait_outlet(eng.h))  # the host buffer is free again (copy of the previous step done)
        eng._check(eng.L.hbk_get_outlet_async(eng.h, 0, P, out_ptr))

    for _ in range(min(args.warmup, 1)):
        e2e_step()
    eng.wait_outlet()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    eng.wait_outlet()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units_done / t.item()
    h2d = pin_params.numel() * 4 + sum(x.numel() * 8 for x in pin_forcing)
    d2h = P * T * 8

    # ---- the same with the blocking fetch (hbk_get_outlet): copy after compute, nothing overlaps ----
    def e2e_blocking_step():
        upload()
        run_engine(eng)
        eng._check(eng.L.hbk_get_outlet(eng.h, 0, P, out_ptr))

    e2e_blocking_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_blocking_step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_blocking_value = units_done / t.item()

    # ---- the calibration loop's variant: only the per-set objective scores come back ----
    def e2e_scores_step():
        upload()
        run_engine(eng)
        eng.compute_scores("nse", pin_obs.numpy(), score_warmup, out=pin_scores.numpy())

    e2e_scores_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_scores_step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_scores_value = units_done / t.item()

    if rank == 0:
        main_ms = kern_ms / args.steps  # device time of one hbk_run (state reset + unit kernel), CUDA events
        flops_unit, flops_detail = algorithmic_flops(args.solver, wl["glacier"], wl["soil"])
        achieved = flops_unit * P * U * T / (main_ms * 1e-3) / 1e12 if flops_unit else None
        hbm_bytes = P * T * 8 + 3 * T * 8 + P * hbk_engine.N_PARAMS * 4
        try:
            hbm_peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        prof = kernel_profile() or {}
        default_shape = args.workload == "c2" and args.solver == "rk4" and (args.sets, U, T) == (P_DEFAULT, U_DEFAULT, T_DEFAULT)
        traffic = prof.get("dram_bytes_full_c2_launch") if (default_shape and not strong) else None
        line = {
            "metric": "hru_timesteps_per_s", "value": value, "unit": "HRU*timesteps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "evals_per_s": P_total * args.steps / (dev_ms * 1e-3),
            "config": {"workload": wl["name"], "sets_per_gpu": P, "sets_total": P_total, "hydro_units": U_total,
                       "hydro_units_per_gpu": U,
                       "multi_gpu": ("one basin of %d units in contiguous blocks of equal integration cost (units with ice "
                                     "weigh %.2f; rank 0 holds %d); per step one all-reduce of the partial outlet and the "
                                     "two sub-basin deposit series [T], then the sub-basin stores"
                                     % (U_total, parallel.GLACIER_UNIT_COST, U)) if unit_sharded else
                                    "parameter axis sharded, no collective on the path, one all_gather of scores",
                       "timesteps": T, "solver": args.solver,
                       "structure": f"socont {wl['soil']} soil store(s), runoff:socont" + (", glacier (GSM)" if wl["glacier"] else ""),
                       "forcing": "station series + fused elevation-gradient spatialisation",
                       "l2": "256 MiB flush between timed iterations; per-step output %.3g GB" % (P * T * 8 / 1e9),
                       "block_shape": block_shape_text(shape),
                       "launch_shape": shape,
                       "arithmetic": "every quotient of the reference correctly rounded, verification sweep kept "
                                     "(bit-identical to the division-for-division build, DESIGN.md §4)"},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": "HRU*timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "note": "per step: pinned host params + station series -> device, run, full outlet[P x T] -> pinned "
                            "host through the C-ABI; the fetch is hbk_get_outlet_async, so the copy of step i overlaps "
                            "the run of step i+1; all outlets have landed when the clock stops",
                    "host_outlet_buffer_pinned": host_out_pinned,
                    "blocking_fetch": {"value": e2e_blocking_value,
                                       "note": "same bytes, hbk_get_outlet after every run (copy not overlapped)"},
                    "scores_only": {"value": e2e_scores_value, "d2h_bytes_per_step": int(P * 8),
                                    "note": "same, but the device computes NSE per set (hbk_compute_scores) and only "
                                            "the scores come back: the calibration loop's step; NSE follows the published "
                                            "formula, parity against HydroErr is unpinned (absent from the image)"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": {"unconverged_sweeps_last_run": int(unconverged_sweeps), "sets_with_unconverged_sweeps": unconverged_sets,
                       "sets_per_gpu": P,
                       "note": "constraint passes in which the reference logs 'did not stabilize after 10 sweeps' "
                               "(Processor.cpp:191-209, soil capacities below ~60 mm): there the reference's own result "
                               "hangs on the last bit (its -ffp-contract=fast build disagrees with itself on 130 of 207 "
                               "such sets, tests/golden/ensemble_fuzz_reference_fma.json); everywhere else 1e-10"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s",
                         "frac": achieved / peak_fp64 if (peak_fp64 and achieved) else None, "traffic": traffic,
                         "traffic_note": "DRAM bytes of the full-size C2 launch from ncu (profiles/kernel_profile.json, "
                                         "kernel version %s); algorithmic bytes are in roofline.hbm; the excess is the "
                                         "unit state of the time-multiplexed tiles passing through L2/HBM (<3 %% of HBM "
                                         "bandwidth)" % prof.get("version"),
                         "kernel": f"hbkRunUnits<{args.solver},{wl['soil']} soil store(s),{'glacier' if wl['glacier'] else 'no glacier'}>",
                         "kernel_ms": main_ms,
                         "flops_per_hru_step": flops_unit, "flops_detail": flops_detail, "flops_source": FP64_FLOPS_SOURCE,
                         "frac_of_non_fma_peak": 2 * achieved / peak_fp64 if (peak_fp64 and achieved) else None,
                         "flops_executed_per_hru_step": prof.get("flops_executed_per_hru_step") if default_shape else None,
                         "fp64_pipe_pct": prof.get("fp64_pipe_pct") if default_shape else None,
                         "issue_slots_pct": prof.get("issue_active_pct") if default_shape else None,
                         "profile_version": prof.get("version"),
                         "peak_source": "DFMA micro-benchmark measured in this run (hbk_measure_fp64_peak, builder-"
                                        "measured: MEASURED_PEAKS.json has no FP64 figure); 2 flop per DFMA",
                         "note": "not a dense contraction: neither HBM nor tensor binds, the FP64 pipe / instruction "
                                 "issue does (SURVEY.md §8d). frac counts the reference's algorithmic operations (one "
                                 "per add / mul / div / pow) against the DFMA peak (two per instruction): an algorithm "
                                 "whose adds and multiplies may not be contracted tops out at 0.5 "
                                 "(frac_of_non_fma_peak); fp64_pipe_pct / issue_slots_pct are the ncu counters of the "
                                 "committed profile of this kernel version",
                         "hbm": {"achieved": hbm_bytes / (main_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_bytes / (main_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                                 "algorithmic_bytes_per_launch": int(hbm_bytes)}},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --sets parameter sets (c3: --units hydro units) PER GPU; strong: in total, split over the GPUs")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--solver", default="rk4", choices=["rk4", "heun_explicit", "euler_explicit", "crank_nicolson",
                                                        "implicit_euler", "exponential_euler", "analytic_linear"])
    ap.add_argument("--sets", type=int, default=None, help="parameter sets per GPU")
    ap.add_argument("--units", type=int, default=None)
    ap.add_argument("--timesteps", type=int, default=None)
    ap.add_argument("--ref-sets-per-core", type=int, default=None,
                    help="reference arm / cpu_baseline: parameter sets per host core and step; default: >= 64 per core "
                         "over the timed steps (SURVEY.md §8d), i.e. ceil(64 / steps) per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    for key in ("sets", "units", "timesteps"):
        if getattr(args, key) is None:
            setattr(args, key, WORKLOADS[args.workload][key])
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference_arm(args)
    else:
        run_engine_arm(args)


if __name__ == "__main__":
    main()
