#!/usr/bin/env python
"""kbench.py — kernel-only timing of one libhbk build on a reduced C2-shaped workload (no torch import: seconds per
variant on the GPU box). Used to rank compile-time variants of the hot kernel before a full bench.py run.

    HBK_LIB=variants/libhbk_x.so python tools/kbench.py [--sets 28416] [--timesteps 1461] [--reps 3] [--workload c2]

Prints one JSON line: kernel ms (min / median of reps, CUDA events inside hbk_run), HRU*timesteps/s, the number of
non-converged constraint sweeps and an md5 of the outlet (identical md5 = bit-identical results between variants).
"""
import argparse
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--sets", type=int, default=28416)  # 148 SMs x 6 blocks x 32 sets: one wave of the C2 kernel
    ap.add_argument("--units", type=int, default=None)
    ap.add_argument("--timesteps", type=int, default=1461)
    ap.add_argument("--solver", default="rk4")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default=os.environ.get("HBK_LIB", "default"))
    # experiments on glacier workloads: no unit carries ice (the glacier kernel on plain units) / the plain structure
    ap.add_argument("--no-ice", action="store_true", help="glacier structure, glacier fraction 0 on every unit")
    ap.add_argument("--plain", action="store_true", help="same units and soil stores without the glacier cover")
    args = ap.parse_args()

    import bench
    from hydrobricks_b200 import models

    wl = bench.WORKLOADS[args.workload]
    P, U, T = args.sets, args.units or wl["units"], args.timesteps
    glacier = wl["glacier"] and not args.plain
    units = bench.hydro_units(U, area=wl["area"], glacier=glacier)
    if args.no_ice and glacier:
        for u in units:
            u["land_covers"] = [("ground", 1.0), ("glacier", 0.0)]
    precip, temp, pet = bench.station_series(T)
    ps = bench.parameter_sets(P)
    if glacier:
        rng = np.random.Generator(np.random.MT19937(91))
        ps.update({"a_ice": ps["a_snow"] + rng.uniform(0.5, 6, P), "k_snow": rng.uniform(0.05, 0.6, P),
                   "k_ice": rng.uniform(0.05, 0.6, P)})
    if wl["soil"] == 2:
        rng = np.random.Generator(np.random.MT19937(92))
        ps.update({"percol": rng.uniform(0, 10, P), "k_slow_2": ps["k_slow_1"] * rng.uniform(0.05, 0.9, P)})
    covers = ["ground", "glacier"] if glacier else ["ground"]
    m = models.Socont(solver=args.solver, soil_storage_nb=wl["soil"], surface_runoff="socont_runoff",
                      land_cover_names=covers, land_cover_types=covers,
                      glacier_infinite_storage=not wl.get("finite_ice", False), backend="python")
    end = str(np.datetime64(bench.START) + np.timedelta64(T - 1, "D"))
    m.setup(units, bench.START, end, device=0)
    matrix = m.parameter_matrix(ps, P)
    eng = m.engine(P)
    eng.set_parameters(matrix)
    S = bench.SPATIAL
    eng.set_station_forcing("precipitation", precip, S["zref"], S["grad_p"], 1.0)
    eng.set_station_forcing("temperature", temp, S["zref"], S["grad_t"], 1.0)
    eng.set_station_forcing("pet", pet)
    ms = []
    st = {}
    for _ in range(max(args.reps, 0) + 1):
        eng.run()
        st = eng.last_run_stats()
        ms.append(st["kernel_ms"])
    ms = ms[1:] or ms
    q = eng.get_outlet()
    line = {"tag": args.tag, "workload": args.workload, "solver": args.solver, "P": P, "U": U, "T": T,
            "kernel_ms_min": min(ms), "kernel_ms_median": float(np.median(ms)),
            "hru_steps_per_s": P * U * T / (min(ms) * 1e-3), "unconverged": st.get("unconverged_sweeps"),
            "outlet_md5": hashlib.md5(np.ascontiguousarray(q).tobytes()).hexdigest(),
            "outlet_sum": float(q.sum())}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
