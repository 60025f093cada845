#!/usr/bin/env python
"""Algorithmic FP64 work per HRU*timestep of the hot path, counted — not estimated — in an instrumented
build of the CPU restatement (oracle/flop_count.h: every `double` of oracle/hb_oracle.cpp counts its own additions,
multiplications, divisions, comparisons, |x|, pow/sqrt and exp calls; results stay bit-identical to the plain oracle).

    python tools/count_flops.py            -> profiles/flops_per_hru_step.json

Workloads: the bench's synthetic forcing and parameter distribution (bench.py), 8 parameter sets x 20 units x 1 461
steps per (solver, structure). `flops` = add + mul + div + pow (a pow / sqrt call counts ONE operation, as the
reference's source writes it); comparisons and |x| are listed beside it. This is the figure bench.py multiplies by the
HRU*timesteps of a launch for `roofline.achieved`.
"""
import ctypes
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
SOLVERS = ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson"]
STRUCTURES = {"socont_1store": dict(glacier=False, soil=1), "gsm_socont_2store_glacier": dict(glacier=True, soil=2)}
P, U, T = 8, 20, 1461


def run_case(solver, st, count):
    import bench
    from hydrobricks_b200 import models
    from oracle import hb_oracle

    units = bench.hydro_units(U, glacier=st["glacier"])
    precip, temp, pet = bench.station_series(T)
    ps = bench.parameter_sets(P)
    if st["glacier"]:
        rng = np.random.Generator(np.random.MT19937(91))
        ps.update({"a_ice": ps["a_snow"] + rng.uniform(0.5, 6, P), "k_snow": rng.uniform(0.05, 0.6, P),
                   "k_ice": rng.uniform(0.05, 0.6, P)})
    if st["soil"] == 2:
        rng = np.random.Generator(np.random.MT19937(92))
        ps.update({"percol": rng.uniform(0, 10, P), "k_slow_2": ps["k_slow_1"] * rng.uniform(0.05, 0.9, P)})
    covers = ["ground", "glacier"] if st["glacier"] else ["ground"]
    m = models.Socont(solver=solver, soil_storage_nb=st["soil"], surface_runoff="socont_runoff", land_cover_names=covers,
                      land_cover_types=covers, backend="python")
    m.units = sorted(units, key=lambda u: -u["elevation"])
    forcing = bench.spatialise(m.units, precip, temp, pet)
    L = hb_oracle.lib()
    if count:
        L.hbo_flop_counts.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    totals = np.zeros(7, dtype=np.uint64)
    outs = []
    for k in range(P):
        for alias in ps:
            comp, name = m.aliases[alias]
            m.settings.set_parameter_value(comp, name, float(ps[alias][k]))
        om = hb_oracle.OracleModel(m.settings.get_structure(), m.units, solver, T)
        for name, v in zip(("precipitation", "temperature", "pet"), forcing):
            om.set_forcing(name, v)
        buf = (ctypes.c_ulonglong * 7)()
        if count:
            L.hbo_flop_counts(buf, 1)  # reset: graph building is not on the path
        outs.append(om.run())
        if count:
            L.hbo_flop_counts(buf, 1)
            totals += np.array(list(buf), dtype=np.uint64)
    return totals, np.array(outs)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        solver, name, count = sys.argv[2], sys.argv[3], sys.argv[4] == "1"
        totals, outs = run_case(solver, STRUCTURES[name], count)
        np.save(sys.argv[5], outs)
        print(json.dumps([int(x) for x in totals]))
        return
    subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "count", "oracle"], stdout=subprocess.DEVNULL)
    result = {"method": "instrumented CPU restatement (oracle/flop_count.h), bench.py forcing and parameter distribution, "
                        f"{P} sets x {U} units x {T} steps per case; flops = add + mul + div + pow (pow/sqrt = 1)",
              "per_hru_step": {}}
    for solver in SOLVERS:
        for name in STRUCTURES:
            res = []
            for count in (1, 0):
                env = dict(os.environ)
                if count:
                    env["HBO_LIB"] = str(ROOT / "oracle" / "libhboracle_count.so")
                tmp = f"/tmp/count_flops_{solver}_{name}_{count}.npy"
                out = subprocess.check_output([sys.executable, __file__, "--child", solver, name, str(count), tmp], env=env)
                res.append((json.loads(out.decode().strip().splitlines()[-1]), np.load(tmp)))
            assert np.array_equal(res[0][1], res[1][1]), "the instrumented build changed a result"
            n = P * U * T
            add, mul, div, cmp_, abs_, pw, ex = (x / n for x in res[0][0])
            result["per_hru_step"].setdefault(solver, {})[name] = {
                "add": round(add, 2), "mul": round(mul, 2), "div": round(div, 2), "pow_sqrt": round(pw, 2),
                "exp": round(ex, 3), "compare": round(cmp_, 2), "abs": round(abs_, 2),
                "flops": round(add + mul + div + pw + ex, 2)}
            print(solver, name, result["per_hru_step"][solver][name], flush=True)
    (ROOT / "profiles" / "flops_per_hru_step.json").write_text(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
