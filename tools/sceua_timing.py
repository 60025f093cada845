"""Wall time of the batched SCE-UA (hydrobricks_b200/trainer.py BatchedSceUa) on one GPU for a few complex counts: the
launch of an evolution step costs about the same for 16 complexes as for 512, which is the point of evolving them side by
side. Synthetic basin (bench.py's generators), observations = the outlet of a known parameter set.

    python tools/sceua_timing.py [--units 50] [--steps 3653] > profiles/<round>_sceua_timing.jsonl
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import bench
    from hydrobricks_b200 import models, trainer

    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3653)
    ap.add_argument("--ngs", type=int, nargs="+", default=[16, 128, 512])
    a = ap.parse_args()
    U, T, warm = a.units, a.steps, 365
    units = bench.hydro_units(U)
    precip, temp, pet = bench.station_series(T)
    end = str(np.datetime64(bench.START) + np.timedelta64(T - 1, "D"))
    m = models.Socont(solver="rk4", soil_storage_nb=1, surface_runoff="socont_runoff", land_cover_names=["ground"],
                      land_cover_types=["ground"], backend="python")
    m.setup(units, bench.START, end)
    sp = bench.SPATIAL
    m.model.set_station_forcing("precipitation", precip, sp["zref"], sp["grad_p"], 1.0)
    m.model.set_station_forcing("temperature", temp, sp["zref"], sp["grad_t"], 1.0)
    m.model.set_station_forcing("pet", pet)
    truth = {"a_snow": 5.0, "A": 400.0, "k_slow_1": 0.05, "beta": 2000.0}
    m.run_sets(m.parameter_matrix({k: [v] for k, v in truth.items()}, 1))
    obs = m.model.get_outlet_discharge_batch()[0].copy()
    space = trainer.ParameterSpace.for_socont(list(truth), transforms=None)
    for ngs in a.ngs:
        sce = trainer.BatchedSceUa(m, space, obs, warmup=warm, objective="nse")
        t0 = time.perf_counter()
        best, score = sce.run(ngs * 250, ngs=ngs, seed=3)
        dt = time.perf_counter() - t0
        print(json.dumps({"ngs": ngs, "units": U, "timesteps": T, "solver": "rk4", "seconds": round(dt, 3),
                          "launches": sce.launches, "ms_per_launch": round(1e3 * dt / sce.launches, 3),
                          "sets_run": int(len(sce.scores)), "evaluations_counted": sce.evaluations, "loops": sce.loops,
                          "stop": sce.stop_reason, "nse": score, "one_minus_nse": 1.0 - score,
                          "sets_per_s": round(len(sce.scores) / dt, 1), "best": best}), flush=True)


if __name__ == "__main__":
    main()
