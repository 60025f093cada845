"""The two calibration GPU tests (tests/test_gpu_parity.py: test_batched_monte_carlo_calibration,
test_batched_sceua_calibration) replayed on the CPU with the ORACLE standing in for the engine (TEST INFRASTRUCTURE:
checks the tests' thresholds and the samplers' host logic when no GPU is at hand; the oracle agrees with the engine to
1e-10, so the draws, the trajectories and the scores are the GPU run's).

    python tools/calibration_tests_on_oracle.py [mc] [sceua]
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class OracleBackedSocont:
    def __init__(self, solver, units, forcing, n_steps):
        from hydrobricks_b200 import models

        self.m = models.Socont(solver=solver, soil_storage_nb=1, surface_runoff="socont_runoff", land_cover_names=["ground"],
                               land_cover_types=["ground"], backend="python")
        self.solver, self.units, self.forcing, self.T = solver, units, forcing, n_steps
        self.names = None

    def parameter_matrix(self, sets, n):
        self.names = list(sets)
        return np.column_stack([np.asarray(sets[k], dtype=np.float64) for k in self.names])

    def run_sets(self, matrix):
        from oracle import hb_oracle

        self.q = np.empty((len(matrix), self.T))
        for i, row in enumerate(matrix):
            for alias, v in zip(self.names, row):
                comp, name = self.m.aliases[alias]
                assert self.m.settings.set_parameter_value(comp, name, float(np.float32(v)))  # the engine holds float32
            om = hb_oracle.OracleModel(self.m.settings.get_structure(), self.units, self.solver, self.T)
            for k, v in self.forcing.items():
                om.set_forcing(k, v)
            self.q[i] = om.run()

    def evaluate_batch(self, metric, obs, warmup):
        from hydrobricks_b200 import evaluation

        return np.asarray(getattr(evaluation, metric)(self.q, obs, warmup), dtype=np.float64)


def setup(solver):
    import bench

    U, T = 12, 900
    units = sorted(bench.hydro_units(U), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(T)
    p2, t2, e2 = bench.spatialise(units, precip, temp, pet)
    m = OracleBackedSocont(solver, units, {"precipitation": p2, "temperature": t2, "pet": e2}, T)
    truth = {"a_snow": 5.0, "A": 400.0, "k_slow_1": 0.05, "beta": 2000.0}
    m.run_sets(m.parameter_matrix({k: [v] for k, v in truth.items()}, 1))
    return m, truth, m.q[0].copy()


def main():
    from hydrobricks_b200 import trainer

    which = sys.argv[1:] or ["mc", "sceua"]
    warm = 120
    if "mc" in which:
        t0 = time.time()
        m, truth, obs = setup("crank_nicolson")
        obs[[200, 555]] = np.nan
        space = trainer.ParameterSpace.for_socont(list(truth), transforms=None)
        mc = trainer.BatchedMonteCarlo(m, space, obs, warmup=warm, objective="nse")
        best, score = mc.run(3000, batch=1024, seed=7)
        assert len(mc.scores) == 3000 and np.isfinite(mc.scores).all()
        assert score > 0.9 and score == mc.scores.max()
        assert abs(mc.evaluate({k: np.array([v]) for k, v in truth.items()})[0] - 1.0) < 1e-12
        print(f"mc: best NSE {score:.6f} (> 0.9) {best} in {time.time() - t0:.0f} s", flush=True)
    if "sceua" in which:
        t0 = time.time()
        m, truth, obs = setup("heun_explicit")
        space = trainer.ParameterSpace.for_socont(list(truth), transforms=None)
        sce = trainer.BatchedSceUa(m, space, obs, warmup=warm, objective="nse")
        best, score = sce.run(4000, ngs=16, seed=11)
        assert score > 0.999 and score == sce.scores.max()
        assert sce.launches == 1 + sce.loops * 9 and len(sce.scores) == 16 * 9 + sce.loops * 9 * 48
        mc = trainer.BatchedMonteCarlo(m, space, obs, warmup=warm, objective="nse")
        mc_score = mc.run(4000, batch=4000, seed=11)[1]
        assert mc_score < score
        print(f"sceua: best NSE {score:.9f} (> 0.999), Monte-Carlo with as many draws {mc_score:.6f}, {best} "
              f"in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
