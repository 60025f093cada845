"""Calibration bridge to the batched engine.

The reference calibrates through SPOTPY: `SpotpySetup.parameters()` draws ONE parameter set and rejects it until the
ParameterSet's constraints and ranges hold (trainer.py:934-964), `simulation()` runs the model once for it
(trainer.py:966-1075), `objectivefunction()` scores the outlet discharge after the warm-up against the observations
(trainer.py:1091-1141, evaluation/metrics.py:40-62). One evaluation = one `ModelHydro::Run()`.

Here the same three steps are batched: `ParameterSpace.sample()` draws n sets with the same rejection rule,
`Socont.run_sets()` integrates them in one `hbk_run`, and the device scores every set (`hbk_compute_scores`), so only
n scores travel back. SPOTPY itself is not in this image; its Monte-Carlo sampler — the one the reference's
`examples/basics/analyse_mc_socont_parallel.py` uses — is what `BatchedMonteCarlo` restates. SCE-UA, the algorithm
most of the reference's examples calibrate with (`calibrate(spot_setup, "sceua", …)`,
examples/basics/calibrate_sceua_socont.py), is restated from the published algorithm in `BatchedSceUa`: its complexes
evolve side by side, one launch per evolution step. Other samplers that propose sets from previous scores (DREAM …) can
use `evaluate()` with whatever population they hold.
"""
from __future__ import annotations

import numpy as np

# Ranges of the reference's ParameterSet for the Socont family (python/src/hydrobricks/parameters.py:66-174, 736-738,
# 1722-1870): alias -> (min, max).
SOCONT_RANGES = {
    "a_snow": (2.0, 20.0), "a_ice": (2.0, 20.0), "A": (0.0, 3000.0), "k_slow_1": (1e-4, 1.0), "k_slow_2": (1e-4, 1.0),
    "k_quick": (1e-4, 1.0), "beta": (100.0, 30000.0), "percol": (0.0, 10.0), "k_snow": (1e-4, 1.0),
    "k_ice": (1e-4, 1.0), "prec_t_start": (-2.0, 2.0), "prec_t_end": (0.0, 4.0),
    # liquid water in the snowpacks (parameters.py:444-467): refreezing factor, water holding capacity
    "cfr": (0.0, 0.1), "whc": (0.0, 0.2), "cwh": (0.0, 0.2),
}
# models/socont.py:280-285 plus the splitter's transition_start < transition_end
SOCONT_CONSTRAINTS = [("a_snow", "<", "a_ice"), ("k_slow_1", "<", "k_quick"), ("k_slow_2", "<", "k_quick"),
                      ("k_slow_2", "<", "k_slow_1"), ("prec_t_start", "<", "prec_t_end")]
# models/socont.py:287-321: the reference's Socont calibrates the slow reservoirs' response factors as lk = log(k [1/h]),
# the original SOCONT formulation, k [1/d] = 24 k [1/h]: alias -> (real -> transformed, transformed -> real)
SOCONT_TRANSFORMS = {
    "k_slow_1": (lambda rf: np.log(np.asarray(rf) / 24.0), lambda lk: 24.0 * np.exp(lk)),
    "k_slow_2": (lambda rf: np.log(np.asarray(rf) / 24.0), lambda lk: 24.0 * np.exp(lk)),
}
_OPS = {"<": np.less, "lt": np.less, "<=": np.less_equal, "le": np.less_equal, ">": np.greater, "gt": np.greater,
        ">=": np.greater_equal, "ge": np.greater_equal}
# the objective functions hbk_compute_scores has on the device (HydroErr's names) and their orientation
# (evaluation/metrics.py:9-38: the error metrics are the ones where lower is better)
HIGHER_IS_BETTER = {"nse": True, "kge_2012": True, "kge_2009": True, "r_squared": True, "mae": False, "mse": False,
                    "rmse": False, "nrmse_mean": False}


class ParameterSpace:
    """The parameters allowed to change, their ranges and the constraints between them
    (ParameterSet.allow_changing / change_range / define_constraint, parameters.py:807-1078)."""

    def __init__(self, ranges, constraints=(), transforms=None):
        """transforms: {name: (to_transformed, to_real)} or objects with those two attributes (the reference's
        ParameterTransform, parameters.py:43-56). The samplers then search that parameter in the transformed space —
        uniform between the transformed bounds, like ParameterSet.get_for_spotpy (parameters.py:1492-1502) — and every
        value that reaches the model, the constraints or the database is mapped back to the real one
        (SpotpySetup.parameters(), trainer.py:955-956)."""
        self.ranges = {k: (float(lo), float(hi)) for k, (lo, hi) in ranges.items()}
        for name, (lo, hi) in self.ranges.items():
            if not lo <= hi:
                raise ValueError(f"range of {name}: min {lo} > max {hi}")
        self.constraints = [tuple(c) for c in constraints]
        for a, op, b in self.constraints:
            if op not in _OPS:
                raise ValueError(f"unknown constraint operator {op!r}")
        self.names = list(self.ranges)
        self.transforms = {}
        for name, t in (transforms or {}).items():
            if name not in self.ranges:
                raise ValueError(f"transform of {name}: the parameter is not part of the space")
            self.transforms[name] = (t.to_transformed, t.to_real) if hasattr(t, "to_real") else tuple(t)

    def optimizer_bounds(self, name):
        """The search interval of a parameter: its range, mapped through its transform if it has one (sorted: the map
        may be decreasing)."""
        lo, hi = self.ranges[name]
        if name in self.transforms:
            forward = self.transforms[name][0]
            lo, hi = sorted([float(forward(lo)), float(forward(hi))])
        return lo, hi

    def to_real(self, name, values):
        """Optimizer-space values of one parameter -> real values."""
        values = np.asarray(values, dtype=np.float64)
        if name not in self.transforms:
            return values
        inverse = self.transforms[name][1]
        try:
            real = np.asarray(inverse(values), dtype=np.float64)
            if real.shape == values.shape:
                return real
        except (TypeError, ValueError):
            pass
        return np.array([float(inverse(float(v))) for v in values.ravel()]).reshape(values.shape)  # scalar-only callables

    @classmethod
    def for_socont(cls, names, constraints=SOCONT_CONSTRAINTS, transforms="reference"):
        """The reference's ranges, constraints and — unless transforms=None (search the real values) or an own table is
        given — its log-space convention for the slow reservoirs' response factors (SOCONT_TRANSFORMS)."""
        unknown = [n for n in names if n not in SOCONT_RANGES]
        if unknown:
            raise ValueError(f"no reference range for {unknown}")
        if isinstance(transforms, str):
            if transforms != "reference":
                raise ValueError(f"transforms: 'reference', None or a table, not {transforms!r}")
            transforms = SOCONT_TRANSFORMS
        transforms = {n: t for n, t in (transforms or {}).items() if n in names}
        return cls({n: SOCONT_RANGES[n] for n in names}, constraints, transforms)

    @classmethod
    def from_parameter_set(cls, parameter_set, allow_changing=None):
        """From the reference's ParameterSet (its `parameters` table with aliases/min/max, its `constraints` and
        `allow_changing` lists), so that existing calibration scripts keep their parameter definitions."""
        names = list(allow_changing if allow_changing is not None else parameter_set.allow_changing)
        ranges, transforms = {}, {}
        for name in names:
            row = None
            for _, r in parameter_set.parameters.iterrows():
                aliases = r["aliases"] if isinstance(r["aliases"], (list, tuple)) else []
                if name in aliases or name == f'{r["component"]}:{r["name"]}':
                    row = r
                    break
            if row is None:
                raise ValueError(f"parameter {name} not found in the parameter set")
            if isinstance(row["min"], list):
                raise ValueError("list parameters are not supported by the batched sampler")
            ranges[name] = (row["min"], row["max"])
            transform = row["transform"] if "transform" in row.index else None
            if transform is not None and hasattr(transform, "to_real"):
                transforms[name] = transform
        return cls(ranges, [tuple(c) for c in parameter_set.constraints], transforms)

    def satisfied(self, values):
        """Mask [n] of the draws that satisfy every constraint whose two parameters are in the space
        (ParameterSet.constraints_satisfied ignores constraints on unused parameters, parameters.py:1089-1091) and
        every range (range_satisfied, parameters.py:1119-1150)."""
        n = len(next(iter(values.values())))
        ok = np.ones(n, dtype=bool)
        for a, op, b in self.constraints:
            if a in values and b in values:
                ok &= _OPS[op](np.asarray(values[a]), np.asarray(values[b]))
        for name, (lo, hi) in self.ranges.items():
            v = np.asarray(values[name])
            ok &= (v >= lo) & (v <= hi)
        return ok

    def sample_optimizer(self, n, rng, max_attempts=1000):
        """n uniform draws in the optimizer space [n x parameters]; a draw whose real values violate a constraint or a
        range is redrawn, at most `max_attempts` times, then the sampler gives up like SpotpySetup.parameters()
        (trainer.py:948-964)."""
        bounds = [self.optimizer_bounds(name) for name in self.names]
        x = np.column_stack([rng.uniform(lo, hi, n) for lo, hi in bounds]) if self.names else np.empty((n, 0))
        bad = ~self.satisfied(self.real_values(x))
        for _ in range(max_attempts):
            if not bad.any():
                return x
            k = int(bad.sum())
            for j, (lo, hi) in enumerate(bounds):
                x[bad, j] = rng.uniform(lo, hi, k)
            bad = ~self.satisfied(self.real_values(x))
        if bad.any():
            raise RuntimeError("The parameter constraints could not be satisfied.")
        return x

    def real_values(self, x):
        """Optimizer-space matrix [n x parameters] -> {name: real values [n]}."""
        return {name: self.to_real(name, np.ascontiguousarray(x[:, j])) for j, name in enumerate(self.names)}

    def sample(self, n, rng, max_attempts=1000):
        """n draws as real values {name: [n]} (uniform in the optimizer space, constraints and ranges satisfied)."""
        return self.real_values(self.sample_optimizer(n, rng, max_attempts))


class BatchedMonteCarlo:
    """Monte-Carlo calibration with every batch of parameter sets integrated by one engine launch.

    model: a set-up hydrobricks_b200.models.Socont with its forcing attached; observations: observed outlet
    discharge [T] (NaN = missing); warmup: leading steps left out of the score (trainer.py `warmup`);
    objective: "nse" | "kge_2012" | "kge_2009" | "r_squared" | "mae" | "mse" | "rmse" | "nrmse_mean"
    (evaluation/metrics.py:40-62; the error metrics are minimised).
    """

    def __init__(self, model, space, observations, warmup=0, objective="nse", fixed=None):
        if objective not in HIGHER_IS_BETTER:
            raise ValueError(f"objective {objective!r} is not available on the device ({', '.join(HIGHER_IS_BETTER)})")
        self.model, self.space, self.objective, self.warmup = model, space, objective, int(warmup)
        self.observations = np.ascontiguousarray(observations, dtype=np.float64)
        self.fixed = dict(fixed or {})
        self.parameters = {name: np.empty(0) for name in space.names}
        self.scores = np.empty(0)

    def evaluate(self, values):
        """Scores [n] of the given sets ({alias: array[n]}); sets outside the constraints or ranges are not run and
        get the worst score, as SpotpySetup.simulation()/objectivefunction() do (trainer.py:1007-1020, 1101-1108)."""
        n = len(next(iter(values.values())))
        ok = self.space.satisfied(values)
        scores = np.full(n, -np.inf if HIGHER_IS_BETTER[self.objective] else np.inf)
        if ok.any():
            run = {k: np.asarray(v)[ok] for k, v in values.items()}
            run.update({k: np.full(int(ok.sum()), v) for k, v in self.fixed.items()})
            self.model.run_sets(self.model.parameter_matrix(run, int(ok.sum())))
            scores[ok] = self.model.evaluate_batch(self.objective, self.observations, self.warmup)
        return scores

    def run(self, n_sets, batch=65536, seed=0):
        """Draw and score n_sets sets, `batch` per launch; keeps every draw and score (SPOTPY's database) and returns
        the best set ({alias: value}) and its score, like trainer.get_best (trainer.py:1607-1650)."""
        rng = np.random.default_rng(seed)
        drawn = {name: [] for name in self.space.names}
        scores = []
        for start in range(0, int(n_sets), int(batch)):
            n = min(int(batch), int(n_sets) - start)
            values = self.space.sample(n, rng)
            scores.append(self.evaluate(values))
            for name in self.space.names:
                drawn[name].append(values[name])
        self.parameters = {name: np.concatenate(v) for name, v in drawn.items()}
        self.scores = np.concatenate(scores)
        return self.best()

    def best(self):
        if not len(self.scores):
            raise RuntimeError("no evaluation yet")
        finite = np.where(np.isnan(self.scores), -np.inf if HIGHER_IS_BETTER[self.objective] else np.inf, self.scores)
        i = int(np.argmax(finite) if HIGHER_IS_BETTER[self.objective] else np.argmin(finite))
        return {name: float(v[i]) for name, v in self.parameters.items()}, float(self.scores[i])


class BatchedLatinHypercube(BatchedMonteCarlo):
    """Latin-hypercube sampling (SPOTPY's `lhs`): every parameter's range is cut into n_sets equal strata and each stratum
    is used exactly once, the strata being paired at random between parameters. As in SPOTPY the design ignores the
    constraints; a set that violates one is not run and gets the worst score (trainer.py:1007-1020, 1101-1108)."""

    def run(self, n_sets, batch=65536, seed=0):
        rng = np.random.default_rng(seed)
        n = int(n_sets)
        design = {}
        for name in self.space.names:  # strata of the optimizer space, values stored as real ones
            lo, hi = self.space.optimizer_bounds(name)
            design[name] = self.space.to_real(name, lo + (hi - lo) * (rng.permutation(n) + rng.uniform(0.0, 1.0, n)) / n)
        scores = [self.evaluate({k: v[i:i + int(batch)] for k, v in design.items()}) for i in range(0, n, int(batch))]
        self.parameters, self.scores = design, np.concatenate(scores) if scores else np.empty(0)
        return self.best()


class BatchedSceUa(BatchedMonteCarlo):
    """Shuffled complex evolution (SCE-UA; Duan, Sorooshian & Gupta, Water Resour. Res. 28(4) 1992 and J. Hydrol. 158, 1994)
    with its complexes evolved SIDE BY SIDE — what `calibrate(spot_setup, "sceua", repetitions, ngs=…)` does one model run
    at a time through SPOTPY (reference trainer.py:1272-1401; SPOTPY is absent from this image, so this follows the
    published algorithm with SPOTPY's argument names and defaults, not its code).

    The `ngs` complexes of a shuffling loop are independent: at each of the 2n+1 evolution steps every complex proposes
    its reflection point, its contraction point and a random point, and the engine integrates all of them in ONE launch
    (3 x ngs sets); each complex then keeps the first of the three that the algorithm would have accepted. The trajectory
    is the sequential algorithm's — `speculative=False` evaluates only what is needed, in up to three launches per step,
    and gives the same result — and only the evaluations the sequential algorithm would have made count against
    `repetitions`. On the batched engine a launch of 60 sets costs what a launch of one does, so `ngs` can be raised to
    hundreds of complexes for the price of one.
    """

    ALPHA, BETA = 1.0, 0.5  # reflection and contraction coefficients of the competitive complex evolution

    def _cost(self, values):
        scores = self.evaluate(values)
        for name in self.space.names:
            self._drawn[name].append(np.asarray(values[name], dtype=np.float64))
        self._scores.append(scores)
        self.launches += 1
        cost = -scores if HIGHER_IS_BETTER[self.objective] else scores.copy()
        return np.where(np.isnan(cost), np.inf, cost)

    def _as_values(self, x):
        return self.space.real_values(x)

    def run(self, repetitions, ngs=20, kstop=100, pcento=1e-7, peps=1e-7, seed=0, speculative=True):
        """repetitions: evaluation budget; ngs: number of complexes; kstop / pcento: stop when the best cost changed by
        less than pcento percent over kstop shuffling loops; peps: stop when the population's normalised geometric range
        falls below it. Returns the best set and its score; `evaluations`, `launches`, `loops` and `stop_reason` say how
        it got there, `parameters` / `scores` hold every set that was run (SPOTPY's database)."""
        names = self.space.names
        n = len(names)
        lo = np.array([self.space.optimizer_bounds(k)[0] for k in names])  # the population lives in the optimizer space
        hi = np.array([self.space.optimizer_bounds(k)[1] for k in names])
        span = np.where(hi > lo, hi - lo, 1.0)
        ngs, npg, nps = int(ngs), 2 * n + 1, n + 1
        nspl, npt = npg, int(ngs) * (2 * n + 1)
        rng = np.random.default_rng(seed)
        self._drawn, self._scores = {name: [] for name in names}, []
        self.launches = self.loops = 0

        x = self.space.sample_optimizer(npt, rng)
        f = self._cost(self._as_values(x))
        self.evaluations = npt
        order = np.argsort(f, kind="stable")
        x, f = x[order], f[order]
        history = [f[0]]
        self.stop_reason = "repetitions"
        # sub-complex membership: the triangular distribution that favours the better points of a complex
        weights = 2.0 * (npg - np.arange(npg)) / (npg * (npg + 1.0))
        while self.evaluations < int(repetitions):
            cx = np.stack([x[k::ngs] for k in range(ngs)])  # [ngs, npg, n]: complex k holds the points k, k + ngs, …
            cf = np.stack([f[k::ngs] for k in range(ngs)])
            for _ in range(nspl):
                members = np.stack([np.sort(rng.choice(npg, size=nps, replace=False, p=weights)) for _ in range(ngs)])
                k_idx = np.arange(ngs)
                worst = members[:, -1]
                sw, fw = cx[k_idx, worst], cf[k_idx, worst]
                centroid = np.stack([cx[k, members[k, :-1]].mean(axis=0) for k in range(ngs)])
                random_point = self.space.sample_optimizer(ngs, rng)
                reflection = centroid + self.ALPHA * (centroid - sw)
                outside = ((reflection < lo) | (reflection > hi)).any(axis=1)
                reflection[outside] = random_point[outside]  # a reflection that leaves the feasible space: mutation
                contraction = sw + self.BETA * (centroid - sw)
                if speculative:
                    cost = self._cost(self._as_values(np.concatenate([reflection, contraction, random_point])))
                    f_refl, f_contr, f_rand = cost[:ngs], cost[ngs:2 * ngs], cost[2 * ngs:]
                else:
                    f_refl = self._cost(self._as_values(reflection))
                    f_contr, f_rand = np.full(ngs, np.inf), np.full(ngs, np.inf)
                    need = f_refl > fw
                    if need.any():
                        f_contr[need] = self._cost(self._as_values(contraction[need]))
                        need = need & (f_contr > fw)
                        if need.any():
                            f_rand[need] = self._cost(self._as_values(random_point[need]))
                take_refl = f_refl <= fw
                take_contr = ~take_refl & (f_contr <= fw)
                take_rand = ~take_refl & ~take_contr
                new_x = np.where(take_refl[:, None], reflection, np.where(take_contr[:, None], contraction, random_point))
                new_f = np.where(take_refl, f_refl, np.where(take_contr, f_contr, f_rand))
                self.evaluations += int(ngs + (~take_refl).sum() + take_rand.sum())
                cx[k_idx, worst], cf[k_idx, worst] = new_x, new_f
                inner = np.argsort(cf, axis=1, kind="stable")
                cx, cf = np.take_along_axis(cx, inner[:, :, None], axis=1), np.take_along_axis(cf, inner, axis=1)
            for k in range(ngs):  # back into the population, then shuffle by sorting it as a whole
                x[k::ngs], f[k::ngs] = cx[k], cf[k]
            order = np.argsort(f, kind="stable")
            x, f = x[order], f[order]
            self.loops += 1
            history.append(f[0])
            spread = (x.max(axis=0) - x.min(axis=0)) / span
            if np.exp(np.mean(np.log(np.maximum(spread, 1e-300)))) < peps:
                self.stop_reason = "parameter space converged (peps)"
                break
            if len(history) > kstop:
                recent = np.array(history[-kstop - 1:])
                scale = np.mean(np.abs(recent))
                if scale == 0 or abs(recent[-1] - recent[0]) * 100.0 / scale < pcento:
                    self.stop_reason = "objective converged (kstop, pcento)"
                    break
        self.parameters = {name: np.concatenate(v) for name, v in self._drawn.items()}
        self.scores = np.concatenate(self._scores)
        del self._drawn, self._scores
        best_score = -f[0] if HIGHER_IS_BETTER[self.objective] else f[0]
        self._result = ({name: float(v[0]) for name, v in self._as_values(x[:1]).items()}, float(best_score))
        return self._result

    def best(self):
        """The best point of the final population (the algorithm's answer, the same with and without speculation)."""
        if not hasattr(self, "_result"):
            raise RuntimeError("no evaluation yet")
        return self._result


ALGORITHMS = {"mc": BatchedMonteCarlo, "lhs": BatchedLatinHypercube, "sceua": BatchedSceUa}


def calibrate(model, space, observations, algorithm, repetitions, warmup=0, objective="nse", fixed=None, seed=0,
              **algorithm_kwargs):
    """The reference's `calibrate(spot_setup, algorithm, repetitions, **algorithm_kwargs)` (trainer.py:1272-1401) on the
    batched engine: `algorithm` is SPOTPY's name ("mc", "lhs", "sceua"), `algorithm_kwargs` its sample() arguments
    (sceua: ngs, kstop, pcento, peps; mc / lhs: batch). Returns the sampler; `sampler.best()` is `get_best`
    (trainer.py:1607-1650), `sampler.parameters` / `sampler.scores` the database of every set that was run."""
    if algorithm not in ALGORITHMS:
        raise ValueError(f"algorithm {algorithm!r} is not available on the batched engine ({', '.join(ALGORITHMS)})")
    sampler = ALGORITHMS[algorithm](model, space, observations, warmup=warmup, objective=objective, fixed=fixed)
    sampler.run(repetitions, seed=seed, **algorithm_kwargs)
    return sampler


def get_results(sampler):
    """Every set that was run, as a table: `score` (the metric's own value, no optimiser sign) and one column of real
    values per parameter — the reference's get_results(sampler) (trainer.py:1560-1605)."""
    import pandas as pd

    return pd.DataFrame({"score": np.asarray(sampler.scores, dtype=np.float64), **sampler.parameters})


def get_best(sampler):
    """{'score', 'parameters', 'index'}: the best row of get_results (highest skill, or lowest error for the error
    metrics), like the reference's get_best(sampler) (trainer.py:1607-1672); raises when no run has a finite score."""
    scores = np.asarray(sampler.scores, dtype=np.float64)
    if not np.isfinite(scores).any():
        raise RuntimeError("No finite objective values in the calibration results.")
    higher = HIGHER_IS_BETTER[sampler.objective]
    masked = np.where(np.isfinite(scores), scores, -np.inf if higher else np.inf)
    i = int(np.argmax(masked) if higher else np.argmin(masked))
    return {"score": float(scores[i]), "parameters": {k: float(v[i]) for k, v in sampler.parameters.items()}, "index": i}
