"""Calibration bridge to the batched engine.

The reference calibrates through SPOTPY: `SpotpySetup.parameters()` draws ONE parameter set and rejects it until the
ParameterSet's constraints and ranges hold (trainer.py:934-964), `simulation()` runs the model once for it
(trainer.py:966-1075), `objectivefunction()` scores the outlet discharge after the warm-up against the observations
(trainer.py:1091-1141, evaluation/metrics.py:40-62). One evaluation = one `ModelHydro::Run()`.

Here the same three steps are batched: `ParameterSpace.sample()` draws n sets with the same rejection rule,
`Socont.run_sets()` integrates them in one `hbk_run`, and the device scores every set (`hbk_compute_scores`), so only
n scores travel back. SPOTPY itself is not in this image; its Monte-Carlo sampler — the one the reference's
`examples/basics/analyse_mc_socont_parallel.py` uses — is what `BatchedMonteCarlo` restates. Samplers that propose
one set from the previous score (SCE-UA, DREAM …) can use `evaluate()` with whatever population they hold.
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
_OPS = {"<": np.less, "lt": np.less, "<=": np.less_equal, "le": np.less_equal, ">": np.greater, "gt": np.greater,
        ">=": np.greater_equal, "ge": np.greater_equal}
# the objective functions hbk_compute_scores has on the device (HydroErr's names) and their orientation
# (evaluation/metrics.py:9-38: the error metrics are the ones where lower is better)
HIGHER_IS_BETTER = {"nse": True, "kge_2012": True, "kge_2009": True, "r_squared": True, "mae": False, "mse": False,
                    "rmse": False, "nrmse_mean": False}


class ParameterSpace:
    """The parameters allowed to change, their ranges and the constraints between them
    (ParameterSet.allow_changing / change_range / define_constraint, parameters.py:807-1078)."""

    def __init__(self, ranges, constraints=()):
        self.ranges = {k: (float(lo), float(hi)) for k, (lo, hi) in ranges.items()}
        for name, (lo, hi) in self.ranges.items():
            if not lo <= hi:
                raise ValueError(f"range of {name}: min {lo} > max {hi}")
        self.constraints = [tuple(c) for c in constraints]
        for a, op, b in self.constraints:
            if op not in _OPS:
                raise ValueError(f"unknown constraint operator {op!r}")
        self.names = list(self.ranges)

    @classmethod
    def for_socont(cls, names, constraints=SOCONT_CONSTRAINTS):
        unknown = [n for n in names if n not in SOCONT_RANGES]
        if unknown:
            raise ValueError(f"no reference range for {unknown}")
        return cls({n: SOCONT_RANGES[n] for n in names}, constraints)

    @classmethod
    def from_parameter_set(cls, parameter_set, allow_changing=None):
        """From the reference's ParameterSet (its `parameters` table with aliases/min/max, its `constraints` and
        `allow_changing` lists), so that existing calibration scripts keep their parameter definitions."""
        names = list(allow_changing if allow_changing is not None else parameter_set.allow_changing)
        ranges = {}
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
        return cls(ranges, [tuple(c) for c in parameter_set.constraints])

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

    def sample(self, n, rng, max_attempts=1000):
        """n uniform draws; a draw that violates a constraint is redrawn, at most `max_attempts` times, then the
        sampler gives up like SpotpySetup.parameters() (trainer.py:948-964)."""
        values = {name: rng.uniform(lo, hi, n) for name, (lo, hi) in self.ranges.items()}
        bad = ~self.satisfied(values)
        for _ in range(max_attempts):
            if not bad.any():
                return values
            k = int(bad.sum())
            for name, (lo, hi) in self.ranges.items():
                values[name][bad] = rng.uniform(lo, hi, k)
            bad = ~self.satisfied(values)
        if bad.any():
            raise RuntimeError("The parameter constraints could not be satisfied.")
        return values


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
