"""Host logic of the calibration bridge (hydrobricks_b200/trainer.py) on CPU: the batched restatement of
SpotpySetup.parameters() / simulation() / objectivefunction() (reference trainer.py:934-1141)."""
import sys
from pathlib import Path

import numpy as np
import pytest

from hydrobricks_b200 import trainer

ROOT = Path(__file__).resolve().parent.parent


def test_sample_respects_ranges_and_constraints():
    space = trainer.ParameterSpace.for_socont(["a_snow", "a_ice", "k_slow_1", "k_slow_2", "k_quick", "A"], transforms=None)
    assert not space.transforms
    v = space.sample(5000, np.random.default_rng(1))
    # without transforms the draws are plain uniform ones in the order of the names, the rejected sets redrawn together
    rng = np.random.default_rng(1)
    plain = {name: rng.uniform(lo, hi, 5000) for name, (lo, hi) in space.ranges.items()}
    bad = ~space.satisfied(plain)
    while bad.any():
        for name, (lo, hi) in space.ranges.items():
            plain[name][bad] = rng.uniform(lo, hi, int(bad.sum()))
        bad = ~space.satisfied(plain)
    assert all(np.array_equal(v[name], plain[name]) for name in space.names)
    assert set(v) == set(space.names) and all(len(x) == 5000 for x in v.values())
    assert (v["a_snow"] < v["a_ice"]).all()  # models/socont.py:280-285
    assert (v["k_slow_2"] < v["k_slow_1"]).all() and (v["k_slow_1"] < v["k_quick"]).all()
    for name, (lo, hi) in space.ranges.items():
        assert (v[name] >= lo).all() and (v[name] <= hi).all()
    assert space.satisfied(v).all()
    # constraints on parameters that are not part of the space are ignored (parameters.py:1089-1091)
    only = trainer.ParameterSpace.for_socont(["a_snow", "A"])
    assert only.satisfied(only.sample(100, np.random.default_rng(2))).all()


def test_sampler_gives_up_like_the_reference():
    space = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (2.0, 3.0)}, [("x", ">", "y")])
    with pytest.raises(RuntimeError, match="constraints could not be satisfied"):
        space.sample(10, np.random.default_rng(0), max_attempts=20)
    with pytest.raises(ValueError):
        trainer.ParameterSpace({"x": (1.0, 0.0)})
    with pytest.raises(ValueError):
        trainer.ParameterSpace({"x": (0.0, 1.0)}, [("x", "~", "x")])


class _FakeModel:
    """Stands in for models.Socont: the 'outlet' of a set is its parameter x repeated, the score 1 - (x - 0.3)^2."""

    def __init__(self):
        self.runs = []

    def parameter_matrix(self, sets, n):
        assert all(len(np.atleast_1d(v)) == n for v in sets.values())
        return np.column_stack([np.asarray(sets["x"], dtype=np.float32), np.asarray(sets["y"], dtype=np.float32)])

    def run_sets(self, matrix):
        self.runs.append(len(matrix))
        self._m = matrix

    def evaluate_batch(self, metric, obs, warmup):
        return 1.0 - (self._m[:, 0].astype(np.float64) - 0.3) ** 2


def test_monte_carlo_batches_scores_and_best():
    space = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (0.0, 1.0)}, [("x", "<", "y")])
    model = _FakeModel()
    mc = trainer.BatchedMonteCarlo(model, space, np.zeros(10), warmup=2, objective="nse")
    best, score = mc.run(1000, batch=300, seed=3)
    assert model.runs == [300, 300, 300, 100]  # one launch per batch
    assert len(mc.scores) == 1000 and all(len(v) == 1000 for v in mc.parameters.values())
    assert abs(best["x"] - 0.3) < 0.02 and score == mc.scores.max() and best["x"] < best["y"]
    # rejected sets are not run and get the worst score (trainer.py:1007-1020, 1101-1108)
    s = mc.evaluate({"x": np.array([0.2, 0.9, 0.1, 1.5]), "y": np.array([0.5, 0.1, 0.4, 2.0])})
    assert model.runs[-1] == 2 and np.isneginf(s[[1, 3]]).all() and np.isfinite(s[[0, 2]]).all()
    with pytest.raises(ValueError):
        trainer.BatchedMonteCarlo(model, space, np.zeros(10), objective="mape")  # not on the device


@pytest.mark.skipif(not Path("/root/reference/python/src/hydrobricks").exists(), reason="reference sources not available")
def test_space_from_the_reference_parameter_set():
    """The reference's own ParameterSet (ranges, aliases, constraints of hb.models.Socont) feeds the batched sampler."""
    sys.path.insert(0, str(ROOT / "tools"))
    import refpy

    refpy.load_reference_python("b200")
    import hydrobricks.models as models

    socont = models.Socont(soil_storage_nb=2, solver="rk4")
    ps = socont.generate_parameters()
    space = trainer.ParameterSpace.from_parameter_set(ps, ["a_snow", "A", "k_slow_1", "k_slow_2", "percol", "beta"])
    assert space.ranges["A"] == (0.0, 3000.0) and space.ranges["beta"] == (100.0, 30000.0)
    assert space.ranges["a_snow"] == (2.0, 20.0) and space.ranges["percol"] == (0.0, 10.0)
    v = space.sample(2000, np.random.default_rng(5))
    assert (v["k_slow_2"] < v["k_slow_1"]).all()
    for name in space.names:
        assert trainer.SOCONT_RANGES[name] == space.ranges[name]
    # a transform attached to the reference's ParameterSet (set_transform, parameters.py:948-988; scalar-only callables
    # as in the reference's own tests) moves the search of that parameter to the transformed space
    import math

    # the reference's Socont searches the response factors as lk = log(k / 24) (models/socont.py:287-321): the same
    # convention as SOCONT_TRANSFORMS
    assert sorted(space.transforms) == ["k_slow_1", "k_slow_2"]
    own = trainer.ParameterSpace.for_socont(["k_slow_1", "k_slow_2", "A"])
    assert sorted(own.transforms) == ["k_slow_1", "k_slow_2"]
    for name in own.transforms:
        assert np.allclose(own.optimizer_bounds(name), space.optimizer_bounds(name), rtol=1e-15)
        assert np.allclose(own.to_real(name, np.array([-9.0, -5.5])), space.to_real(name, np.array([-9.0, -5.5])), rtol=1e-15)
    ps.set_transform("k_slow_1", math.log, math.exp)
    space = trainer.ParameterSpace.from_parameter_set(ps, ["k_slow_1", "k_slow_2", "A"])
    assert np.allclose(space.optimizer_bounds("k_slow_1"), sorted([math.log(space.ranges["k_slow_1"][0]), 0.0]))
    assert space.optimizer_bounds("A") == (0.0, 3000.0)
    v = space.sample(4000, np.random.default_rng(6))
    assert space.satisfied(v).all() and (v["k_slow_2"] < v["k_slow_1"]).all()


class _FakeModelN:
    """A model whose score is an analytic function of its parameter sets (so the optimiser can be checked on CPU)."""

    def __init__(self, names, score):
        self.names, self.score, self.runs = list(names), score, []

    def parameter_matrix(self, sets, n):
        return np.column_stack([np.asarray(sets[k], dtype=np.float64) for k in self.names])

    def run_sets(self, matrix):
        self.runs.append(len(matrix))
        self._m = matrix

    def evaluate_batch(self, metric, obs, warmup):
        return self.score(self._m)


def test_sceua_evolves_its_complexes_side_by_side():
    """Shuffled complex evolution on the batched engine: every evolution step of all complexes is ONE launch of 3 x ngs
    candidate sets; the speculative launches change how many launches there are, not the trajectory."""
    names = ["a", "b", "c", "d"]
    target = np.array([0.3, 0.7, 0.45, 0.9])

    def score(m):  # a curved valley (Rosenbrock-like) with its maximum 1 at `target`
        z = (m - target) * 4
        return 1.0 - np.sum((z[:, 1:] - z[:, :-1] ** 2) ** 2 + z[:, :-1] ** 2, axis=1) - z[:, -1] ** 2

    space = trainer.ParameterSpace({k: (0.0, 1.0) for k in names}, [("a", "<", "b")])
    ngs, n = 6, len(names)
    model = _FakeModelN(names, score)
    sce = trainer.BatchedSceUa(model, space, np.zeros(4), objective="nse")
    best, best_score = sce.run(4000, ngs=ngs, seed=5)
    assert best_score > 1 - 1e-4 and np.abs(np.array([best[k] for k in names]) - target).max() < 0.02
    assert best["a"] < best["b"] and best_score == sce.scores.max()
    npg = 2 * n + 1
    assert model.runs[0] == ngs * npg  # the initial population in one launch
    assert sce.launches == 1 + sce.loops * npg == len(model.runs)  # then one launch per evolution step
    assert all(r <= 3 * ngs for r in model.runs[1:]) and max(model.runs[1:]) > 2 * ngs  # (sets outside the constraint do not run)
    assert sce.evaluations <= len(sce.scores) and sce.evaluations >= ngs * npg + sce.loops * npg * ngs
    assert len(sce.scores) == sum(len(v) for v in [sce.parameters["a"]]) == ngs * npg + sce.loops * npg * 3 * ngs

    # the same trajectory without speculation: fewer sets run, up to three launches per step, same answer
    model2 = _FakeModelN(names, score)
    seq = trainer.BatchedSceUa(model2, space, np.zeros(4), objective="nse")
    best2, score2 = seq.run(4000, ngs=ngs, seed=5, speculative=False)
    assert best2 == best and score2 == best_score and seq.evaluations == sce.evaluations and seq.loops == sce.loops
    assert seq.launches > sce.launches and len(seq.scores) <= seq.evaluations < len(sce.scores)


def test_sceua_minimises_error_metrics_and_stops_on_convergence():
    names = ["x", "y"]
    model = _FakeModelN(names, lambda m: np.sqrt((m[:, 0] - 0.25) ** 2 + (m[:, 1] - 0.6) ** 2))
    space = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (0.0, 1.0)})
    sce = trainer.BatchedSceUa(model, space, np.zeros(4), objective="rmse")
    best, err = sce.run(100000, ngs=4, seed=1, peps=1e-4)
    assert sce.stop_reason.startswith("parameter space converged") and sce.evaluations < 100000
    assert err < 1e-3 and abs(best["x"] - 0.25) < 1e-3 and abs(best["y"] - 0.6) < 1e-3 and err == sce.scores.min()
    # the budget ends the search when nothing converges first
    sce2 = trainer.BatchedSceUa(_FakeModelN(names, model.score), space, np.zeros(4), objective="rmse")
    sce2.run(60, ngs=4, seed=1)
    assert sce2.stop_reason == "repetitions" and sce2.loops >= 1
    # a flat objective stops on kstop / pcento
    flat = trainer.BatchedSceUa(_FakeModelN(names, lambda m: np.ones(len(m))), space, np.zeros(4), objective="rmse")
    flat.run(100000, ngs=2, seed=0, kstop=3, peps=0.0)
    assert flat.stop_reason.startswith("objective converged")


def test_latin_hypercube_uses_every_stratum_once_and_calibrate_dispatches():
    names = ["x", "y"]
    space = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (2.0, 4.0)}, [])
    model = _FakeModelN(names, lambda m: 1.0 - (m[:, 0] - 0.3) ** 2 - (m[:, 1] - 3.1) ** 2)
    lhs = trainer.calibrate(model, space, np.zeros(4), "lhs", 500, batch=200)
    assert isinstance(lhs, trainer.BatchedLatinHypercube) and model.runs == [200, 200, 100]
    assert sorted(np.floor(lhs.parameters["x"] * 500).astype(int)) == list(range(500))
    assert sorted(np.floor((lhs.parameters["y"] - 2.0) / 2.0 * 500).astype(int)) == list(range(500))
    best, score = lhs.best()
    assert score == lhs.scores.max() and abs(best["x"] - 0.3) < 0.1 and abs(best["y"] - 3.1) < 0.2
    # a constraint is not part of the design: violating sets are not run and score worst
    constrained = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (0.0, 1.0)}, [("x", "<", "y")])
    model2 = _FakeModelN(names, lambda m: -np.abs(m[:, 0] - 0.2))
    lhs2 = trainer.calibrate(model2, constrained, np.zeros(4), "lhs", 400)
    assert 100 < model2.runs[0] < 300 and np.isneginf(lhs2.scores).sum() == 400 - model2.runs[0]
    assert lhs2.best()[0]["x"] < lhs2.best()[0]["y"]
    sce = trainer.calibrate(model2, constrained, np.zeros(4), "sceua", 600, ngs=3, seed=2)
    assert isinstance(sce, trainer.BatchedSceUa) and abs(sce.best()[0]["x"] - 0.2) < 0.01
    mc = trainer.calibrate(model2, constrained, np.zeros(4), "mc", 300, batch=100)
    assert isinstance(mc, trainer.BatchedMonteCarlo) and len(mc.scores) == 300
    with pytest.raises(ValueError, match="not available on the batched engine"):
        trainer.calibrate(model2, constrained, np.zeros(4), "dream", 10)


def test_transforms_move_the_search_to_the_transformed_space():
    """ParameterTransform (parameters.py:43-56): uniform in log space = log-uniform real values (half of the draws below
    the geometric mean of the range), real values everywhere outside the sampler, decreasing maps handled, and SCE-UA /
    the Latin hypercube search the same space."""
    space = trainer.ParameterSpace({"k": (1e-4, 1.0), "y": (0.0, 1.0)}, [],
                                   transforms={"k": (np.log10, lambda t: 10.0 ** t)})
    assert np.allclose(space.optimizer_bounds("k"), (-4.0, 0.0)) and space.optimizer_bounds("y") == (0.0, 1.0)
    v = space.sample(20000, np.random.default_rng(0))
    assert v["k"].min() >= 1e-4 and v["k"].max() <= 1.0 and abs(np.mean(v["k"] < 1e-2) - 0.5) < 0.02
    assert abs(np.mean(v["y"] < 0.5) - 0.5) < 0.02
    # a decreasing map (1 / k) and a constraint checked on the real values
    inv = trainer.ParameterSpace({"k": (0.5, 4.0), "y": (0.0, 4.0)}, [("k", "<", "y")],
                                 transforms={"k": (lambda r: 1.0 / r, lambda t: 1.0 / t)})
    assert np.allclose(inv.optimizer_bounds("k"), (0.25, 2.0))
    w = inv.sample(5000, np.random.default_rng(1))
    assert (w["k"] < w["y"]).all() and w["k"].min() >= 0.5 and w["k"].max() <= 4.0
    with pytest.raises(ValueError, match="not part of the space"):
        trainer.ParameterSpace({"k": (0.0, 1.0)}, [], transforms={"z": (np.log, np.exp)})

    # the optimum sits three decades below the top of the range: found in log space to a relative 1e-3
    names = ["k", "y"]
    model = _FakeModelN(names, lambda m: 1.0 - (np.log10(m[:, 0]) + 3.0) ** 2 - (m[:, 1] - 0.5) ** 2)
    sce = trainer.calibrate(model, space, np.zeros(4), "sceua", 3000, ngs=4, seed=4)
    best, score = sce.best()
    assert abs(best["k"] / 1e-3 - 1.0) < 2e-2 and abs(best["y"] - 0.5) < 1e-2 and score > 1 - 1e-4
    assert sce.parameters["k"].min() >= 1e-4 * (1 - 1e-12) and sce.parameters["k"].max() <= 1.0 + 1e-12  # real values
    lhs = trainer.calibrate(model, space, np.zeros(4), "lhs", 1000)
    strata = np.floor((np.log10(lhs.parameters["k"]) + 4.0) / 4.0 * 1000 + 1e-9).astype(int)
    assert len(set(strata)) >= 998  # one draw per stratum of the LOG axis (up to rounding at the stratum edges)


def test_get_results_and_get_best_read_the_database():
    names = ["x", "y"]
    space = trainer.ParameterSpace({"x": (0.0, 1.0), "y": (0.0, 1.0)}, [("x", "<", "y")])
    model = _FakeModelN(names, lambda m: np.abs(m[:, 0] - 0.2))
    lhs = trainer.calibrate(model, space, np.zeros(4), "lhs", 300, objective="mae")
    df = trainer.get_results(lhs)
    assert list(df.columns) == ["score", "x", "y"] and len(df) == 300
    best = trainer.get_best(lhs)
    assert best["score"] == df["score"].min() and df["x"].iloc[best["index"]] == best["parameters"]["x"]
    assert np.isposinf(df["score"]).sum() > 50 and best["parameters"]["x"] < best["parameters"]["y"]  # rejected sets kept
    sce = trainer.calibrate(_FakeModelN(names, lambda m: 1 - (m[:, 0] - 0.2) ** 2), space, np.zeros(4), "sceua", 300, ngs=2)
    assert trainer.get_best(sce)["score"] >= sce.best()[1]  # the database also holds the speculative candidates
    empty = trainer.BatchedMonteCarlo(model, space, np.zeros(4))
    empty.scores, empty.parameters = np.array([np.nan, -np.inf]), {"x": np.zeros(2), "y": np.ones(2)}
    with pytest.raises(RuntimeError, match="No finite objective"):
        trainer.get_best(empty)
