"""Randomised parity of the device arithmetic (hbk_device.cuh compiled for the host, tests/host_emul) against the oracle
restatement over the parameter space (tools/fuzz_cases.py): every solver, glacier / 1-2 soil stores / both quick-flow
laws, and — for two thirds of the cases — soil capacities below 50 mm where the capacity branches of the flux
constraint bind.

The bar is the GPU one (1e-10 relative + 1e-12 mm) wherever the reference's constraint sweep CONVERGES. Where it does
not ("The storage constraints did not stabilize after 10 sweeps", Processor.cpp:191-209: the overflow assignment of
WaterContainer.cpp:147-153 flips between two values and the result is whatever the 10th sweep leaves), the reference's
own result hangs on the last bit of `content + change*dt > capacity`: the reference built with -march=native
-ffp-contract=fast differs from its -O3 build by up to 70 % of the discharge on 5 of the first 120 cases here
(DESIGN.md §4), so no implementation can be held to 1e-10 there. Those cases must be exactly the ones that exceed the
bar, and the engine reports them (hbk_last_run_stats: unconverged_sweeps)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

import fuzz_cases  # noqa: E402
from test_device_math_on_host import emul, emul_faithful, emul_fast, run_emul  # noqa: E402,F401

RTOL, ATOL = 1e-10, 1e-12
N = 48


def _excess(a, b):
    return float(np.max(np.abs(a - b) / (RTOL * np.abs(b) + ATOL)))


SOLVERS = ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "implicit_euler", "exponential_euler",
           "analytic_linear"]


def _worst(L, seed, solver):
    meta, forcing, om, vals = fuzz_cases.run_oracle(seed, solver)
    _, q, rec, unc = run_emul(L, meta, forcing)
    return max(_excess(q, om.outlet), _excess(rec["slow"], om.records["slow_reservoir:water_content"])), om, vals, q, rec, unc


@pytest.mark.parametrize("solver", SOLVERS)
def test_device_arithmetic_fuzz(solver, emul):  # noqa: F811
    """The shipped arithmetic: inside the bar on every case whose reference sweep converges; where it does not, RK4 still
    holds it on every case of this sample (the quotients round like the reference's divisions), Heun / Euler on most."""
    converged = beyond = 0
    for seed in range(N):
        worst, om, vals, _, _, _ = _worst(emul, seed, solver)
        if om.unconverged == 0:
            converged += 1
            assert worst <= 1.0, f"seed {seed} ({vals}): {worst:.3g} x the tolerance with a converged sweep"
        elif worst > 1.0:
            beyond += 1
            assert float(vals["A"]) < 50.0, f"seed {seed}: non-converged sweep at A = {vals['A']}"
    assert converged >= N * 0.6
    if solver == "rk4":
        assert beyond == 0
    elif solver in ("heun_explicit", "euler_explicit"):
        assert beyond <= 3  # the reference's own -ffp-contract=fast build differs on 5 of 120 such cases
    else:
        assert converged == N  # the sequential solvers apply the constraint once: nothing to converge


@pytest.mark.parametrize("solver", ["rk4", "heun_explicit", "euler_explicit"])
def test_default_build_is_bit_identical_to_the_faithful_build_on_the_fuzz_cases(solver, emul, emul_faithful):  # noqa: F811
    """divExact (reciprocal + two fused multiply-adds) against the divisions written as divisions, over the randomised
    parameter space incl. the non-converging small-capacity regime, where one different bit in a quotient would show."""
    non_converged = 0
    for seed in range(N):
        _, om, vals, qa, ra, ua = _worst(emul, seed, solver)
        meta, forcing, _, _ = fuzz_cases.run_oracle(seed, solver)
        _, qb, rb, ub = run_emul(emul_faithful, meta, forcing)
        non_converged += om.unconverged > 0
        assert np.array_equal(qa, qb) and np.array_equal(ra["slow"], rb["slow"]) and ua == ub, f"seed {seed} ({vals})"
    assert non_converged > 0


def test_fast_arithmetic_variant_loses_the_non_converging_regime(emul_fast):  # noqa: F811
    """-DHBK_FAST_ARITH=1 (round-1 kernels: bare reciprocals, verification sweep skipped) is fine wherever the
    reference's sweep converges and NOT where it does not — the reason it is no longer the default."""
    beyond = 0
    for seed in range(N):
        worst, om, vals, _, _, _ = _worst(emul_fast, seed, "rk4")
        if om.unconverged == 0:
            assert worst <= 1.0, f"seed {seed} ({vals})"
        else:
            beyond += worst > 1.0
    assert beyond > 0


@pytest.mark.parametrize("solver", ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson"])
def test_snow_retention_fuzz(solver, emul):  # noqa: F811
    """Liquid water in the snowpacks over the parameter space: holding capacity 0-0.2, refreezing factor 0-0.1 (two cases
    out of three), the rain through the snowpack (one out of two), with / without a glacier cover, soil capacities down to
    0 — snowpackStepRetention against the oracle's outflow:snow_holding / refreeze:degree_day (pinned on the reference's
    goldens snow_retention_*). Same rule as above: inside the bar wherever the reference's sweep converges."""
    converged = 0
    for seed in range(24):
        meta, forcing, om, vals = fuzz_cases.run_oracle(seed, solver, retention=True)
        _, q, rec, _ = run_emul(emul, meta, forcing)
        worst = max(_excess(q, om.outlet), _excess(rec["slow"], om.records["slow_reservoir:water_content"]),
                    _excess(rec["snow"], om.records["ground_snowpack:snow_content"]))
        if om.unconverged == 0:
            converged += 1
            assert worst <= 1.0, f"seed {seed} ({vals}): {worst:.3g} x the tolerance with a converged sweep"
        else:
            assert float(vals["A"]) < 50.0, f"seed {seed}: non-converged sweep at A = {vals['A']}"
    assert converged >= 14
