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
from test_device_math_on_host import emul, emul_faithful, run_emul  # noqa: E402,F401

RTOL, ATOL = 1e-10, 1e-12
N = 48


def _excess(a, b):
    return float(np.max(np.abs(a - b) / (RTOL * np.abs(b) + ATOL)))


@pytest.mark.parametrize("solver", ["rk4", "heun_explicit", "euler_explicit", "crank_nicolson", "implicit_euler",
                                    "exponential_euler", "analytic_linear"])
def test_device_arithmetic_fuzz(solver, emul):  # noqa: F811
    converged = beyond = 0
    for seed in range(N):
        meta, forcing, om, vals = fuzz_cases.run_oracle(seed, solver)
        _, q, rec, _ = run_emul(emul, meta, forcing)
        worst = max(_excess(q, om.outlet), _excess(rec["slow"], om.records["slow_reservoir:water_content"]))
        if om.unconverged == 0:
            converged += 1
            assert worst <= 1.0, f"seed {seed} ({vals}): {worst:.3g} x the tolerance with a converged sweep"
        elif worst > 1.0:
            beyond += 1
            assert float(vals["A"]) < 50.0, f"seed {seed}: non-converged sweep at A = {vals['A']}"
    assert converged >= N * 0.6
    if solver in ("rk4", "heun_explicit", "euler_explicit"):
        assert beyond > 0  # the sample does contain the ill-conditioned regime
    else:
        assert converged == N  # the sequential solvers apply the constraint once: nothing to converge


def test_faithful_build_holds_the_bar_in_the_non_converging_regime(emul_faithful):  # noqa: F811
    """-DHBK_FAITHFUL=1 (divisions as the reference writes them, verification sweep kept): RK4 stays within 1e-10 of
    the reference on every case, converged or not — the shipped reciprocal forms lose about half of the non-converged
    ones (test above). Evidence for the option documented in hbk_device.cuh / DESIGN.md §4."""
    non_converged = 0
    for seed in range(N):
        meta, forcing, om, vals = fuzz_cases.run_oracle(seed, "rk4")
        _, q, rec, _ = run_emul(emul_faithful, meta, forcing)
        worst = max(_excess(q, om.outlet), _excess(rec["slow"], om.records["slow_reservoir:water_content"]))
        non_converged += om.unconverged > 0
        assert worst <= 1.0, f"seed {seed} ({vals}): {worst:.3g} x the tolerance (unconverged sweeps: {om.unconverged})"
    assert non_converged > 0
