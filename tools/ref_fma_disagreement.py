#!/usr/bin/env python
"""How reproducible is the REFERENCE on the ensemble fuzz sample of tests/test_gpu_parity.py? (TEST INFRASTRUCTURE)

Runs the unmodified reference core twice over the sample of test_ensemble_at_scale_matches_oracle_incl_small_capacities
— oracle/_ref (its Release flags, no FMA contraction) and the same sources built with `-O3 -march=native
-ffp-contract=fast` (path given on the command line; recipe: `make -C oracle -B ref OUT=/tmp/ref_fma REFFLAGS='… -march=native
-ffp-contract=fast …'`) — and writes tests/golden/ensemble_fuzz_reference_fma.json: for every parameter set whether the
two builds of the reference agree within the GPU bar (1e-10 relative + 1e-12), and the oracle's count of non-converged
constraint sweeps. The GPU test holds the engine to the reference's own disagreement count on the non-converging sets.

    python tools/ref_fma_disagreement.py /tmp/ref_fma/_hydrobricks.cpython-312-x86_64-linux-gnu.so
"""
import importlib.util
import json
import multiprocessing as mp
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
_W = {}


def sample():
    import bench

    P0, EXTRA, U, T = 4096, 256 + 17, 13, 2000
    P = P0 + EXTRA
    ps = bench.parameter_sets(P, seed=2024)
    rng = np.random.default_rng(7)
    ps["A"][P0:] = rng.uniform(0.0, 50.0, EXTRA)
    ps["A"][P0] = 0.0
    return ps, P, U, T


def _init(so, U, T):
    import bench
    from hydrobricks_b200 import models

    spec = importlib.util.spec_from_file_location("_hydrobricks", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.init()
    try:
        ref.set_max_log_level()
    except Exception:
        pass
    units = bench.hydro_units(U)
    precip, temp, pet = bench.station_series(T)
    m = models.Socont(solver="rk4", soil_storage_nb=1, surface_runoff="socont_runoff", land_cover_names=["ground"],
                      land_cover_types=["ground"], backend=ref)
    m.setup(units, bench.START, str(np.datetime64(bench.START) + np.timedelta64(T - 1, "D")))
    p2, t2, e2 = bench.spatialise(m.units, precip, temp, pet)
    m.set_forcing(bench.mjd_axis(T), p2, t2, e2)
    _W["m"] = m


def _run(values):
    m = _W["m"]
    m.set_parameters(values)
    m.model.reset()
    m.model.run()
    return np.array(m.model.get_outlet_discharge())


def run_build(so, sets, U, T):
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)  # the reference logs "did not stabilize" warnings by the thousand
    try:
        with mp.get_context("fork").Pool(os.cpu_count(), initializer=_init, initargs=(str(so), U, T)) as pool:
            return pool.map(_run, sets, chunksize=16)
    finally:
        os.dup2(saved, 2)


def main():
    import test_gpu_parity as tg

    fma_so = Path(sys.argv[1])
    ref_so = sorted((ROOT / "oracle" / "_ref").glob("_hydrobricks*.so"))[0]
    ps, P, U, T = sample()
    sets = [{a: float(ps[a][k]) for a in ps} for k in range(P)]
    a = run_build(ref_so, sets, U, T)
    b = run_build(fma_so, sets, U, T)
    with mp.get_context("fork").Pool(os.cpu_count(), initializer=tg._oracle_init, initargs=(U, T)) as pool:
        orc = pool.map(tg._oracle_run, sets, chunksize=16)
    disagree, unconverged, oracle_exact = [], [], 0
    for k in range(P):
        ok, _ = tg.close(b[k], a[k])
        disagree.append(0 if ok else 1)
        unconverged.append(int(orc[k][1]))
        oracle_exact += bool(np.array_equal(orc[k][0], a[k]))
    rest = [k for k in range(P) if unconverged[k] > 0]
    out = {"sample": "bench.parameter_sets(4369, seed=2024), A[4096:] ~ U(0, 50) (default_rng(7)), A[4096] = 0; 13 units, "
                     "2000 steps; tools/ref_fma_disagreement.py",
           "n_sets": P, "oracle_bit_identical_to_reference": oracle_exact,
           "n_nonconverged": len(rest),
           "reference_fma_disagrees_nonconverged": int(sum(disagree[k] for k in rest)),
           "reference_fma_disagrees_converged": int(sum(disagree[k] for k in range(P) if unconverged[k] == 0)),
           "unconverged_sets": rest,
           "reference_fma_disagrees": [k for k in range(P) if disagree[k]]}
    path = ROOT / "tests" / "golden" / "ensemble_fuzz_reference_fma.json"
    path.write_text(json.dumps(out))
    print({k: v for k, v in out.items() if not isinstance(v, list)})


if __name__ == "__main__":
    main()
