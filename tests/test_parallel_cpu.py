"""Host logic of the multi-GPU path on CPU: world_size-2 gloo process group (parameter axis
sharded, no collective on the path, one final gather)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydrobricks_b200 import evaluation, parallel


def test_shard_bounds_cover_and_balance():
    for n in (1, 2, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_sets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_sets * 3, dtype=torch.float64).reshape(n_sets, 3)
    local = parallel.shard_rows(full, rank, world)
    # "scores" of the local shard: something that depends on the set only
    scores = (local ** 2).sum(dim=1)
    gathered = parallel.gather_rows(scores, n_sets)
    rows = parallel.gather_rows(local, n_sets)
    ok = torch.equal(gathered, (full ** 2).sum(dim=1)) and torch.equal(rows, full)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sets", [8, 9, 33])
def test_gather_rows_gloo_world2(n_sets):
    world = 2
    port = 29500 + (os.getpid() + n_sets) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_sets, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def test_batched_nse_kge_against_numpy_formulas():
    rng = np.random.default_rng(0)
    obs = rng.gamma(2.0, 2.0, 400)
    obs[[5, 77, 300]] = np.nan
    sim = obs[None, :] * rng.uniform(0.6, 1.4, (6, 1)) + rng.normal(0, 0.3, (6, 400))
    sim = np.nan_to_num(sim, nan=1.0)
    sim[0] = np.nan_to_num(obs, nan=3.0)
    warm = 30
    n1 = evaluation.nse(sim, obs, warm).numpy()
    k1 = evaluation.kge_2012(sim, obs, warm).numpy()
    keep = ~np.isnan(obs[warm:])
    o = obs[warm:][keep]
    for i in range(6):
        s = sim[i, warm:][keep]
        n0 = 1 - np.sum((s - o) ** 2) / np.sum((o - o.mean()) ** 2)
        r = np.corrcoef(s, o)[0, 1]
        k0 = 1 - np.sqrt((r - 1) ** 2 + ((s.std(ddof=1) / s.mean()) / (o.std(ddof=1) / o.mean()) - 1) ** 2 +
                         (s.mean() / o.mean() - 1) ** 2)
        assert abs(n1[i] - n0) < 1e-12 and abs(k1[i] - k0) < 1e-12
    assert abs(n1[0] - 1.0) < 1e-12  # identical series: NSE = 1 (python/tests/test_models.py:731-738)


def _reduce_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # unit-sharded basin: every rank holds the partial outlet and deposit sums of its block of
    # units; one all-reduce(sum) per series; a series that does not exist (no glacier) is None on every rank
    U, T = 11, 6
    lo, hi = parallel.shard_bounds(U, rank, world)
    contrib = torch.arange(U * T, dtype=torch.float64).reshape(U, T)
    partial = [contrib[lo:hi].sum(dim=0, keepdim=True), None, 2.0 * contrib[lo:hi].sum(dim=0, keepdim=True)]
    parallel.reduce_partials(partial)
    ok = torch.equal(partial[0], contrib.sum(dim=0, keepdim=True)) and partial[1] is None and \
        torch.equal(partial[2], 2.0 * contrib.sum(dim=0, keepdim=True))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_partials_gloo_world2():
    world = 2
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_reduce_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def test_basin_area_is_the_running_sum_in_unit_order():
    rng = np.random.default_rng(3)
    areas = rng.uniform(1e3, 1e7, 5000)
    total = 0.0
    for a in areas:
        total += float(a)
    assert parallel.basin_area(areas) == total


def test_weighted_shard_bounds_balance_the_cost():
    """Unit shards cut by integration cost (parallel.weighted_shard_bounds): contiguous, covering, and no block more than
    one unit's weight away from the mean — for a basin whose expensive (glacier) units sit at one end."""
    import bench

    units = bench.hydro_units(997, glacier=True)
    costs = parallel.unit_costs(units)
    assert set(costs) == {1.0, parallel.GLACIER_UNIT_COST}
    for world in (1, 2, 3, 4, 8):
        spans = [parallel.weighted_shard_bounds(costs, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == len(units)
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sums = [sum(costs[lo:hi]) for lo, hi in spans]
        assert max(sums) - min(sums) <= 2 * parallel.GLACIER_UNIT_COST
    assert parallel.weighted_shard_bounds([], 0, 4) == (0, 0)
    assert parallel.weighted_shard_bounds([1.0] * 10, 1, 1) == (0, 10)
