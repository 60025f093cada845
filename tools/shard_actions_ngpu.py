#!/usr/bin/env python
"""Unit shards with spin-up and dated actions across REAL ranks (NCCL): one block of a glacierised basin per GPU, the
in-run sums through parallel.make_reduce_callback (one ncclAllReduce per request of the engine), against the unsharded
run of the same basin on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/shard_actions_ngpu.py [--units 4000] [--sets 2] [--years 8]

Prints one JSON line on rank 0: max relative difference of the outlet between the sharded and the whole run, the device
time of both, and the number of reductions the engine asked for."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=4000)
    ap.add_argument("--sets", type=int, default=2)
    ap.add_argument("--years", type=int, default=8)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from hydrobricks_b200 import models, parallel

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, U, T = args.sets, args.units, 365 * args.years + 30
    units = sorted(bench.hydro_units(U, glacier=True, area=1.0e4), key=lambda u: -u["elevation"])
    precip, temp, pet = bench.station_series(T)
    rng = np.random.default_rng(5)
    ps = {"a_snow": rng.uniform(2, 8, P), "A": rng.uniform(50, 800, P), "k_slow_1": rng.uniform(0.01, 0.2, P),
          "beta": rng.uniform(300, 9000, P), "a_ice": rng.uniform(9, 25, P), "k_snow": rng.uniform(0.05, 0.6, P),
          "k_ice": rng.uniform(0.05, 0.6, P)}
    kw = dict(solver="rk4", soil_storage_nb=1, surface_runoff="socont_runoff", land_cover_names=["ground", "glacier"],
              land_cover_types=["ground", "glacier"], glacier_infinite_storage=False, backend="python")
    end = str(np.datetime64(bench.START) + np.timedelta64(T - 1, "D"))
    gl, areas, volumes = bench.c4_tables(units)
    volumes = volumes * 0.1
    table_ids = np.array([units[i]["id"] for i in gl], dtype=np.int32)
    changes = bench.c4_changes(units, gl, areas, T)
    sp = bench.SPATIAL
    keep = []

    def make(block):
        m = models.Socont(**kw)
        m.setup(block, bench.START, end, spinup_days=200, device=local)
        hb = m.backend
        s2i = hb.ActionGlacierSnowToIceTransformation(9, 30, "glacier")
        dh = hb.ActionGlacierEvolutionDeltaH()
        dh.add_lookup_tables(10, "glacier", table_ids, areas, volumes)
        lc = hb.ActionLandCoverChange()
        for date, unit_id, area in changes:
            lc.add_change(date, unit_id, "glacier", area)
        keep.extend([s2i, dh, lc])
        for a in (s2i, dh, lc):
            assert m.model.add_action(a)
        return m

    def load(m):
        m.model.set_parameter_sets(m.parameter_matrix(ps, P))
        m.model.set_station_forcing("precipitation", precip, sp["zref"], sp["grad_p"], 1.0)
        m.model.set_station_forcing("temperature", temp, sp["zref"], sp["grad_t"], 1.0)
        m.model.set_station_forcing("pet", pet)

    calls = [0]
    inner = parallel.make_reduce_callback()

    def counted(ptr, n):
        calls[0] += 1
        inner(ptr, n)

    lo, hi = parallel.shard_bounds(U, rank, world)
    m = make(units[lo:hi])
    m.model.set_unit_shard(parallel.basin_area([u["area"] for u in units]), counted)
    load(m)
    dist.barrier()
    m.model.run()  # warm-up (allocations, NCCL communicator)
    calls[0] = 0
    dist.barrier()
    m.model.run()
    ms_shard = torch.tensor([m.engine(P).last_run_stats()["kernel_ms"]], device="cuda")
    dist.all_reduce(ms_shard, op=dist.ReduceOp.MAX)
    q = m.model.get_outlet_discharge_batch()
    if rank == 0:
        whole = make(units)
        load(whole)
        whole.model.run()
        whole.model.run()
        qw = whole.model.get_outlet_discharge_batch()
        rel = float(np.max(np.abs(q - qw) / np.maximum(np.abs(qw), 1e-2)))
        print(json.dumps({"n_gpus": world, "units": U, "sets": P, "timesteps": T, "max_rel_diff_vs_whole_basin": rel,
                          "sharded_ms_max_over_ranks": float(ms_shard.item()),
                          "whole_basin_ms_one_gpu": whole.engine(P).last_run_stats()["kernel_ms"],
                          "reductions_per_run": calls[0],
                          "launches_per_run": m.engine(P).last_run_stats()["launches"]}), flush=True)
        assert rel < 1e-10, rel
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
