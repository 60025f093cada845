"""Multi-GPU plumbing (one process per GPU, torch.distributed): parameter sets are independent, so the
parameter axis is sharded across ranks with NO collective on the hot path; the only exchange is
the final gather of per-set results (scores [P] or outlet series [P x T]) over NCCL (gloo on CPU in the tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_sets: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of the parameter axis: the first (n_sets % world) ranks get one extra set."""
    base, extra = divmod(int(n_sets), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(matrix, rank: int, world: int):
    lo, hi = shard_bounds(len(matrix), rank, world)
    return matrix[lo:hi]


def gather_rows(local, n_total: int, group=None):
    """All-gather per-set rows (torch tensor [n_local, ...]) from every rank into [n_total, ...], in set order.
    Shards may differ by one row: pad to the largest shard, gather, drop the padding."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    width = max(sizes)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * width: r * width + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


class ShardedEnsemble:
    """Runs a parameter-set ensemble across the ranks of a process group.

    `model` is a set-up hydrobricks_b200.models.Socont on this rank's device (every rank holds all units and
    the forcing). run(matrix) integrates this rank's shard of `matrix` [n_sets x HBK_N_PARAMS] and returns the
    gathered per-set score tensor [n_sets] computed by `score_fn(outlet_local [n_local x T]) -> [n_local]`.
    """

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def run(self, matrix, score_fn):
        import torch
        import torch.distributed as dist

        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        local = np.ascontiguousarray(shard_rows(matrix, rank, world), dtype=np.float32)
        self.model.model.set_parameter_sets(local)
        self.model.model.run()
        scores = score_fn(device_outlet(self.model.model._engine))
        return gather_rows(scores, len(matrix), self.group)


class _DeviceArray:
    def __init__(self, ptr, shape, strides):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": strides}


def device_outlet(engine):
    """Zero-copy torch view [n_sets x T] of the engine's outlet discharge in HBM."""
    import torch

    ptr, ld = engine.outlet_device_pointer()
    return torch.as_tensor(_DeviceArray(ptr, (engine.n_sets, engine.n_steps), (ld * 8, 8)), device="cuda")
