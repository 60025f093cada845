"""Multi-GPU plumbing (one process per GPU, torch.distributed), SURVEY.md §8e.

* Ensembles: parameter sets are independent, so the parameter axis is sharded across ranks with NO collective on the
  hot path; the only exchange is the final gather of per-set results (scores [P] or outlet series [P x T]).
* Large distributed basins: blocks of hydro units are sharded across ranks (`ShardedBasin`); every rank integrates its
  block, then ONE all-reduce(sum) of the area-weighted partial outlet [P x T] and of the two per-step deposit sums into
  the sub-basin stores, after which the two linear sub-basin stores are integrated on the summed deposits (exactly — no
  superposition across shards, the 1e-8 thresholds of the reference are not linear).
NCCL on GPUs, gloo on CPU in the tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_sets: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of the parameter axis: the first (n_sets % world) ranks get one extra set."""
    base, extra = divmod(int(n_sets), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# Relative cost of a hydro unit that carries ice (the glacier kernel variant, 128 registers) against a unit that never does
# (the plain variant): 51.5 ms / 35.3 ms on the same 1 M units x 730 steps (profiles/r02s_kbench.jsonl)
GLACIER_UNIT_COST = 1.46


def unit_costs(units):
    """Relative integration cost per hydro unit, for a cost-balanced split of a basin over the ranks."""
    return [GLACIER_UNIT_COST if any(abs(f) > 0 for n, f in u["land_covers"] if "glacier" in n) else 1.0 for u in units]


def weighted_shard_bounds(weights, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split with (nearly) equal summed weight per rank: boundary r sits where the prefix sum of the weights
    first reaches r/world of the total. A basin sorted by elevation has its glacier units — the expensive ones — at
    one end, so equal unit counts leave the rank with the glaciers 40 % behind (2xB200, C3: x1.58 instead of x1.85)."""
    import numpy as np

    w = np.asarray(weights, dtype=np.float64)
    if len(w) == 0 or world <= 1:
        return 0, len(w)
    csum = np.cumsum(w)
    cuts = [0] + [int(np.searchsorted(csum, csum[-1] * r / world, side="left")) for r in range(1, world)] + [len(w)]
    for i in range(1, len(cuts)):  # monotone, and no empty block while units are left
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


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
        # Socont.engine() gives the ctypes view of the hbk_engine for either host backend (native C++ module or mirror)
        scores = score_fn(device_outlet(self.model.engine(len(local))))
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


def basin_area(areas) -> float:
    """Area of the whole basin as SubBasin::AddHydroUnit accumulates it (SubBasin.cpp:173): a running float64 sum
    in unit order — NOT numpy's pairwise sum, the outlet weights area/basinArea must be those of the full model."""
    a = np.asarray(areas, dtype=np.float64)
    return float(np.cumsum(a)[-1]) if a.size else 0.0  # cumsum is the plain running sum (np.sum is pairwise)


def reduce_partials(partials, group=None):
    """In-place sum over the ranks of the partial series (torch tensors), one collective per tensor."""
    import torch.distributed as dist

    for t in partials:
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return partials


def make_reduce_callback(group=None):
    """The in-place sum over the ranks that a unit shard asks for inside its run (hbk_set_shard_reduce,
    Engine.set_shard_reduce): one all-reduce(sum) of n float64 at a device address, complete when it returns."""
    import torch
    import torch.distributed as dist

    def reduce(ptr, n):
        t = torch.as_tensor(_DeviceArray(ptr, (n,), None), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.current_stream().synchronize()

    return reduce


def shard_partials(engine):
    """Zero-copy torch views [n_sets x ld] of a unit shard's partial outlet and deposit sums in HBM."""
    import torch

    q, d1, d2, n, ld = engine.shard_partials_device_pointers()
    view = lambda ptr: None if not ptr else torch.as_tensor(  # noqa: E731
        _DeviceArray(ptr, (n, ld), (ld * 8, 8)), device="cuda")
    return [view(q), view(d1), view(d2)]


class ShardedBasin:
    """One large basin split into contiguous blocks of hydro units, one block per rank.

    `make_model(units_block)` builds and sets up a hydrobricks_b200.models.Socont for a block of units on this rank's
    device (same structure, time axis and forcing on every rank); `units` is the full list in basin order
    (descending elevation, as HydroUnits sorts them). run() integrates the block and returns nothing; afterwards
    get_outlet_discharge() gives the outlet of the WHOLE basin on every rank.

    in_run_reduce=True hands the engine a reduction callback (hbk_set_shard_reduce): the sums over the ranks then happen
    inside the run — which is what spin-up (ModelHydro::RunSpinup) and dated actions (the delta-h water equivalent of the
    whole glacier at every update) need — and run() goes through the model's own run(), so that actions registered on
    the model (tables / changes of the WHOLE basin, the same on every rank) are scheduled. Needs the Python host.

    balance_cost=True cuts the blocks by integration cost (units with ice cost GLACIER_UNIT_COST) instead of unit count.
    """

    def __init__(self, make_model, units, group=None, in_run_reduce=False, balance_cost=False):
        import torch.distributed as dist

        self.group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        units = sorted(units, key=lambda u: -u["elevation"])
        lo, hi = weighted_shard_bounds(unit_costs(units), rank, world) if balance_cost else \
            shard_bounds(len(units), rank, world)
        self.bounds = (lo, hi)
        self.total_area = basin_area([u["area"] for u in units])
        self.model = make_model(units[lo:hi])
        self._engine = None
        self.in_run_reduce = in_run_reduce
        if in_run_reduce:
            self.model.model.set_unit_shard(self.total_area, make_reduce_callback(group))

    def engine(self, n_sets=1):
        if self._engine is None:
            self._engine = self.model.engine(n_sets)
            self._engine.set_unit_shard(self.total_area)
        self._engine.n_sets = n_sets
        return self._engine

    def run(self, n_sets=1):
        if self.in_run_reduce:
            self._engine = self.model.engine(n_sets)
            self.model.model.run()
            return
        eng = self.engine(n_sets)
        import torch

        eng.run()  # returns after the engine's stream has drained
        reduce_partials(shard_partials(eng), self.group)
        torch.cuda.current_stream().synchronize()  # the collective runs on torch's stream, the engine on its own
        eng.finish_sharded_run()

    def get_outlet_discharge(self):
        return self._engine.get_outlet()
