"""Wavenumber sharding of the spectral reduction across ranks (BASELINE north_star:
"sharding wavenumber ranges for the reduction, NCCL allreduce over NVLink only for the small
per-g-point partial accumulators that straddle shards").

Every rank owns a contiguous wavenumber range and its own g-point plan.  Per reduction the
ranks exchange only
    sums   [nlay][nspec][ng][3] = {sum f(tau)*w, sum w, count(tau <= 0)}   all-reduce SUM
    minmax [nlay][nspec][ng][2] = {min tau, -max tau}                      all-reduce MIN
and, once per g-point map, the per-g wavenumber counts (SUM).  The arrays can be numpy arrays
(gloo, used by the CPU tests of this host-side logic) or torch CUDA tensors (NCCL).
"""
from __future__ import annotations

from typing import Tuple

CHUNK = 480  # shard boundaries are multiples of the longest kernel chunk (spectral_reduce.cuh: CH_MAX), hence 16-byte aligned


def shard_range(nwav: int, rank: int, world: int, align: int = CHUNK) -> Tuple[int, int]:
    """[begin, end) of this rank's wavenumbers: contiguous, chunk-aligned, covering [0, nwav)
    exactly once over the ranks, sizes differing by at most one chunk."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    nchunks = (nwav + align - 1) // align
    c0 = (nchunks * rank) // world
    c1 = (nchunks * (rank + 1)) // world
    return min(c0 * align, nwav), min(c1 * align, nwav)


def segment_ranges(nwav: int, nseg: int, align: int = CHUNK):
    """The nseg wavenumber segments [begin, end) of the N-independent decomposition (chunk-aligned
    boundaries that depend on nwav and nseg only)."""
    return [shard_range(nwav, s, nseg, align) for s in range(nseg)]


def segment_owners(nseg: int, world: int):
    """seg_begin[world + 1]: rank r owns segments [seg_begin[r], seg_begin[r + 1]) - contiguous, in rank order,
    counts differing by at most one (ranks may own none when world > nseg)."""
    return [(nseg * r) // world for r in range(world + 1)]


def exchange_partials_reference(seg_sums, seg_minmax, seg_begin, rank, group=None):
    """What ecckd_average_optical_depth_exchange_dev does, on torch CPU tensors over any backend (the gloo tests
    of the host logic): every owner broadcasts its run of slots, then the slots are combined in segment order."""
    import torch
    import torch.distributed as dist

    world = len(seg_begin) - 1
    if world > 1:
        for r in range(world):
            a, b = seg_begin[r], seg_begin[r + 1]
            if b > a:
                dist.broadcast(seg_sums[a:b], src=r, group=group)
                dist.broadcast(seg_minmax[a:b], src=r, group=group)
    sums = seg_sums[0].clone()
    minmax = seg_minmax[0].clone()
    for s in range(1, seg_sums.shape[0]):
        sums += seg_sums[s]
        minmax = torch.minimum(minmax, seg_minmax[s])
    return sums, minmax


def allreduce_partials(sums, minmax, group=None):
    """In-place all-reduce of the partial accumulators (torch tensors, any backend)."""
    import torch.distributed as dist

    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(minmax, op=dist.ReduceOp.MIN, group=group)
    return sums, minmax


def allreduce_counts(count, group=None):
    import torch.distributed as dist

    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return count


def shard_columns(ncol: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of this rank's columns for the flux solver (independent columns: no exchange)."""
    return (ncol * rank) // world, (ncol * (rank + 1)) // world


# ------------------------------------------------------------------ sort keys + ranking
def allgather_wavenumber_shards(local, nwav: int, rank: int, world: int, group=None, align: int = CHUNK):
    """All-gather of a per-wavenumber torch vector whose rank-r piece covers shard_range(nwav, r, world):
    every rank gets the full [nwav] vector (the sort keys of find_g_points: each rank computes the
    keys of its range, the stable radix sort is then replicated on every rank - deterministic and
    bit-exact, no exchange inside the sort).  Pieces differ by at most one chunk, so they are padded
    to the longest before the collective (NCCL and gloo want equal sizes)."""
    import torch
    import torch.distributed as dist

    spans = [shard_range(nwav, r, world, align) for r in range(world)]
    lo, hi = spans[rank]
    if local.numel() != hi - lo:
        raise ValueError("local piece does not match this rank's wavenumber range")
    if world == 1:
        return local.clone()
    longest = max(b - a for a, b in spans)
    padded = torch.zeros(longest, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(pieces, spans)])


def sharded_sort_keys(key_fn, nwav: int, rank: int, world: int, group=None):
    """key_fn(lo, hi) -> torch vector of the sort keys of wavenumbers [lo, hi) (e.g. a call of
    ecckd_sort_key_lw_dev on the column slice od[:, lo:hi] with od_stride = nwav).  Returns the full
    key vector on every rank; rank it with rank_bands_dev (replicated)."""
    lo, hi = shard_range(nwav, rank, world)
    return allgather_wavenumber_shards(key_fn(lo, hi), nwav, rank, world, group)


# ------------------------------------------------------------------ forward cost
def sharded_forward_cost(cost_fn, nprof: int, rank: int, world: int, per_profile: dict, shared: dict, group=None,
                         device=None):
    """optimize_lut's forward cost of one scene with its profiles (columns) sharded over the ranks:
    cost_fn(**shared, **{name: array[lo:hi]}) is api.cost_function_ckd_lw / _sw (or, in the CPU tests
    of this logic, the oracle) on this rank's profiles; the only exchange is ONE all-reduce of
    {cost_total, penalty} (both are plain sums over profiles: solve_adept.cpp:107-116,157,183).
    `per_profile` holds the arguments whose leading dimension is the profile, `shared` the rest
    (None entries pass through).  Returns dict(cost_total, penalty, cost_profile=<this rank's>, range=(lo, hi))."""
    import torch
    import torch.distributed as dist

    lo, hi = shard_columns(nprof, rank, world)
    acc = torch.zeros(2, dtype=torch.float64, device=device if device is not None else "cpu")
    local = None
    if hi > lo:
        kw = dict(shared)
        for name, a in per_profile.items():
            kw[name] = None if a is None else a[lo:hi]
        local = cost_fn(**kw)
        acc[0] = float(local["cost_total"])
        acc[1] = float(local["penalty"])
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    tot = acc.cpu().numpy()
    return dict(cost_total=float(tot[0]), penalty=float(tot[1]),
                cost_profile=None if local is None else local["cost_profile"], range=(lo, hi))
