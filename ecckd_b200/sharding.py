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
