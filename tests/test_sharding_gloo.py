"""The N > 1 path on CPU: two gloo ranks shard the wavenumber axis, compute their partial
accumulators (with the NumPy oracle standing in for the device kernel), all-reduce them exactly
as bench.py does over NCCL, finalize, and must reproduce the unsharded oracle result."""
import os
import socket

import numpy as np
import pytest

from ecckd_b200 import sharding
from tests import synthetic as S


def test_shard_range_partitions_exactly():
    for nwav in (1, 223, 224, 225, 10_000, 8_388_608):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(nwav, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nwav
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[0] % sharding.CHUNK == 0 for s in spans)
    assert [sharding.shard_columns(50, r, 8) for r in range(8)][-1][1] == 50


def _partials(c, lo, hi, kd):
    """{sum f w, sum w, zero count}, {min, -max} per (layer, g) for wavenumbers [lo, hi)."""
    nlay, ng = c["nlay"], c["ng"]
    sums = np.zeros((nlay, 1, ng, 3))
    minmax = np.full((nlay, 1, ng, 2), np.inf)
    g = c["g_point"][lo:hi]
    od = c["od"][:, lo:hi]
    w = c["planck"][:, lo:hi]
    for ig in range(ng):
        m = g == ig
        if not m.any():
            continue
        sums[:, 0, ig, 0] = np.sum((1.0 - np.exp(-od[:, m] * kd)) * w[:, m], axis=1)
        sums[:, 0, ig, 1] = np.sum(w[:, m], axis=1)
        minmax[:, 0, ig, 0] = od[:, m].min(axis=1)
        minmax[:, 0, ig, 1] = -od[:, m].max(axis=1)
    count = np.bincount(g[g >= 0], minlength=ng).astype(np.int64)
    return sums, minmax, count


def _finalize(c, sums, minmax, count, kd):
    """average_optical_depth.cpp:47-52,135-193 on the reduced accumulators."""
    s0, s1 = sums[:, 0, :, 0], sums[:, 0, :, 1]
    mn, mx = minmax[:, 0, :, 0], -minmax[:, 0, :, 1]
    with np.errstate(all="ignore"):
        fit = np.abs(-np.log(1.0 - np.minimum(0.9999999999999999, s0 / s1)) / kd)
    fit = np.maximum(mn, np.minimum(fit, mx))
    fit[:, count == 0] = 0.0
    return fit * S.out_scale(c)[:, None]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_np as N

        c = S.lw_spectrum_case(nlay=7, nwav=5003, ng=9, nband=2, seed=3)
        c["planck"] = N.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
        kd = 1.66
        lo, hi = sharding.shard_range(c["nwav"], rank, world)
        sums, minmax, count = _partials(c, lo, hi, kd)
        t_s, t_m, t_c = torch.from_numpy(sums), torch.from_numpy(minmax), torch.from_numpy(count)
        sharding.allreduce_partials(t_s, t_m)
        sharding.allreduce_counts(t_c)
        out = _finalize(c, t_s.numpy(), t_m.numpy(), t_c.numpy(), kd)
        if rank == 0:
            ref = N.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"],
                                                     c["g_point"], c["od"], c["planck"], "transmission")[0]
            q.put(float(np.max(np.abs(out - ref) / np.abs(ref))))
    finally:
        dist.destroy_process_group()


def test_two_rank_wavenumber_sharding_matches_unsharded():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) < 1e-12
