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


# ------------------------------------------------------------------ sort keys and forward cost
def _oracle_cost_fn(O, weights):
    """The oracle standing in for api.cost_function_ckd_lw on a shard of profiles."""
    def fn(pressure_hl, planck_hl, surf_emissivity, surf_planck, optical_depth, flux_dn, flux_up, hr):
        costs = []
        for i in range(pressure_hl.shape[0]):
            lw = np.diff(np.sqrt(pressure_hl[i]))
            costs.append(O.calc_cost_function_ckd_lw(pressure_hl[i], planck_hl[i], surf_emissivity[i], surf_planck[i],
                                                     np.maximum(optical_depth[i], 0.0), flux_dn[i], flux_up[i], hr[i],
                                                     *weights, lw / lw.sum()))
        pen = 1.0e4 * float((np.minimum(optical_depth, 0.0) ** 2).sum())
        return dict(cost_profile=np.array(costs), penalty=pen, cost_total=pen + float(np.sum(costs)))
    return fn


def _cost_scene(O, nprof=5, nlay=12, ng=8):
    from tests.test_reference_sources import ckd_cost_case

    cases = [ckd_cost_case(nlay=nlay, ng=ng, nband=ng, seed=900 + i) for i in range(nprof)]
    rng = np.random.default_rng(4)
    s = dict(pressure_hl=np.stack([c["p"] for c in cases]), planck_hl=np.stack([c["planck"] for c in cases]),
             surf_planck=np.stack([c["surf"] for c in cases]), optical_depth=np.stack([c["od"] for c in cases]),
             surf_emissivity=rng.uniform(0.9, 1.0, (nprof, ng)))
    s["optical_depth"][1, 2, 3] = -0.02
    s["optical_depth"][nprof - 1, 0, 0] = -0.3
    dn, up, hr = [], [], []
    for i in range(nprof):
        a, b = O.radiative_transfer_lw(s["planck_hl"][i], np.maximum(s["optical_depth"][i], 0) * 1.1,
                                       s["surf_emissivity"][i], s["surf_planck"][i])
        dn.append(a), up.append(b), hr.append(O.heating_rate(s["pressure_hl"][i], a, b))
    s.update(flux_dn=np.stack(dn), flux_up=np.stack(up), hr=np.stack(hr))
    return s


def _worker2(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.oracle_lib import oracle

        O = oracle()
        # (1) sort keys: each rank computes its wavenumber range, all-gather, replicated stable rank
        c = S.lw_spectrum_case(nlay=9, nwav=5003, ng=4, nband=1, seed=8, zero_frac=0.0, vmr=1.0)
        od = c["od"] * 3e4
        key_full, _ = O.sort_key_lw(c["pressure_hl"], c["wavenumber"], c["d_wavenumber"], od, 0.5)

        def key_fn(lo, hi):
            k, _ = O.sort_key_lw(c["pressure_hl"], c["wavenumber"][lo:hi], c["d_wavenumber"][lo:hi],
                                 np.ascontiguousarray(od[:, lo:hi]), 0.5)
            return torch.from_numpy(k)

        key = sharding.sharded_sort_keys(key_fn, c["nwav"], rank, world).numpy()
        b1, b2 = [c["wavenumber"][0], c["wavenumber"][2000]], [c["wavenumber"][2000], c["wavenumber"][-1]]
        same_key = bool(np.array_equal(key, key_full, equal_nan=True))
        same_rank = all(np.array_equal(a, b) for a, b in zip(O.rank_bands(c["wavenumber"], key, b1, b2),
                                                             O.rank_bands(c["wavenumber"], key_full, b1, b2)))
        # (2) forward cost: 5 profiles over 2 ranks (2 + 3), one all-reduce of {cost_total, penalty}
        weights = (0.2, 0.2, 0.8, 0.1)
        fn = _oracle_cost_fn(O, weights)
        s = _cost_scene(O)
        got = sharding.sharded_forward_cost(fn, 5, rank, world, per_profile=s, shared={})
        ref = fn(**s)
        lo, hi = got["range"]
        ok_prof = bool(np.array_equal(got["cost_profile"], ref["cost_profile"][lo:hi]))
        if rank == 0:
            q.put((same_key, same_rank, ok_prof, abs(got["cost_total"] - ref["cost_total"]) / ref["cost_total"],
                   abs(got["penalty"] - ref["penalty"]) / ref["penalty"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sort_keys_and_forward_cost_match_unsharded():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker2, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    same_key, same_rank, ok_prof, d_cost, d_pen = q.get(timeout=10)
    assert same_key and same_rank and ok_prof
    assert d_cost < 1e-13 and d_pen < 1e-13


# ------------------------------------------------------------------ N-independent segments
def test_segment_ranges_and_owners():
    for nwav in (1, 5003, 8_388_608):
        for nseg in (1, 6, 8):
            spans = sharding.segment_ranges(nwav, nseg)
            assert spans[0][0] == 0 and spans[-1][1] == nwav
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for nseg in (1, 6, 8):
        for world in (1, 2, 3, 8, 11):
            sb = sharding.segment_owners(nseg, world)
            assert sb[0] == 0 and sb[-1] == nseg and len(sb) == world + 1
            assert all(0 <= b - a <= (nseg + world - 1) // world for a, b in zip(sb, sb[1:]))


def _worker_segments(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_np as N

        nseg, kd = 6, 1.66
        c = S.lw_spectrum_case(nlay=7, nwav=5003, ng=9, nband=2, seed=3)
        c["planck"] = N.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
        spans = sharding.segment_ranges(c["nwav"], nseg)
        sb = sharding.segment_owners(nseg, world)
        seg_s = torch.zeros((nseg, c["nlay"], 1, c["ng"], 3), dtype=torch.float64)
        seg_m = torch.zeros((nseg, c["nlay"], 1, c["ng"], 2), dtype=torch.float64)
        for s in range(sb[rank], sb[rank + 1]):  # this rank's segments only
            a, b, _ = _partials(c, spans[s][0], spans[s][1], kd)
            seg_s[s], seg_m[s] = torch.from_numpy(a), torch.from_numpy(b)
        sums, minmax = sharding.exchange_partials_reference(seg_s, seg_m, sb, rank)
        count = np.bincount(c["g_point"][c["g_point"] >= 0], minlength=c["ng"]).astype(np.int64)
        out = _finalize(c, sums.numpy(), minmax.numpy(), count, kd)
        q.put((rank, out.tobytes()))
    finally:
        dist.destroy_process_group()


def test_segment_exchange_is_bitwise_independent_of_the_number_of_ranks():
    """One, two and three gloo ranks own 6, 3+3 and 2+2+2 of the same six segments: after the exchange and the
    fixed-order combination every rank of every world size holds the same bits."""
    import torch.multiprocessing as mp

    results = {}
    for world in (1, 2, 3):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker_segments, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=240) for _ in range(world)]
        for p in procs:
            p.join(timeout=240)
            assert p.exitcode == 0
        results[world] = dict(got)
    base = results[1][0]
    for world, by_rank in results.items():
        for rank, blob in by_rank.items():
            assert blob == base, (world, rank)
