"""The N-independent form of the sharded reduction: the wavenumber axis is cut into
segments that do not depend on the number of ranks, every segment is reduced with its own plan into its own
slot, ecckd_average_optical_depth_exchange_dev combines the slots in segment order.  One GPU plays the ranks of
N = 1, 2 and 3 in turn (every "rank" holds ONLY its own wavenumber range, in its own allocations, at its own
alignment); the finalized look-up-table slices must be bitwise identical for every N and match the oracle."""
import numpy as np
import pytest

import ecckd_b200 as E
from ecckd_b200 import sharding
from tests import synthetic as S

pytestmark = pytest.mark.gpu

NSEG = 6


def _run_as(ctx, c, ods, vmrs, method, world):
    nlay, ng, nwav, nspec = c["nlay"], c["ng"], c["nwav"], len(ods)
    spans = sharding.segment_ranges(nwav, NSEG)
    seg_begin = sharding.segment_owners(NSEG, world)
    n3, n2 = nlay * nspec * ng * 3, nlay * nspec * ng * 2
    seg_sums, seg_minmax = ctx.zeros((NSEG, n3)), ctx.zeros((NSEG, n2))
    seg_count = ctx.zeros((NSEG, ng), np.int64)
    keep = []
    for rank in range(world):
        s0, s1 = seg_begin[rank], seg_begin[rank + 1]
        if s1 == s0:
            continue
        lo, hi = spans[s0][0], spans[s1 - 1][1]
        pad = rank % 2  # an odd number of leading elements: this rank's rows start 8 bytes off a 16-byte boundary

        def own(a, dtype=np.float64):
            host = np.zeros(pad + hi - lo, dtype=dtype)
            host[pad:] = a[lo:hi]
            return ctx.to_device(host, dtype)

        d_g, d_wn, d_dwn = own(c["g_point"], np.int32), own(c["wavenumber"]), own(c["d_wavenumber"])
        d_ods = []
        for od in ods:
            host = np.zeros((nlay, pad + hi - lo))
            host[:, pad:] = od[:, lo:hi]
            d_ods.append(ctx.to_device(host))
        keep += [d_g, d_wn, d_dwn] + d_ods
        for s in range(s0, s1):
            a, b = spans[s][0] - lo + pad, spans[s][1] - lo + pad
            if b == a:
                continue
            plan = E.GPointPlan(ctx, d_g.offset(a), b - a, ng)
            seg_count.row(s).copy_from(plan.counts())
            E.average_optical_depth_partial_dev(ctx, plan, nlay, c["pressure_fl"], [d.offset(a) for d in d_ods],
                                                pad + hi - lo, E.WEIGHT_PLANCK, None, 0, c["temperature_fl"],
                                                d_wn.offset(a), d_dwn.offset(a), method, seg_sums.row(s), seg_minmax.row(s))
            keep.append(plan)
    sums, minmax, count = ctx.empty(n3), ctx.empty(n2), ctx.empty(ng, np.int64)
    E.average_optical_depth_exchange_dev(ctx, None, seg_begin, seg_sums, seg_minmax, sums, minmax, seg_count, count)
    outs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    mins = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    maxs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    E.average_optical_depth_finalize_dev(ctx, ng, count, vmrs, c["pressure_fl"], c["pressure_hl"], method, sums, minmax,
                                         outs, mins, maxs, ng)
    res = [(o.numpy(), a.numpy(), b.numpy()) for o, a, b in zip(outs, mins, maxs)]
    for p in keep:
        if isinstance(p, E.GPointPlan):
            p.close()
    return res, count.numpy()


@pytest.mark.parametrize("method", ["transmission", "linear", "logarithmic"])
def test_segmented_reduction_is_bitwise_the_same_for_one_two_and_three_ranks(ctx, O, method):
    c = S.lw_spectrum_case(nlay=6, nwav=40013, ng=14, nband=3, seed=17)
    rng = np.random.default_rng(2)
    ods = [c["od"], c["od"] * rng.uniform(0.2, 4.0)]
    vmrs = [c["vmr"], c["vmr"] * 3.0]
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    base, count = _run_as(ctx, c, ods, vmrs, method, 1)
    assert np.array_equal(count, np.bincount(c["g_point"][c["g_point"] >= 0], minlength=c["ng"]))
    for s, od in enumerate(ods):
        ref = O.average_optical_depth_to_g_point(c["ng"], vmrs[s], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                                 planck, method)
        S.assert_avg_close(base[s] + (None,), ref, method=method, scale=S.out_scale(dict(c, vmr=vmrs[s])),
                           what=f"segments, {method}, spectrum {s}: ")
    for world in (2, 3):
        got, cnt = _run_as(ctx, c, ods, vmrs, method, world)
        assert np.array_equal(cnt, count)
        for s in range(len(ods)):
            for a, b in zip(got[s], base[s]):
                assert np.array_equal(a, b), f"{world} ranks differ from 1 rank ({method}, spectrum {s})"
