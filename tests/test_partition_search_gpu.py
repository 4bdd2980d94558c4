"""Config D, second half: the partition search of find_g_points on the GPU evaluator
(ecckd_partition_search_* over ecckd_interval_error_{lw,sw}_eval = batched CkdEquipartition::calc_error)
against the reference's own search - Equipartition::equipartition_e / _n driven by the reference's
CkdEquipartition::calc_error (find_g_points.cpp:291-405, equipartition.cpp:201-805), compiled in place into
oracle/_ref/libecckd_ref_tools.so.

The GPU errors equal the reference's to ~1e-13, not bit for bit, so the comparison is made in two steps:
  (1) the driver on the GPU evaluator equals, bit for bit, the REFERENCE's Equipartition class driven by the
      same GPU evaluator through a callback (control flow);
  (2) against the all-CPU reference search: the same number of g points and status, bounds that snap to the
      same rank1 / rank2 (find_g_points.cpp:1396-1400) and errors within COST_RTOL."""
import numpy as np
import pytest

import ecckd_b200 as E
from ecckd_b200 import api as A
from tests import ref_tools as RT
from tests.test_reference_sources import interval_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RT.available(), reason="oracle/_ref/libecckd_ref_tools.so not built")]

COST_RTOL = 1e-10


@pytest.fixture(scope="module")
def R():
    return RT.ref_tools()


def lw_band(O, n, nlay, seed, method):
    """A band in rank order: columns sorted by their total optical depth, as reorder_spectrum leaves them."""
    c = interval_case(nlay=nlay, n=n, seed=seed)
    order = np.argsort(c["od"].sum(axis=0), kind="stable")
    for k in ("od", "planck"):
        c[k] = np.ascontiguousarray(c[k][:, order])
    for k in ("emis", "surf"):
        c[k] = np.ascontiguousarray(c[k][order])
    od_gas = c["od"] * c["rng"].uniform(0.5, 1.5, c["od"].shape)
    true_dn, true_up = O.radiative_transfer_lw(c["planck"], c["od"] + od_gas, c["emis"], c["surf"])
    hr = O.heating_rate(c["p"], true_dn, true_up)
    if method in ("transmission", "transmission-2"):
        metric = 1.0 - np.exp(-1.66 * (2.0 if method.endswith("2") else 1.0) * od_gas)
    elif method == "square-root":
        metric = np.sqrt(od_gas)
    else:
        metric = od_gas
    c.update(metric=metric, dn=np.ascontiguousarray(true_dn[-1]), up=np.ascontiguousarray(true_up[0]), hr=hr)
    return c


def snapped(bounds, npoints):
    r1 = np.ceil(bounds[:-1] * (npoints - 1)).astype(np.int64)
    r2 = np.floor(bounds[1:] * (npoints - 1)).astype(np.int64)
    return r1, r2


@pytest.mark.parametrize("method,n,offset", [("linear", 6000, 0), ("transmission", 20000, 0), ("logarithmic", 9000, 2500)])
def test_partition_search_lw_equals_the_reference_search(ctx, O, R, method, n, offset):
    nlay, fw = 16, 0.2
    c = lw_band(O, n, nlay, 11, method)
    nband = n - offset  # the band is [offset, n - 1] of the rank axis (ibegin..iend)
    d = {k: ctx.to_device(c[k]) for k in ("p", "planck", "emis", "surf", "od", "metric", "dn", "up", "hr", "lw")}
    ev = A.interval_error_lw_evaluator(ctx, method, nlay, nband, n, d["p"], d["planck"], d["emis"], d["surf"], d["od"],
                                       d["metric"], d["dn"], d["up"], d["hr"], fw, d["lw"], offset=offset)
    band = R.band_lw(method, fw, c["lw"], c["p"], c["emis"], c["surf"], c["dn"], c["up"], c["planck"], c["od"],
                     c["metric"], c["hr"], ibegin=offset, iend=n - 1)
    # a target that needs of the order of ten g points
    whole = R.calc_error_lw(band, [0.0], [1.0])[0]
    target = 0.02 * whole
    opt = dict(partition_max_iterations=60, partition_tolerance=0.02, resolution=1.0 / nband)

    ps = A.PartitionSearch(ev, speculative_line_search=4, **opt)
    st, b, e = ps.equipartition_e(target, 0.0, 1.0, capacity=512)
    stats = ps.stats()
    assert 3 <= e.size <= 200, e.size

    # (1) the reference's Equipartition on the same GPU evaluator: bit for bit
    L = A._lib.lib()

    def gpu_error(b1, b2):
        x1, x2, out = np.array([b1]), np.array([b2]), np.zeros(1)
        rc = L.ecckd_interval_error_lw_eval(A.C.addressof(ev[1]), 1, x1.ctypes.data, x2.ctypes.data, out.ctypes.data)
        assert rc == 0
        return out[0]
    st_r, b_r, e_r, calls = R.equipartition_e_callback(gpu_error, target, 0.0, 1.0, **opt)
    assert st == A.EP_STATUS[st_r]
    assert np.array_equal(b, b_r) and np.array_equal(e, e_r)
    assert stats["batches"] < calls  # the batches did the work of `calls` single evaluations

    # (2) the all-CPU reference search
    st_c, b_c, e_c = R.equipartition_e_lw(band, target, max_iterations=60, tolerance_tolerance=0.02)
    assert A.EP_STATUS[st_c] == st and b_c.size == b.size
    assert np.array_equal(snapped(b, nband), snapped(b_c, nband))
    np.testing.assert_allclose(e, e_c, rtol=COST_RTOL)

    # the band call: min / max g points and rank1 / rank2
    st2, b2, e2, r1, r2 = ps.band(nband, target, min_g_points=e.size + 3, max_g_points=e.size + 5)
    assert e2.size == e.size + 3
    st_n, b_n, e_n = R.equipartition_n_lw(band, e.size + 3, max_iterations=60, tolerance_tolerance=0.02)
    assert st2 == A.EP_STATUS[st_n]
    s1, s2 = snapped(b_n, nband)
    assert np.array_equal(r1, s1) and np.array_equal(r2, s2)
    np.testing.assert_allclose(e2, e_n, rtol=COST_RTOL)
    ps.close()


def test_partition_search_sw_equals_the_reference_search(ctx, O, R):
    n, nlay, fw, cos_sza, albedo = 12000, 16, 0.2, 0.5, 0.15
    c = interval_case(nlay=nlay, n=n, seed=21)
    rng = c["rng"]
    order = np.argsort(c["od"].sum(axis=0), kind="stable")
    od = np.ascontiguousarray(c["od"][:, order]) * 0.05
    ssi = rng.uniform(0.1, 2.0, n) * 1e-3
    od_gas = od * rng.uniform(0.5, 1.5, od.shape)
    true_dn, true_up = O.radiative_transfer_norayleigh_sw(cos_sza, ssi, od + od_gas, np.full(n, albedo))
    hr = O.heating_rate(c["p"], true_dn, None)
    metric = 1.0 - np.exp(-1.66 * od_gas)
    dn, up = np.ascontiguousarray(true_dn[-1]), np.ascontiguousarray(true_up[0])
    d = [ctx.to_device(a) for a in (c["p"], ssi, od, metric, dn, up, hr, c["lw"])]
    ev = A.interval_error_sw_evaluator(ctx, "transmission", nlay, n, n, cos_sza, d[0], d[1], albedo, d[2], d[3], d[4], d[5],
                                       d[6], fw, d[7])
    band = R.band_sw("transmission", fw, cos_sza, albedo, c["lw"], c["p"], ssi, dn, up, od, metric, hr)
    whole = R.calc_error_sw(band, [0.0], [1.0])[0]
    opt = dict(partition_max_iterations=60, partition_tolerance=0.02, resolution=1.0 / n)
    ps = A.PartitionSearch(ev, **opt)
    ng = 9
    st, b, e = ps.equipartition_n(np.sqrt(np.arange(ng + 1) / ng))
    L = A._lib.lib()

    def gpu_error(b1, b2):
        x1, x2, out = np.array([b1]), np.array([b2]), np.zeros(1)
        assert L.ecckd_interval_error_sw_eval(A.C.addressof(ev[1]), 1, x1.ctypes.data, x2.ctypes.data, out.ctypes.data) == 0
        return out[0]
    st_r, b_r, e_r, _ = R.equipartition_n_callback(gpu_error, np.sqrt(np.arange(ng + 1) / ng), **opt)
    assert st == A.EP_STATUS[st_r] and np.array_equal(b, b_r) and np.array_equal(e, e_r)
    assert whole > 0 and np.all(e > 0)
    ps.close()
