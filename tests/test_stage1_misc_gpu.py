"""GPU parity of the stage-1 companions: planck_function, merge, gpoint_fraction,
ssi intervals, Planck LUT, g-point compaction."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu


def test_planck_function(ctx, O):
    wn = np.linspace(0.0002, 3260.0, 50001)
    dwn = np.full(wn.size, wn[1] - wn[0])
    t = np.array([120.0, 187.3, 250.0, 299.9, 350.0])
    got = E.planck_function(ctx, t, wn, dwn)
    ref = O.planck_function(t, wn, dwn)
    # exp(x) amplifies a last-bit difference in x = (h/k)(nu/T) by x (<= 40 here)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)
    with pytest.raises(E.EcckdError) as ei:  # planck_function.cpp:35-38
        E.planck_function(ctx, t, wn, dwn[:-1])
    assert ei.value.code == 147


def test_merge_axpy(ctx, O):
    rng = np.random.default_rng(0)
    nlay, nwav = 7, 3001
    a, b = rng.uniform(0, 1, (nlay, nwav)), rng.uniform(0, 1, (nlay, nwav))
    prof = rng.uniform(0.5, 2, nlay)
    for scaling, profile in [(1.0, None), (0.37, None), (1.0, prof)]:
        d_a, d_b = ctx.to_device(a), ctx.to_device(b)
        E.merge_axpy_dev(ctx, nlay, nwav, d_a, nwav, d_b, nwav, scaling, profile)
        ref = O.merge_axpy(a, b, scaling, profile)
        np.testing.assert_allclose(d_a.numpy(), ref, rtol=2.3e-16)


@pytest.mark.parametrize("nwav,pad", [(2999, 7), (3000, 7), (5121, 1), (1, 1)])
def test_merge_spectra_one_pass_equals_the_axpy_sequence_bitwise(ctx, O, nwav, pad):
    """read_merged_spectrum.cpp:149-168: the fused merge must give the bits of the loop of AXPYs on a zeroed array
    (and of the oracle's loop), for scalings of 1, below 0 (= 1) and arbitrary ones, with row strides (even strides
    take the two-wavenumbers-per-load kernel, with an odd last wavenumber; odd ones the scalar kernel)."""
    rng = np.random.default_rng(1)
    nlay = 5
    srcs = [rng.uniform(0, 3, (nlay, nwav + pad)) for _ in range(4)]
    scalings = [1.0, 0.37, -1.0, 2.5]
    d_srcs = [ctx.to_device(a) for a in srcs]
    d_out = ctx.zeros((nlay, nwav + 3))
    E.merge_spectra_dev(ctx, nlay, nwav, d_srcs, [nwav + pad] * 4, scalings, d_out, nwav + 3)
    seq = ctx.zeros((nlay, nwav))
    ref = np.zeros((nlay, nwav))
    for a, d_a, sc in zip(srcs, d_srcs, scalings):
        sc = 1.0 if sc < 0.0 else sc
        E.merge_axpy_dev(ctx, nlay, nwav, seq, nwav, d_a, nwav + pad, sc, None)
        ref = O.merge_axpy(ref, np.ascontiguousarray(a[:, :nwav]), sc, None)
    got = d_out.numpy()[:, :nwav]
    assert np.array_equal(got, seq.numpy())
    assert np.array_equal(got, ref)
    assert np.all(d_out.numpy()[:, nwav:] == 0.0)  # nothing written past the row


def test_gpoint_fraction_ssi_intervals_planck_lut(ctx, O):
    c = S.lw_spectrum_case(nlay=3, nwav=30000, ng=14, nband=3, seed=8)
    w1, w2 = O.gpoint_fraction_grid(False, [0.0, 500.0, 1500.0], [500.0, 1500.0, 3260.0])
    d_g, d_wn, d_dwn = ctx.to_device(c["g_point"], np.int32), ctx.to_device(c["wavenumber"]), ctx.to_device(c["d_wavenumber"])
    d_out = ctx.empty((c["ng"], w1.size))
    E.gpoint_fraction_dev(ctx, c["ng"], c["nwav"], d_g, d_wn, d_dwn, w1, w2, d_out)
    ref = O.gpoint_fraction(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], w1, w2)
    np.testing.assert_allclose(d_out.numpy(), ref, rtol=1e-12, atol=0)
    assert np.array_equal(d_out.numpy() == 0.0, ref == 0.0)

    ssi = S.solar_spectrum(c["wavenumber"])
    d_ssi, d_si = ctx.to_device(ssi), ctx.empty(w1.size)
    E.ssi_intervals_dev(ctx, c["nwav"], d_wn, d_ssi, w1, w2, d_si)
    np.testing.assert_allclose(d_si.numpy(), O.ssi_intervals(c["wavenumber"], ssi, w1, w2), rtol=1e-12)

    plan = E.GPointPlan(ctx, d_g, c["nwav"], c["ng"])
    t_lut = np.arange(120.0, 351.0)
    d_lut = ctx.empty((t_lut.size, c["ng"]))
    E.planck_lut_dev(ctx, plan, d_wn, d_dwn, t_lut, d_lut)
    ref = O.planck_lut(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], t_lut)
    np.testing.assert_allclose(d_lut.numpy(), ref, rtol=1e-12)


@pytest.mark.parametrize("nwav,ng,nt,holes", [(8192 * 3 + 77, 14, 231, False), (20011, 300, 40, False),
                                               (9000, 700, 33, False), (1, 1, 1, False), (70001, 40, 65, True)])
def test_planck_lut_with_the_temperatures_across_the_lanes(ctx, O, monkeypatch, nwav, ng, nt, holes):
    """planck_lut_kernel (temperatures on the lanes, create_look_up_table.cpp:565-587) against the oracle and against
    the reduction path it replaces: ragged last block, fewer warps per block as ng grows, nt not a multiple of 32."""
    rng = np.random.default_rng(nwav + ng)
    wn = np.sort(rng.uniform(0.5, 3260.0, nwav))
    dwn = rng.uniform(0.5e-3, 2e-3, nwav)
    g = rng.integers(0, ng, nwav).astype(np.int32)
    g[:ng] = np.arange(ng)[:nwav]  # every g point used
    if holes:  # wavenumbers of no g point, and a g point without wavenumbers (its sums are 0)
        g[rng.random(nwav) < 0.1] = -1
        g[g == 3] = 4
    t_lut = np.linspace(120.0, 350.0, nt)
    d_g, d_wn, d_dwn = ctx.to_device(g, np.int32), ctx.to_device(wn), ctx.to_device(dwn)
    plan = E.GPointPlan(ctx, d_g, nwav, ng)
    d_lut = ctx.empty((nt, ng))
    E.planck_lut_dev(ctx, plan, d_wn, d_dwn, t_lut, d_lut)
    got = d_lut.numpy().copy()
    np.testing.assert_allclose(got, O.planck_lut(ng, g, wn, dwn, t_lut), rtol=1e-12)
    monkeypatch.setenv("ECCKD_PLANCK_LUT_GENERAL", "1")
    E.planck_lut_dev(ctx, plan, d_wn, d_dwn, t_lut, d_lut)
    np.testing.assert_allclose(got, d_lut.numpy(), rtol=1e-13)
    monkeypatch.delenv("ECCKD_PLANCK_LUT_GENERAL")
    E.planck_lut_dev(ctx, plan, d_wn, d_dwn, t_lut, d_lut)
    assert np.array_equal(got, d_lut.numpy())  # the same bits on every call


def test_gpoint_fraction_bin_edges_exact(ctx, O):
    """Wavenumbers exactly on bin edges: bin i is (w1, w2] (create_look_up_table.cpp:544-545)."""
    wn = np.array([0.0, 10.0, 10.000000000000002, 20.0, 29.999999999999996, 30.0, 3259.9, 3260.0, 3260.1])
    dwn = np.arange(1.0, wn.size + 1)
    g = np.array([0, 0, 1, 1, 0, 1, 0, 1, 0], dtype=np.int32)
    w1, w2 = O.gpoint_fraction_grid(False, [0.0], [3260.0])
    d_out = ctx.empty((2, w1.size))
    E.gpoint_fraction_dev(ctx, 2, wn.size, ctx.to_device(g), ctx.to_device(wn), ctx.to_device(dwn), w1, w2, d_out)
    ref = O.gpoint_fraction(2, g, wn, dwn, w1, w2)
    np.testing.assert_allclose(d_out.numpy(), ref, rtol=1e-15, atol=0)


def test_compact_g_points(ctx, O):
    rng = np.random.default_rng(3)
    g = rng.choice([0, 2, 3, 7, 8], size=5000).astype(np.int32)
    d_g = ctx.to_device(g)
    n, gmap = E.compact_g_points_dev(ctx, 10, g.size, d_g)
    n_o, g_o, gmap_o = O.compact_g_points(10, g)
    assert n == n_o == 5 and np.array_equal(gmap, gmap_o) and np.array_equal(d_g.numpy(), g_o)
    g[17] = -1
    d_g = ctx.to_device(g)
    with pytest.raises(E.EcckdError) as ei:  # THROW(1), create_look_up_table.cpp:146-149
        E.compact_g_points_dev(ctx, 10, g.size, d_g)
    assert ei.value.code == 1
