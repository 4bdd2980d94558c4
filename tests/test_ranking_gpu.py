"""GPU parity of g-point ranking and assignment (reorder_spectrum.cpp:111-300,
single_gas_data.h:56-62, find_g_points.cpp:1455-1481).  Integer outputs are
bit-exact given identical FP64 keys."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu


def _rank(ctx, wn, key, b1, b2):
    n = wn.size
    d_wn, d_key = ctx.to_device(wn), ctx.to_device(key)
    d_ib, d_rank, d_ord = ctx.empty(n, np.int32), ctx.empty(n, np.int32), ctx.empty(n, np.int32)
    E.rank_bands_dev(ctx, n, d_wn, d_key, b1, b2, d_ib, d_rank, d_ord)
    return d_ib.numpy(), d_rank.numpy(), d_ord.numpy()


@pytest.mark.parametrize("n", [1, 2, 255, 2048, 2049, 70001])
def test_rank_single_band_bit_exact(ctx, O, n):
    rng = np.random.default_rng(n)
    wn = np.linspace(0.5, 100.0, n)
    key = rng.normal(size=n) * 10.0 ** rng.integers(-8, 8, n)  # all exponent bytes vary
    got = _rank(ctx, wn, key, [0.0], [101.0])
    ref = O.rank_bands(wn, key, [0.0], [101.0])
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


def test_rank_ties_signed_zeros_and_bands(ctx, O):
    rng = np.random.default_rng(5)
    n = 50000
    wn = np.linspace(0.5, 100.0, n)
    key = np.round(rng.normal(size=n), 1)
    key[::50] = 0.0
    key[25::50] = -0.0
    key[7::1000] = -np.inf
    key[9::1000] = np.inf
    key[11::1000] = 5e-324      # denormals
    key[13::1000] = -5e-324
    for b1, b2 in [([0.0], [101.0]),
                   ([0.5, 20.0, 60.0], [20.0, 60.0, 100.0]),     # last band closed above
                   ([10.0, 50.0], [50.0, 90.0]),                 # points outside every band keep their place
                   (list(np.linspace(0.5, 97.5, 44)), list(np.linspace(2.76, 100.0, 44)))]:
        got = _rank(ctx, wn, key, b1, b2)
        ref = O.rank_bands(wn, key, b1, b2)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        assert np.array_equal(np.sort(got[1]), np.arange(n))


def test_rank_overlapping_bands_follow_the_reference_loop(ctx, O):
    """Bands whose index ranges overlap are sorted one after the other on the running order
    (reorder_spectrum.cpp:278-294); disjoint bands take the single-sort path."""
    rng = np.random.default_rng(8)
    n = 30011
    wn = np.linspace(0.5, 100.0, n)
    key = np.round(rng.normal(size=n), 2)
    for b1, b2 in [([0.5, 30.0], [60.0, 100.0]), ([0.5, 10.0, 5.0], [20.0, 40.0, 100.0])]:
        got = _rank(ctx, wn, key, b1, b2)
        ref = O.rank_bands(wn, key, b1, b2)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)


def test_rank_many_tiles_smooth_keys_gaps_and_unordered_bands(ctx, O):
    """2^20+3 wavenumbers (257 tiles, ragged last tile), smooth keys (whole warps on one digit), bands
    listed out of positional order with gaps between them and at both ends."""
    n = (1 << 20) + 3
    wn = np.linspace(1.0, 3000.0, n)
    rng = np.random.default_rng(21)
    key = 6.0 + 5.0 * np.sin(wn * 0.05) + 1e-3 * np.round(rng.normal(size=n), 1)
    key[::9973] = key[1::9973]          # exact ties far apart
    b1 = [1500.0, 10.0, 2200.0, 700.0]
    b2 = [2000.0, 650.0, 2990.0, 1400.0]
    got = _rank(ctx, wn, key, b1, b2)
    ref = O.rank_bands(wn, key, b1, b2)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # size-independent properties: a permutation, the identity outside the bands, sorted inside
    ib, rank, order = got
    assert np.array_equal(np.sort(rank), np.arange(n))
    assert np.array_equal(order[ib < 0], np.nonzero(ib < 0)[0])
    for j in range(4):
        m = np.nonzero(ib == j)[0]
        k = key[order[m[0]:m[-1] + 1]]
        assert np.all(np.diff(k) >= 0)


def test_rank_constant_key_is_identity(ctx):
    n = 10000
    wn = np.linspace(1.0, 2.0, n)
    got = _rank(ctx, wn, np.full(n, 3.25), [0.0], [3.0])
    assert np.array_equal(got[1], np.arange(n))


def test_rank_empty_band_is_parameter_error(ctx):
    wn = np.linspace(1.0, 2.0, 100)
    with pytest.raises(E.EcckdError) as ei:
        _rank(ctx, wn, wn.copy(), [0.0, 5.0], [5.0, 6.0])
    assert ei.value.code == 147


def test_sort_keys(ctx, O):
    c = S.lw_spectrum_case(nlay=54, nwav=20000, ng=8, nband=1, seed=17, zero_frac=0.0, vmr=1.0)
    od = c["od"] * 3e4
    od[:, ::7] *= 1e-6     # thin columns take the column-optical-depth branch
    d_od, d_wn, d_dwn = ctx.to_device(od), ctx.to_device(c["wavenumber"]), ctx.to_device(c["d_wavenumber"])
    d_key, d_col = ctx.empty(c["nwav"]), ctx.empty(c["nwav"])
    E.sort_key_lw_dev(ctx, 54, c["nwav"], c["pressure_hl"], d_wn, d_dwn, d_od, c["nwav"], 0.5, d_key, d_col)
    key_o, col_o = O.sort_key_lw(c["pressure_hl"], c["wavenumber"], c["d_wavenumber"], od, 0.5)
    np.testing.assert_allclose(d_col.numpy(), col_o, rtol=1e-15)
    ok = np.isfinite(key_o)
    assert np.array_equal(np.isfinite(d_key.numpy()), ok)
    # key = ratio of sums of heating rates that are differences of fluxes: conditioned comparison
    np.testing.assert_allclose(d_key.numpy()[ok], key_o[ok], rtol=1e-9, atol=1e-12)
    assert (col_o < 0.5).any() and (col_o > 0.5).any()

    E.sort_key_sw_dev(ctx, 54, c["nwav"], c["pressure_hl"], d_od, c["nwav"], 0.25, d_key, d_col)
    key_s, col_s, rc = O.sort_key_sw(c["pressure_hl"], od, 0.25)
    assert rc == 0
    np.testing.assert_allclose(d_key.numpy(), key_s, rtol=1e-12, atol=1e-15)


def test_store_overlap_assign_bit_exact(ctx, O):
    rng = np.random.default_rng(6)
    n, ngas, nband = 60000, 5, 3
    n_g = rng.integers(1, 5, size=(ngas, nband)).astype(np.int32)
    gas_g, svs, d_list = [], [], []
    for a in range(ngas):
        rank = rng.permutation(n).astype(np.int32)
        ng_a = int(n_g[a].sum())
        cuts = np.sort(rng.choice(np.arange(1, n - 1), ng_a - 1, replace=False))
        r1 = np.concatenate([[0], cuts]).astype(np.int32)
        r2 = np.concatenate([cuts - 1, [n - 1]]).astype(np.int32)
        if a == 1:
            r2[0] -= 5      # a gap: some ranks fall in no g point
            r1[-1] -= 3     # an overlap: later g point wins
        d_rank, d_g = ctx.to_device(rank), ctx.empty(n, np.int32)
        E.store_g_points_dev(ctx, n, d_rank, r1, r2, d_g)
        ref = O.store_g_points(rank, r1, r2)
        assert np.array_equal(d_g.numpy(), ref)
        gas_g.append(ref)
        svs.append(np.sort(rng.uniform(0, 5, ng_a)))
    ng, gmin, gmax, band = E.overlap_g_points(n_g, svs)
    ng_o, gmin_o, gmax_o, band_o = O.overlap_g_points(n_g, svs)
    assert ng == ng_o and np.array_equal(gmin, gmin_o) and np.array_equal(gmax, gmax_o)
    gg = np.stack(gas_g)
    d_gg, d_out = ctx.to_device(gg), ctx.empty(n, np.int32)
    n_un = E.assign_g_points_dev(ctx, n, d_gg, gmin, gmax, d_out)
    ref, n_un_o = O.assign_g_points(gg, gmin, gmax)
    assert np.array_equal(d_out.numpy(), ref) and n_un == n_un_o
