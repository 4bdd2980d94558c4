"""The restated oracle held against the reference's OWN code (runs without a GPU).

oracle/_ref/libecckd_ref.so is built by oracle/build_ref_shim.sh from the reference's
translation units, unmodified and in place (average_optical_depth.cpp, planck_function.cpp,
radiative_transfer_{lw,sw}.cpp, single_gas_data.cpp, calc_cost_function_{lw,sw}.cpp,
heating_rate.h, rayleigh_scattering.h), against the Adept API shim in oracle/adept_shim/
(Adept itself cannot be installed here).  oracle/ref_glue.cpp only wraps raw arrays and
calls the reference's functions; it exports the oracle's entry-point names, so every test
below runs the same Python call on both libraries.

Both sum sequentially and use libm, so agreement is expected to a few ulp; the tolerance
1e-13 leaves room for FMA contraction differences under -O2 -march=native.
"""
import numpy as np
import pytest

from tests import synthetic as S
from tests.oracle_lib import METHODS, reference_oracle

R = reference_oracle()
pytestmark = pytest.mark.skipif(R is None, reason="oracle/_ref/libecckd_ref.so not built (needs /root/reference)")

RTOL = 1e-13


def close(a, b, rtol=RTOL, atol=0.0, what=""):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, err_msg=what)


def test_reference_library_is_the_reference():
    R.L.ref_build_info.restype = __import__("ctypes").c_char_p
    info = R.L.ref_build_info().decode()
    for unit in ("average_optical_depth.cpp", "planck_function.cpp", "radiative_transfer_lw.cpp",
                 "radiative_transfer_sw.cpp", "single_gas_data.cpp", "calc_cost_function_lw.cpp",
                 "calc_cost_function_sw.cpp"):
        assert unit in info


def test_planck_function(O):
    c = S.lw_spectrum_case(nlay=7, nwav=900, ng=4, nband=1, seed=3)
    a = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    b = R.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    close(a, b, what="planck_function.cpp:22-54")


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("vmr", [4.0e-4, -1.0])
def test_average_optical_depth(O, method, vmr):
    c = S.lw_spectrum_case(nlay=9, nwav=2500, ng=7, nband=2, seed=11)
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    a = O.average_optical_depth_to_g_point(c["ng"], vmr, c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"],
                                           planck, method)
    R.L.ref_warning_count(1)
    b = R.average_optical_depth_to_g_point(c["ng"], vmr, c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"],
                                           planck, method)
    scale = S.out_scale(c) if vmr > 0 else None
    S.assert_avg_close(a[:3], b[:3], method=method, scale=scale, rtol=2e-13)
    # the reference prints one warning per flagged condition per g point (:136,:154,:159)
    nwarn = sum(bin(int(w)).count("1") for w in a[3])
    assert R.L.ref_warning_count(1) == nwarn


@pytest.mark.parametrize("method", ["transmission", "logarithmic", "hybrid-logarithmic-transmission-3"])
def test_average_optical_depth_ssi_weights_zeros_and_constant_g(O, method):
    """Shortwave-style call: one weight row for all layers; a g point of constant optical depth
    (equal-bounds fix x0.99 / x1.01) and one with zeros only."""
    c = S.sw_spectrum_case(nlay=6, nwav=1500, ng=6, nband=2, seed=4)
    od = c["od"].copy()
    od[:, c["g_point"] == 1] = 0.0
    od[:, c["g_point"] == 2] = np.linspace(1e-3, 2.0, 6)[:, None]
    a = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                           c["ssi"], method)
    R.L.ref_warning_count(1)
    b = R.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                           c["ssi"], method)
    S.assert_avg_close(a[:3], b[:3], method=method, scale=S.out_scale(c), rtol=2e-13)
    assert np.all(a[3][2] & 4) and R.L.ref_warning_count(1) >= 1


def test_average_optical_depth_unassigned_wavenumbers(O):
    c = S.lw_spectrum_case(nlay=4, nwav=500, ng=4, nband=1, seed=5)
    g = c["g_point"].copy()
    g[::17] = -1
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    for m in ("linear", "square-root", "transmission-10"):
        a = O.average_optical_depth_to_g_point(4, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        b = R.average_optical_depth_to_g_point(4, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        S.assert_avg_close(a[:3], b[:3], method=m, scale=S.out_scale(c), rtol=2e-13)


def test_average_optical_depth_bad_method_and_empty_g_point(O):
    c = S.lw_spectrum_case(nlay=3, nwav=100, ng=2, nband=1, seed=5)
    with pytest.raises(RuntimeError) as ei:
        R.average_optical_depth_to_g_point(2, 1.0, c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"], c["od"], 8)
    assert ei.value.args[0] == 147  # PARAMETER_ERROR, average_optical_depth.cpp:130-133
    # An empty g point: the reference assigns its never-sized min/max vectors into the output
    # (average_optical_depth.cpp:138-140 then :176-181) which Adept rejects with a size mismatch;
    # without min/max outputs it returns zeros.  The restatement (and the CUDA path) return zeros
    # and flag the g point instead of failing.
    g = np.zeros(100, dtype=np.int32)
    with pytest.raises(RuntimeError) as ei:
        R.average_optical_depth_to_g_point(2, 1.0, c["pressure_fl"], c["pressure_hl"], g, c["od"], c["od"], 0)
    assert ei.value.args[0] == -10
    a = O.average_optical_depth_to_g_point(2, 1.0, c["pressure_fl"], c["pressure_hl"], g, c["od"], c["od"], 0)
    assert a[3][1] & 1 and np.all(a[0][:, 1] == 0.0)


def rt_inputs(nlay=11, n=70, seed=0):
    rng = np.random.default_rng(seed)
    od = np.exp(rng.normal(-2.0, 3.0, (nlay, n)))
    od[rng.uniform(size=od.shape) < 0.05] = 0.0
    od[0, :5] = 1e-9  # below the 1e-5 emissivity threshold
    planck = np.cumsum(rng.uniform(0.1, 1.0, (nlay + 1, n)), axis=0)
    emis = rng.uniform(0.9, 1.0, n)
    surf = planck[-1] * rng.uniform(1.0, 1.1, n)
    return od, planck, emis, surf


def test_radiative_transfer_lw(O):
    od, planck, emis, surf = rt_inputs()
    a = O.radiative_transfer_lw(planck, od, emis, surf)
    b = R.radiative_transfer_lw(planck, od, emis, surf)
    close(a[0], b[0], what="flux_dn"), close(a[1], b[1], what="flux_up")
    # heating_rate.h:25-49
    p = np.linspace(1.0, 1.0e5, od.shape[0] + 1)
    # same flux inputs to both, so that the flux differences cancel identically
    close(O.heating_rate(p, a[0], a[1]), R.heating_rate(p, a[0], a[1]))
    close(O.heating_rate(p, a[0], None), R.heating_rate(p, a[0], None))
    close(O.heating_rate_single(p, a[0][:, 0], a[1][:, 0]), R.heating_rate_single(p, a[0][:, 0], a[1][:, 0]))
    close(O.heating_rate_single(p, a[0][:, 0], None), R.heating_rate_single(p, a[0][:, 0], None))


def test_radiative_transfer_lw_bb(O):
    od, planck, emis, surf = rt_inputs(seed=1)
    grey = np.exp(np.random.default_rng(2).normal(-3, 2, od.shape[0]))
    grey[3] = 0.0
    a = O.radiative_transfer_lw_bb(planck, od, grey, emis, surf)
    b = R.radiative_transfer_lw_bb(planck, od, grey, emis, surf)
    close(a[0], b[0]), close(a[1], b[1])


def test_radiative_transfer_sw(O):
    od, _, _, _ = rt_inputs(seed=3)
    n = od.shape[1]
    rng = np.random.default_rng(4)
    ssi = rng.uniform(0.1, 3.0, n)
    alb = rng.uniform(0.0, 0.6, n)
    close(O.radiative_transfer_direct_sw(0.37, ssi, od), R.radiative_transfer_direct_sw(0.37, ssi, od))
    a = O.radiative_transfer_norayleigh_sw(0.5, ssi, od, alb)
    b = R.radiative_transfer_norayleigh_sw(0.5, ssi, od, alb)
    close(a[0], b[0]), close(a[1], b[1])
    grey = rng.uniform(0.0, 0.2, od.shape[0])
    for direct in (False, True):
        a = O.radiative_transfer_sw_bb(0.5, ssi, od, grey, 0.15, direct_only=direct)
        b = R.radiative_transfer_sw_bb(0.5, ssi, od, grey, 0.15, direct_only=direct)
        close(a[0], b[0])
        if not direct:
            close(a[1], b[1])


def test_rayleigh_molar_scattering_coeff(O):
    wn = np.concatenate([np.linspace(275.0, 49975.0, 995), [19999.9, 20000.0, 20000.1]])  # both branches, the 0.5 um seam
    close(O.rayleigh_molar_scattering_coeff(wn), R.rayleigh_molar_scattering_coeff(wn))


def test_store_and_overlap_g_points_bit_exact(O):
    rng = np.random.default_rng(8)
    nwav = 4000
    rank = rng.permutation(nwav).astype(np.int32)
    cuts = np.sort(rng.choice(np.arange(1, nwav - 1), 6, replace=False))
    rank1 = np.concatenate([[0], cuts]).astype(np.int32)
    rank2 = np.concatenate([cuts - 1, [nwav - 1]]).astype(np.int32)
    rank2[2] -= 40  # leaves a gap of unassigned ranks -> -1
    assert np.array_equal(O.store_g_points(rank, rank1, rank2), R.store_g_points(rank, rank1, rank2))
    for seed in range(6):
        r = np.random.default_rng(seed)
        ngas, nband = int(r.integers(1, 6)), int(r.integers(1, 5))
        n_g = r.integers(1, 6, (ngas, nband)).astype(np.int32)
        sv = [np.sort(r.uniform(0, 1, int(n_g[i].sum()))) if seed % 2 else r.uniform(0, 1, int(n_g[i].sum()))
              for i in range(ngas)]
        a = O.overlap_g_points(n_g, sv)
        b = R.overlap_g_points(n_g, sv)
        assert a[0] == b[0]
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)


def interval_case(nlay=10, n=160, seed=0):
    rng = np.random.default_rng(seed)
    od, planck, emis, surf = rt_inputs(nlay, n, seed)
    p = np.sort(np.concatenate([[1.0], rng.uniform(10.0, 9.0e4, nlay - 1), [1.0e5]]))
    lw = np.diff(np.sqrt(p))
    lw /= lw.sum()
    fit = np.exp(rng.normal(-2, 1, nlay))
    return dict(p=p, od=od, planck=planck, emis=emis, surf=surf, lw=lw, fit=fit, rng=rng)


def test_calc_cost_function_lw(O):
    """Interval error of find_g_points (calc_cost_function_lw.cpp:23-110)."""
    c = interval_case()
    true_dn, true_up = O.radiative_transfer_lw(c["planck"], c["od"] * 1.3, c["emis"], c["surf"])
    hr = O.heating_rate(c["p"], true_dn, true_up)
    args = (c["p"], c["planck"], c["emis"], c["surf"], c["od"], c["fit"], true_dn[-1], true_up[0], hr, 0.2, c["lw"])
    a, b = O.calc_cost_function_lw(*args), R.calc_cost_function_lw(*args)
    assert a > 0 and abs(a - b) <= 1e-12 * abs(b)


@pytest.mark.parametrize("albedo", [0.0, 0.15])
def test_calc_cost_function_sw(O, albedo):
    c = interval_case(seed=5)
    ssi = c["rng"].uniform(0.1, 2.0, c["od"].shape[1])
    true_dn, true_up = O.radiative_transfer_norayleigh_sw(0.5, ssi, c["od"] * 1.2, np.full(ssi.size, 0.15))
    hr = O.heating_rate(c["p"], true_dn, None)
    args = (0.5, c["p"], ssi, albedo, c["od"], c["fit"] * 0.1, true_dn[-1], true_up[0], hr, 0.2, c["lw"])
    a, b = O.calc_cost_function_sw(*args), R.calc_cost_function_sw(*args)
    assert a > 0 and abs(a - b) <= 1e-12 * abs(b)


def ckd_cost_case(nlay=12, ng=16, nband=5, seed=0):
    rng = np.random.default_rng(seed)
    od, planck, _, surf = rt_inputs(nlay, ng, seed)
    band_mapping = np.sort(rng.integers(0, nband, ng)).astype(np.int32)
    band_mapping[:nband] = np.arange(nband)
    band_mapping.sort()
    p = np.sort(np.concatenate([[1.0], rng.uniform(10.0, 9.0e4, nlay - 1), [1.0e5]]))
    lw = np.diff(p ** 0.5)
    lw /= lw.sum()
    return dict(p=p, od=od, planck=planck, surf=surf, bm=band_mapping, lw=lw, rng=rng, nband=nband)


@pytest.mark.parametrize("banded", [True, False])
@pytest.mark.parametrize("relative", [False, True])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8, 0.1), (0.5, 0.0, 0.0, 0.0)])
def test_calc_cost_function_ckd_lw(O, banded, relative, weights):
    """Forward cost of optimize_lut (calc_cost_function_lw.cpp:115-232)."""
    c = ckd_cost_case(seed=2)
    ng = c["od"].shape[1]
    nband = c["nband"] if banded else ng
    bm = c["bm"] if banded else None
    rng = c["rng"]
    emis = rng.uniform(0.9, 1.0, nband)
    emis_g = emis[bm] if banded else emis
    t_dn, t_up = O.radiative_transfer_lw(c["planck"], c["od"] * 1.1, emis_g, c["surf"])
    if banded:
        t_dn = np.stack([t_dn[:, bm == b].sum(1) for b in range(nband)], 1)
        t_up = np.stack([t_up[:, bm == b].sum(1) for b in range(nband)], 1)
    hr = O.heating_rate(c["p"], t_dn, t_up)
    rel = (rng.uniform(0, 1, (c["od"].shape[0] + 1, ng)), rng.uniform(0, 1, (c["od"].shape[0] + 1, ng))) if relative \
        else (None, None)
    sfd, sfu = rng.uniform(0, 5, ng), rng.uniform(0, 5, ng)
    kw = dict(spectral_flux_dn_surf=sfd, spectral_flux_up_toa=sfu, rel_dn=rel[0], rel_up=rel[1], band_mapping=bm)
    args = (c["p"], c["planck"], emis, c["surf"], c["od"], t_dn, t_up, hr) + weights + (c["lw"],)
    a, b = O.calc_cost_function_ckd_lw(*args, **kw), R.calc_cost_function_ckd_lw(*args, **kw)
    assert a > 0 and abs(a - b) <= 1e-12 * abs(b), (a, b)


@pytest.mark.parametrize("albedo_on", [True, False])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8), (0.5, 0.0, 0.0)])
def test_calc_cost_function_ckd_sw(O, albedo_on, weights):
    """calc_cost_function_sw.cpp:115-277: x20 TOA term, no upwelling in the heating rate,
    broadband upwelling terms only when every albedo > 0, per-band cost accumulation."""
    c = ckd_cost_case(seed=3)
    ng, nband, bm, rng = c["od"].shape[1], c["nband"], c["bm"], c["rng"]
    ssi = rng.uniform(1.0, 100.0, ng)
    albedo = rng.uniform(0.05, 0.5, nband) if albedo_on else np.zeros(nband)
    t_dn, t_up = O.radiative_transfer_norayleigh_sw(0.6, ssi, c["od"] * 1.1, albedo[bm])
    t_dn = np.stack([t_dn[:, bm == b].sum(1) for b in range(nband)], 1)
    t_up = np.stack([t_up[:, bm == b].sum(1) for b in range(nband)], 1)
    hr = O.heating_rate(c["p"], t_dn, None)
    pa, pb = np.full(nband, 0.25), np.full(nband, 0.25)
    kw = dict(spectral_boundary_weights=rng.uniform(0, 1, ng), spectral_flux_dn_surf=rng.uniform(0, 5, ng),
              spectral_flux_up_toa=rng.uniform(0, 5, ng), band_mapping=bm)
    args = (0.6, c["p"], ssi, albedo, c["od"], t_dn, t_up, hr) + weights + (c["lw"],)
    a = O.calc_cost_function_ckd_sw(*args, cost_fn_per_band=pa, **kw)
    b = R.calc_cost_function_ckd_sw(*args, cost_fn_per_band=pb, **kw)
    assert a > 0 and abs(a - b) <= 1e-12 * abs(b), (a, b)
    close(pa, pb, rtol=1e-12)
