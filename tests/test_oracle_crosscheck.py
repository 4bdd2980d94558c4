"""Pins for the CPU oracle (runs without a GPU).

The reference ships no tests or golden vectors for this path (SURVEY.md section
4), so the oracle is pinned by (i) analytic known-answer tests, (ii) agreement
of two independently written restatements (C++ oracle vs oracle/oracle_np.py),
(iii) the committed golden fixtures under tests/golden/ (test_golden.py).
"""
import numpy as np
import pytest

from oracle import oracle_np as N
from tests import synthetic as S
from tests.oracle_lib import METHODS


@pytest.mark.parametrize("method", METHODS)
def test_avg_cpp_vs_numpy(O, method):
    c = S.lw_spectrum_case(nlay=9, nwav=2500, ng=7, nband=2, seed=11)
    planck = N.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    a = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                           c["od"], planck, method)
    b = N.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                           c["od"], planck, method)
    S.assert_avg_close(a, b, method=method, scale=S.out_scale(c), rtol=2e-13)


def test_avg_known_answers(O):
    """Constant optical depth inside a g point => every method returns it."""
    c = S.lw_spectrum_case(nlay=6, nwav=1200, ng=5, nband=1, seed=2, zero_frac=0.0)
    tau_g = np.array([1e-6, 3e-3, 0.2, 1.5, 4.0])
    od = np.broadcast_to(tau_g[c["g_point"]], c["od"].shape) * np.linspace(0.5, 2.0, 6)[:, None]
    w = np.abs(np.sin(np.arange(1200.0)))[None, :] + 0.1 * np.arange(1, 7)[:, None]
    expect = np.linspace(0.5, 2.0, 6)[:, None] * tau_g[None, :]
    for m in METHODS:
        out, mn, mx, warn = O.average_optical_depth_to_g_point(5, -1.0, c["pressure_fl"], c["pressure_hl"],
                                                               c["g_point"], od, w, m)
        np.testing.assert_allclose(out, expect, rtol=1e-9, err_msg=m)
        # equal bounds are pushed apart by 0.99 / 1.01 (average_optical_depth.cpp:158-165)
        np.testing.assert_allclose(mn, 0.99 * expect, rtol=1e-15)
        np.testing.assert_allclose(mx, 1.01 * expect, rtol=1e-15)
        assert np.all(warn & 4)


def test_avg_empty_g_point_and_unassigned(O):
    c = S.lw_spectrum_case(nlay=4, nwav=500, ng=4, nband=1, seed=5)
    g = c["g_point"].copy()
    g[g == 2] = 1          # g point 2 empty
    g[::17] = -1           # unassigned wavenumbers belong to no g point
    planck = N.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    for m in ("transmission", "logarithmic"):
        a = O.average_optical_depth_to_g_point(4, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        b = N.average_optical_depth_to_g_point(4, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        S.assert_avg_close(a, b, method=m, scale=S.out_scale(c), rtol=2e-13)
        assert a[3][2] & 1 and np.all(a[0][:, 2] == 0.0) and np.all(a[1][:, 2] == 0.0)


def test_avg_bad_method(O):
    c = S.lw_spectrum_case(nlay=3, nwav=100, ng=2, nband=1, seed=5)
    with pytest.raises(RuntimeError) as ei:
        O.average_optical_depth_to_g_point(2, 1.0, c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"], c["od"], 8)
    assert ei.value.args[0] == 147  # PARAMETER_ERROR, average_optical_depth.cpp:130-133


def test_planck_cpp_vs_numpy_and_stefan_boltzmann(O):
    wn = np.linspace(1.0, 8000.0, 80001)
    dwn = np.full(wn.size, wn[1] - wn[0])
    t = np.array([180.0, 250.0, 320.0])
    a = O.planck_function(t, wn, dwn)
    b = N.planck_function(t, wn, dwn)
    # exp(x) amplifies the last-bit difference of x = (h/k)*(nu/T) by x (<= ~64 here)
    np.testing.assert_allclose(a, b, rtol=2e-13)
    # integral over the spectrum ~ sigma T^4 (constants of planck_function.cpp:29-33)
    np.testing.assert_allclose(a.sum(axis=1), 5.670400e-8 * t ** 4, rtol=2e-4)


def test_rt_lw_cpp_vs_numpy_and_limits(O):
    rng = np.random.default_rng(0)
    nlay, n = 14, 40
    od = np.exp(rng.normal(-3, 3, (nlay, n)))
    od[rng.uniform(size=od.shape) < 0.1] = 0.0            # exercises the emissivity <= 1e-5 branch
    planck = rng.uniform(1.0, 5.0, (nlay + 1, n))
    emis = rng.uniform(0.8, 1.0, n)
    sp = rng.uniform(1.0, 5.0, n)
    dn, up = O.radiative_transfer_lw(planck, od, emis, sp)
    dn2, up2 = N.radiative_transfer_lw(planck, od, emis, sp)
    # factor = 1 - emissivity/(D*tau) cancels for thin layers (relative error ~1e-16/emissivity),
    # and the C++ build contracts a*b+c into FMAs: compare at the flux tolerance scale
    np.testing.assert_allclose(dn, dn2, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(up, up2, rtol=1e-10, atol=1e-12)
    # opaque isothermal column: flux_dn -> planck, flux_up = planck everywhere
    pl = np.full((nlay + 1, n), 3.0)
    dn, up = O.radiative_transfer_lw(pl, np.full((nlay, n), 50.0), np.ones(n), np.full(n, 3.0))
    np.testing.assert_allclose(dn[1:], 3.0, rtol=1e-13)
    np.testing.assert_allclose(up, 3.0, rtol=1e-13)
    # transparent column: nothing emitted downward, surface emission passes through
    dn, up = O.radiative_transfer_lw(pl, np.zeros((nlay, n)), np.ones(n), np.full(n, 3.0))
    assert np.all(dn == 0.0) and np.all(up == 3.0)


def test_rt_lw_bb_matches_spectral_sum(O):
    """radiative_transfer_lw_bb == sum over wavenumber of the spectral solver when
    every emissivity is above the 1e-5 guard (the two differ only in the guard)."""
    rng = np.random.default_rng(1)
    nlay, n = 10, 30
    od = np.exp(rng.normal(-1, 1, (nlay, n))) + 1e-3
    grey = rng.uniform(0, 0.1, nlay)
    planck = rng.uniform(1.0, 5.0, (nlay + 1, n))
    emis, sp = np.ones(n), rng.uniform(1, 5, n)
    dn, up = O.radiative_transfer_lw(planck, od + grey[:, None], emis, sp)
    bdn, bup = O.radiative_transfer_lw_bb(planck, od, grey, emis, sp)
    np.testing.assert_allclose(bdn, dn.sum(axis=1), rtol=1e-13)
    np.testing.assert_allclose(bup, up.sum(axis=1), rtol=1e-13)


def test_rt_sw_closed_form(O):
    rng = np.random.default_rng(2)
    nlay, n = 12, 25
    od = np.exp(rng.normal(-2, 2, (nlay, n)))
    ssi = rng.uniform(0.1, 2.0, n)
    mu0 = 0.5
    dn = O.radiative_transfer_direct_sw(mu0, ssi, od)
    cum = np.concatenate([np.zeros((1, n)), np.cumsum(od, axis=0)])
    np.testing.assert_allclose(dn, mu0 * ssi * np.exp(-cum / mu0), rtol=1e-13)
    np.testing.assert_allclose(dn, N.radiative_transfer_direct_sw(mu0, ssi, od), rtol=1e-15)
    alb = rng.uniform(0, 0.5, n)
    dn2, up2 = O.radiative_transfer_norayleigh_sw(mu0, ssi, od, alb)
    dn3, up3 = N.radiative_transfer_norayleigh_sw(mu0, ssi, od, alb)
    np.testing.assert_allclose(dn2, dn3, rtol=1e-15)
    np.testing.assert_allclose(up2, up3, rtol=1e-15)
    np.testing.assert_allclose(up2[0], dn2[-1] * alb * np.exp(-2.0 * cum[-1]), rtol=1e-13)


def test_heating_rate(O):
    rng = np.random.default_rng(3)
    p = S.pressure_grid(8)
    dn, up = rng.uniform(0, 10, (9, 5)), rng.uniform(0, 10, (9, 5))
    np.testing.assert_allclose(O.heating_rate(p, dn, up), N.heating_rate(p, dn, up), rtol=1e-15)
    np.testing.assert_allclose(O.heating_rate(p, dn), N.heating_rate(p, dn), rtol=1e-15)


def test_ckd_optical_depth_and_planck_cpp_vs_numpy(O):
    c = S.ckd_model_case(ncol=4, nlay=11, ng=6, seed=9)
    t_fl = O.temperature_fl(c["temperature_hl"], c["pressure_hl"])
    for gas, vmr in zip(c["gases"], c["vmr_fl"]):
        a = O.ckd_calc_optical_depth(gas, c["log_pressure"], c["temperature"], c["pressure_hl"], t_fl, vmr)
        b = N.ckd_calc_optical_depth(gas["conc_dependence"], gas.get("vmr_lut"), gas.get("reference_vmr", 0.0),
                                     c["log_pressure"], c["temperature"], gas["molar_abs"], c["pressure_hl"], t_fl,
                                     vmr)
        # interpolation weights are differences such as pindex0 - ip0: a last-bit difference
        # of log() (NumPy's vs libm's) is amplified by index/weight, so 1e-13 is not attainable
        np.testing.assert_allclose(a, b, rtol=5e-12, atol=1e-300, err_msg=gas["name"])
    a = O.ckd_calc_planck_function(c["temperature_planck"], c["planck_lut"], c["temperature_hl"])
    b = N.ckd_calc_planck_function(c["temperature_planck"], c["planck_lut"], c["temperature_hl"])
    np.testing.assert_allclose(a, b, rtol=1e-14)
    assert (c["temperature_hl"] < 120).any() and (c["temperature_hl"] > 350).any()


def test_ckd_missing_concentration_is_an_error(O):
    c = S.ckd_model_case(ncol=2, nlay=5, ng=4, seed=1)
    t_fl = O.temperature_fl(c["temperature_hl"], c["pressure_hl"])
    with pytest.raises(RuntimeError):  # THROW(1), ckd_model.cpp:1048-1051
        O.ckd_calc_optical_depth(c["gases"][2], c["log_pressure"], c["temperature"], c["pressure_hl"], t_fl, None)


def test_rank_bands_cpp_vs_numpy_with_ties(O):
    rng = np.random.default_rng(5)
    n = 5000
    wn = np.linspace(0.5, 100.0, n)
    key = np.round(rng.normal(size=n), 1)       # many ties
    key[::50] = 0.0
    key[25::50] = -0.0                          # signed zeros compare equal
    for b1, b2 in [([0.0], [101.0]), ([0.5, 20.0, 60.0], [20.0, 60.0, 100.0]), ([10.0, 50.0], [50.0, 90.0])]:
        a = O.rank_bands(wn, key, b1, b2)
        b = N.rank_bands(wn, key, np.array(b1), np.array(b2))
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.array_equal(np.sort(a[1]), np.arange(n))  # rank is a permutation


def test_g_point_assignment_cpp_vs_numpy(O):
    rng = np.random.default_rng(6)
    n, ngas, nband = 4000, 3, 2
    ranks = np.stack([rng.permutation(n) for _ in range(ngas)]).astype(np.int32)
    n_g = np.array([[3, 2], [2, 2], [1, 3]], dtype=np.int32)
    gas_g, svs = [], []
    for a in range(ngas):
        ng_a = n_g[a].sum()
        cuts = np.sort(rng.choice(np.arange(1, n - 1), ng_a - 1, replace=False))
        r1 = np.concatenate([[0], cuts]).astype(np.int32)
        r2 = np.concatenate([cuts - 1, [n - 1]]).astype(np.int32)
        g = O.store_g_points(ranks[a], r1, r2)
        assert np.array_equal(g, N.store_g_points(ranks[a], r1, r2))
        gas_g.append(g)
        svs.append(np.sort(rng.uniform(0, 5, ng_a)))
    ng, gmin, gmax, band = O.overlap_g_points(n_g, svs)
    assert ng == int(nband * (1 - ngas) + n_g.sum())
    g, n_un = O.assign_g_points(np.stack(gas_g), gmin, gmax)
    assert np.array_equal(g, N.assign_g_points(np.stack(gas_g), gmin, gmax))
    assert n_un == int((g < 0).sum())


def test_gpoint_fraction_and_planck_lut(O):
    c = S.lw_spectrum_case(nlay=3, nwav=4000, ng=6, nband=2, seed=8)
    w1, w2 = O.gpoint_fraction_grid(False, [0.0, 1500.0], [1500.0, 3260.0])
    assert w1.size == 326 and w1[0] == 0.0 and w2[-1] == 3260.0
    a = O.gpoint_fraction(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], w1, w2)
    b = N.gpoint_fraction(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], w1, w2)
    np.testing.assert_allclose(a, b, rtol=1e-13)
    np.testing.assert_allclose(a.sum(axis=1), 1.0, rtol=1e-12)
    t = np.arange(120.0, 351.0, 23.0)
    np.testing.assert_allclose(O.planck_lut(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], t),
                               N.planck_lut(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], t), rtol=1e-13)


def _background_case(nx, ng, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((nx, nx))
    shape = a @ a.T / nx + np.eye(nx)  # symmetric positive definite, like an inverse covariance
    err = rng.uniform(0.5, 4.0, ng)
    dx = rng.standard_normal((nx, ng))
    grad0 = rng.standard_normal((nx, ng))
    return shape, err, dx, grad0


@pytest.mark.parametrize("nx,ng", [(1, 1), (7, 5), (318, 32)])
def test_background_cost_cpp_vs_numpy(O, nx, ng):
    """ckd_model.cpp:838-873: J_b = 1/2 sum_g dx_g^T (S / err_g^2) dx_g, gradient += (S dx_g) / err_g^2 per gas;
    the Rayleigh part adds to the cost only."""
    shape, err, dx, grad0 = _background_case(nx, ng, nx + ng)
    grad = grad0.copy()
    cost = O.background_cost_gas(shape, err, dx, grad)
    gl = (shape @ dx) / err[None, :] ** 2
    np.testing.assert_allclose(cost, 0.5 * np.sum(dx * gl), rtol=1e-13)
    np.testing.assert_allclose(grad - grad0, gl, rtol=1e-12, atol=1e-13 * np.abs(gl).max())
    # the gradient is the derivative of the cost: central differences along a random direction
    d = np.random.default_rng(1).standard_normal(dx.shape)
    h = 1e-6
    cp = O.background_cost_gas(shape, err, dx + h * d, np.zeros_like(dx))
    cm = O.background_cost_gas(shape, err, dx - h * d, np.zeros_like(dx))
    np.testing.assert_allclose((cp - cm) / (2 * h), np.sum(gl * d), rtol=1e-7)
    inv_b = 1.0 / err ** 2
    np.testing.assert_allclose(O.background_cost_rayleigh(inv_b, dx[0]), 0.5 * np.sum(inv_b * dx[0] ** 2), rtol=1e-14)
