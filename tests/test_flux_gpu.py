"""GPU parity of the CKD flux solver (ckd_model.cpp:923-1145, radiative_transfer_*.cpp,
heating_rate.h, run_ckd.cpp:119-367) against the CPU oracle.  Flux tolerance: 1e-6 W m-2
(north_star); optical depths and Planck functions relative 1e-12."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu
FLUX_ATOL = S.FLUX_ATOL


def _model(ctx, c, sw=False):
    if sw:
        return E.CkdModel(ctx, c["gases"], c["log_pressure"], c["temperature"],
                          solar_irradiance=c["solar_irradiance"], rayleigh_molar_scat=c["rayleigh_molar_scat"])
    return E.CkdModel(ctx, c["gases"], c["log_pressure"], c["temperature"], c["temperature_planck"], c["planck_lut"])


def test_rt_lw_spectral(ctx, O):
    rng = np.random.default_rng(0)
    nlay, n = 54, 3000
    od = np.exp(rng.normal(-3, 3, (nlay, n)))
    od[rng.uniform(size=od.shape) < 0.1] = 0.0
    planck = rng.uniform(0.0, 30.0, (nlay + 1, n))
    emis, sp = rng.uniform(0.8, 1.0, n), rng.uniform(0.0, 30.0, n)
    dn, up = E.radiative_transfer_lw(ctx, planck, od, emis, sp)
    dn_o, up_o = O.radiative_transfer_lw(planck, od, emis, sp)
    assert np.max(np.abs(dn - dn_o)) <= FLUX_ATOL * 1e-3 and np.max(np.abs(up - up_o)) <= FLUX_ATOL * 1e-3
    # limits: opaque isothermal and transparent columns
    pl = np.full((nlay + 1, n), 3.0)
    dn, up = E.radiative_transfer_lw(ctx, pl, np.full((nlay, n), 50.0), np.ones(n), np.full(n, 3.0))
    np.testing.assert_allclose(dn[1:], 3.0, rtol=1e-13)
    dn, up = E.radiative_transfer_lw(ctx, pl, np.zeros((nlay, n)), np.ones(n), np.full(n, 3.0))
    assert np.all(dn == 0.0) and np.all(up == 3.0)


def test_rt_sw(ctx, O):
    rng = np.random.default_rng(2)
    nlay, n = 54, 2000
    od = np.exp(rng.normal(-2, 2, (nlay, n)))
    ssi, alb = rng.uniform(0.1, 2.0, n), rng.uniform(0, 0.5, n)
    for mu0 in (0.5, 0.1, 1.0):
        dn = E.radiative_transfer_direct_sw(ctx, mu0, ssi, od)
        assert np.max(np.abs(dn - O.radiative_transfer_direct_sw(mu0, ssi, od))) <= FLUX_ATOL * 1e-3
        dn2, up2 = E.radiative_transfer_norayleigh_sw(ctx, mu0, ssi, od, alb)
        dn_o, up_o = O.radiative_transfer_norayleigh_sw(mu0, ssi, od, alb)
        assert np.max(np.abs(dn2 - dn_o)) <= FLUX_ATOL * 1e-3 and np.max(np.abs(up2 - up_o)) <= FLUX_ATOL * 1e-3
    cum = np.concatenate([np.zeros((1, n)), np.cumsum(od, axis=0)])
    np.testing.assert_allclose(E.radiative_transfer_direct_sw(ctx, 0.5, ssi, od), 0.5 * ssi * np.exp(-cum / 0.5),
                               rtol=1e-12, atol=1e-300)


def test_heating_rate(ctx, O):
    rng = np.random.default_rng(3)
    p = S.pressure_grid(54)
    dn, up = rng.uniform(0, 400, (55, 777)), rng.uniform(0, 400, (55, 777))
    np.testing.assert_allclose(E.heating_rate(ctx, p, dn, up), O.heating_rate(p, dn, up), rtol=1e-14, atol=1e-18)
    np.testing.assert_allclose(E.heating_rate(ctx, p, dn), O.heating_rate(p, dn), rtol=1e-14, atol=1e-18)


def test_ckd_optical_depth_each_gas(ctx, O):
    c = S.ckd_model_case(ncol=7, nlay=54, ng=32, seed=9)
    m = _model(ctx, c)
    t_fl = O.temperature_fl(c["temperature_hl"], c["pressure_hl"])
    d_p, d_t = ctx.to_device(c["pressure_hl"]), ctx.to_device(t_fl)
    d_od = ctx.empty((7, 54, 32))
    total = np.zeros((7, 54, 32))
    d_tot = ctx.zeros((7, 54, 32))
    for i, (gas, vmr) in enumerate(zip(c["gases"], c["vmr_fl"])):
        d_v = None if vmr is None else ctx.to_device(vmr)
        m.calc_optical_depth_dev(i, 7, 54, d_p, d_t, d_v, d_od)
        ref = O.ckd_calc_optical_depth(gas, c["log_pressure"], c["temperature"], c["pressure_hl"], t_fl, vmr)
        # index weights are differences (pindex0 - ip0): last-bit differences of log() are amplified
        np.testing.assert_allclose(d_od.numpy(), ref, rtol=5e-12, atol=1e-300, err_msg=gas["name"])
        m.calc_optical_depth_dev(i, 7, 54, d_p, d_t, d_v, d_tot, accumulate=True)
        total += ref
    np.testing.assert_allclose(d_tot.numpy(), total, rtol=1e-11, atol=1e-18)
    with pytest.raises(E.EcckdError) as ei:  # concentration not provided: THROW(1)
        m.calc_optical_depth_dev(2, 7, 54, d_p, d_t, None, d_od)
    assert ei.value.code == 1


def test_ckd_planck_function(ctx, O):
    c = S.ckd_model_case(ncol=9, nlay=54, ng=64, seed=3)
    m = _model(ctx, c)
    d_t = ctx.to_device(c["temperature_hl"])
    d_pl = ctx.empty((9, 55, 64))
    m.calc_planck_function_dev(9 * 55, d_t, d_pl)
    ref = O.ckd_calc_planck_function(c["temperature_planck"], c["planck_lut"], c["temperature_hl"])
    np.testing.assert_allclose(d_pl.numpy(), ref, rtol=1e-14)


@pytest.mark.parametrize("ng", [32, 64, 20])
def test_run_ckd_lw_config_a(ctx, O, ng):
    """BASELINE config A: 50 columns x 54 layers x 32 g points, 8 gases."""
    c = S.ckd_model_case(ncol=50, nlay=54, ng=ng, seed=1234)
    m = _model(ctx, c)
    got = m.run_ckd(c["vmr_fl"], c["pressure_hl"], c["temperature_hl"])
    ref = O.run_ckd(c, is_sw=False)
    np.testing.assert_allclose(got["optical_depth"], ref["optical_depth"], rtol=2e-11, atol=1e-18)
    np.testing.assert_allclose(got["planck_hl"], ref["planck_hl"], rtol=1e-14)
    np.testing.assert_allclose(got["planck_surf"], ref["planck_surf"], rtol=1e-14)
    for k in ("flux_dn", "flux_up", "bb_flux_dn", "bb_flux_up"):
        err = np.max(np.abs(got[k] - ref[k]))
        assert err <= FLUX_ATOL, (k, err)
    assert np.all(got["optical_depth"] >= 0.0)
    assert ref["bb_flux_up"].max() > 100.0  # fluxes of physical magnitude: the 1e-6 W m-2 tolerance is meaningful


def test_run_ckd_sw(ctx, O):
    c = S.ckd_model_case(ncol=50, nlay=54, ng=32, seed=77, sw=True)
    m = _model(ctx, c, sw=True)
    got = m.run_ckd(c["vmr_fl"], c["pressure_hl"], c["temperature_hl"], tsi=1361.0)
    ref = O.run_ckd(c, is_sw=True, tsi=1361.0)
    np.testing.assert_allclose(got["optical_depth"], ref["optical_depth"], rtol=2e-11, atol=1e-18)
    for k in ("flux_dn", "bb_flux_dn"):
        err = np.max(np.abs(got[k] - ref[k]))
        assert err <= FLUX_ATOL, (k, err)
    np.testing.assert_allclose(got["bb_flux_dn"][:, 0], 0.5 * 1361.0, rtol=1e-13)


def test_run_ckd_negative_od_is_clamped(ctx, O):
    """relative-linear gases below their reference concentration give negative optical
    depth; the sum is clamped at zero (run_ckd.cpp:317)."""
    c = S.ckd_model_case(ncol=6, nlay=20, ng=8, seed=5)
    for i in (0, 1, 2, 3, 6, 7):
        c["gases"][i]["molar_abs"] = c["gases"][i]["molar_abs"] * 1e-9
    c["vmr_fl"][4] = c["vmr_fl"][4] * 0.01
    m = _model(ctx, c)
    got = m.run_ckd(c["vmr_fl"], c["pressure_hl"], c["temperature_hl"])
    ref = O.run_ckd(c, is_sw=False)
    assert (ref["optical_depth"] == 0.0).any()
    assert np.array_equal(got["optical_depth"] == 0.0, ref["optical_depth"] == 0.0)
    assert np.max(np.abs(got["flux_up"] - ref["flux_up"])) <= FLUX_ATOL


@pytest.mark.parametrize("sw", [False, True])
def test_calc_total_optical_depth(ctx, O, sw):
    """calc_total_optical_depth (solve_adept.cpp:23-69): Rayleigh + gases; a gas missing from the training
    scene is included only when it has no concentration dependence."""
    c = S.ckd_model_case(ncol=6, nlay=54, ng=32, seed=17, sw=sw)
    m = _model(ctx, c, sw=sw)
    t_fl = O.temperature_fl(c["temperature_hl"], c["pressure_hl"])
    vmr = list(c["vmr_fl"])
    absent = [i for i, g in enumerate(c["gases"]) if g["conc_dependence"] != E.CONC_NONE][-1]
    vmr[absent] = None  # "not present in LBL file": skipped
    ref = np.zeros((6, 54, 32))
    if sw:
        dp = np.diff(c["pressure_hl"], axis=1)
        ref += (dp * (1.0 / (9.80665 * 0.001 * 28.970)))[:, :, None] * c["rayleigh_molar_scat"][None, None, :]
    for i, (gas, v) in enumerate(zip(c["gases"], vmr)):
        if v is None and gas["conc_dependence"] != E.CONC_NONE:
            continue
        ref += O.ckd_calc_optical_depth(gas, c["log_pressure"], c["temperature"], c["pressure_hl"], t_fl, v)
    d_od = ctx.empty((6, 54, 32))
    m.calc_total_optical_depth_dev([None if v is None else ctx.to_device(v) for v in vmr], 6, 54,
                                   ctx.to_device(c["pressure_hl"]), ctx.to_device(c["temperature_hl"]), d_od)
    np.testing.assert_allclose(d_od.numpy(), ref, rtol=1e-11, atol=1e-18)


@pytest.mark.parametrize("ncol,nlay,n", [(3, 54, 32), (5, 54, 64), (2, 53, 96), (1, 1, 32), (2, 3, 3008), (1500, 7, 32)])
def test_rt_lw_staged_kernel_shapes(ctx, O, ncol, nlay, n):
    """radiative_transfer_lw through the shared-memory staged kernel (n a multiple of 32): contiguous
    columns (n = 32, two bulk copies), per-row copies (n > 32), layer counts that are not a multiple
    of the 4-layer batch, one layer, and more (column, chunk) items than resident warps.  Exact zeros,
    optical depths under the 1e-5 emissivity threshold and opaque layers included."""
    rng = np.random.default_rng(ncol * 1000 + nlay * 10 + n)
    od = np.exp(rng.normal(-3, 3, (ncol, nlay, n)))
    od[rng.uniform(size=od.shape) < 0.1] = 0.0
    od[:, 0, :4] = 1e-9
    od[:, -1, 4:8] = 300.0
    planck = rng.uniform(0.0, 30.0, (ncol, nlay + 1, n))
    emis, sp = rng.uniform(0.8, 1.0, n), rng.uniform(0.0, 30.0, (ncol, n))
    d = [ctx.to_device(a) for a in (planck, od, emis, sp)]
    dn, up = ctx.empty((ncol, nlay + 1, n)), ctx.empty((ncol, nlay + 1, n))
    E.radiative_transfer_lw_dev(ctx, ncol, nlay, n, d[0], d[1], d[2], d[3], dn, up)
    dn, up = dn.numpy(), up.numpy()
    for c in rng.choice(ncol, min(ncol, 6), replace=False):
        dn_o, up_o = O.radiative_transfer_lw(planck[c], od[c], emis, sp[c])
        assert np.max(np.abs(dn[c] - dn_o)) <= FLUX_ATOL * 1e-3, (c, np.max(np.abs(dn[c] - dn_o)))
        assert np.max(np.abs(up[c] - up_o)) <= FLUX_ATOL * 1e-3
    assert np.all(dn[:, 0, :] == 0.0)


@pytest.mark.parametrize("ncol,nlay,n", [(3, 54, 32), (4, 54, 64), (2, 53, 96), (1, 1, 32), (700, 6, 32)])
@pytest.mark.parametrize("with_up", [True, False])
def test_rt_sw_staged_kernel_shapes(ctx, O, ncol, nlay, n, with_up):
    """radiative_transfer_{direct,norayleigh}_sw through the shared-memory staged kernel (n a multiple of 32),
    per-column cos_sza; deep columns whose direct beam underflows keep their RELATIVE accuracy."""
    rng = np.random.default_rng(ncol * 100 + nlay + n)
    od = np.exp(rng.normal(-2, 2, (ncol, nlay, n)))
    od[rng.uniform(size=od.shape) < 0.1] = 0.0
    od[0, :, :3] = 40.0  # cumulative optical depth in the thousands
    ssi, alb = rng.uniform(0.1, 2.0, n), rng.uniform(0.0, 0.5, n)
    mu = rng.uniform(0.1, 1.0, ncol)
    dn, up = ctx.empty((ncol, nlay + 1, n)), ctx.empty((ncol, nlay + 1, n))
    E.radiative_transfer_sw_dev(ctx, ncol, nlay, n, mu, ctx.to_device(ssi), ctx.to_device(od),
                                ctx.to_device(alb) if with_up else None, dn, up if with_up else None)
    dn = dn.numpy()
    for c in rng.choice(ncol, min(ncol, 5), replace=False):
        if with_up:
            dn_o, up_o = O.radiative_transfer_norayleigh_sw(float(mu[c]), ssi, od[c], alb)
            np.testing.assert_allclose(up.numpy()[c], up_o, rtol=1e-12, atol=1e-300)
        else:
            dn_o = O.radiative_transfer_direct_sw(float(mu[c]), ssi, od[c])
        np.testing.assert_allclose(dn[c], dn_o, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("ng,ncol", [(32, 333), (64, 77), (20, 130), (32, 31)])
def test_run_ckd_layer_major_kernel_and_column_major_kernel_agree(ctx, O, monkeypatch, ng, ncol):
    """total_od_planck_kernel (columns grouped by their table corners, several tiles, partial tiles)
    against the oracle and against the column-major kernel it replaces."""
    c = S.ckd_model_case(ncol=ncol, nlay=23, ng=ng, seed=99 + ng)
    # make neighbouring columns share table corners in some layers and not in others
    c["temperature_hl"][::3] += 25.0
    m = _model(ctx, c)
    ref = O.run_ckd(c, is_sw=False)
    got = m.run_ckd(c["vmr_fl"], c["pressure_hl"], c["temperature_hl"])
    np.testing.assert_allclose(got["optical_depth"], ref["optical_depth"], rtol=2e-11, atol=1e-18)
    np.testing.assert_allclose(got["planck_hl"], ref["planck_hl"], rtol=1e-14)
    np.testing.assert_allclose(got["planck_surf"], ref["planck_surf"], rtol=1e-14)
    for k in ("flux_dn", "flux_up", "bb_flux_dn", "bb_flux_up"):
        assert np.max(np.abs(got[k] - ref[k])) <= FLUX_ATOL, k
    # optical depth only (no Planck output, no sweep)
    d_p, d_t = ctx.to_device(c["pressure_hl"]), ctx.to_device(c["temperature_hl"])
    d_v = [None if v is None else ctx.to_device(v) for v in c["vmr_fl"]]
    d_od = ctx.empty((ncol, 23, ng))
    m.run_ckd_dev(d_v, 1361.0, ncol, 23, d_p, d_t, d_od)
    assert np.array_equal(d_od.numpy(), got["optical_depth"])
    monkeypatch.setenv("ECCKD_RUN_CKD_LEGACY", "1")
    old = m.run_ckd(c["vmr_fl"], c["pressure_hl"], c["temperature_hl"])
    np.testing.assert_allclose(got["optical_depth"], old["optical_depth"], rtol=2e-11, atol=1e-18)
    np.testing.assert_allclose(got["planck_hl"], old["planck_hl"], rtol=1e-14)


def test_run_ckd_two_lut_gases_and_nine_gases(ctx, O):
    """More than one gas with a concentration axis (unpipelined path of the layer-major kernel), and a
    gas list longer than the kernel's register budget (falls to the column-major kernel)."""
    c = S.ckd_model_case(ncol=40, nlay=12, ng=32, seed=5)
    rng = np.random.default_rng(1)
    lut2 = dict(c["gases"][1])
    lut2["molar_abs"] = lut2["molar_abs"] * rng.uniform(0.5, 2.0, lut2["molar_abs"].shape)
    for gases, vmr in ((c["gases"][:5] + [lut2], c["vmr_fl"][:5] + [c["vmr_fl"][1] * 0.7]),
                       (c["gases"] + [c["gases"][2]], c["vmr_fl"] + [c["vmr_fl"][2]])):
        cc = dict(c, gases=gases, vmr_fl=vmr)
        m = _model(ctx, cc)
        got = m.run_ckd(vmr, c["pressure_hl"], c["temperature_hl"])
        ref = O.run_ckd(cc, is_sw=False)
        np.testing.assert_allclose(got["optical_depth"], ref["optical_depth"], rtol=2e-11, atol=1e-18)
        for k in ("flux_dn", "flux_up"):
            assert np.max(np.abs(got[k] - ref[k])) <= FLUX_ATOL, k


@pytest.mark.parametrize("ng", [32, 40])
def test_calc_total_optical_depth_adjoint_is_the_transpose_of_the_forward(ctx, O, ng):
    """calc_total_optical_depth (solve_adept.cpp:23-69, ckd_model.cpp:962-1088) is linear in
    the look-up tables, tau = A K, so its reverse mode is A^T.  For random c and a second random set of tables K':
    <c, A K'> == <A^T c, K'> - the forward is the oracle-checked device call on a model holding K', the adjoint
    never sees K'.  Also: an absent gas gets nothing, accumulate adds, and the call is bitwise reproducible."""
    ncol, nlay = 37, 23
    c = S.ckd_model_case(ncol=ncol, nlay=nlay, ng=ng, seed=29)
    rng = np.random.default_rng(5)
    m = _model(ctx, c, sw=False)
    vmr = list(c["vmr_fl"])
    absent = [i for i, g in enumerate(c["gases"]) if g["conc_dependence"] != E.CONC_NONE][-1]
    vmr[absent] = None  # "not present in LBL file": skipped by the forward, so by its transpose
    d_vmr = [None if v is None else ctx.to_device(v) for v in vmr]
    d_p, d_t = ctx.to_device(c["pressure_hl"]), ctx.to_device(c["temperature_hl"])
    cvec = rng.normal(size=(ncol, nlay, ng))
    d_c = ctx.to_device(cvec)
    grads = [ctx.zeros(np.asarray(g["molar_abs"]).shape) for g in c["gases"]]
    m.calc_total_optical_depth_adjoint_dev(d_vmr, ncol, nlay, d_p, d_t, d_c, grads)
    g1 = [g.numpy() for g in grads]
    assert not np.any(g1[absent])
    # a second model with other tables
    c2 = dict(c)
    c2["gases"] = [dict(g, molar_abs=np.asarray(g["molar_abs"]) * rng.uniform(0.2, 3.0, np.asarray(g["molar_abs"]).shape))
                   for g in c["gases"]]
    m2 = _model(ctx, c2, sw=False)
    d_od = ctx.empty((ncol, nlay, ng))
    m2.calc_total_optical_depth_dev(d_vmr, ncol, nlay, d_p, d_t, d_od)
    lhs = float((cvec * d_od.numpy()).sum())
    rhs = float(sum((a * np.asarray(g["molar_abs"])).sum() for a, g in zip(g1, c2["gases"])))
    scale = float(sum(np.abs(a * np.asarray(g["molar_abs"])).sum() for a, g in zip(g1, c2["gases"])))
    assert abs(lhs - rhs) <= 1e-12 * scale, (lhs, rhs, scale)
    # reproducible, and accumulate = True adds
    m.calc_total_optical_depth_adjoint_dev(d_vmr, ncol, nlay, d_p, d_t, d_c, grads, accumulate=True)
    for a, g in zip(g1, grads):
        assert np.array_equal(g.numpy(), a + a)
    only = [None] * len(grads)
    only[1] = ctx.zeros(g1[1].shape)  # a subset of the gases
    m.calc_total_optical_depth_adjoint_dev(d_vmr, ncol, nlay, d_p, d_t, d_c, only)
    assert np.array_equal(only[1].numpy(), g1[1])
