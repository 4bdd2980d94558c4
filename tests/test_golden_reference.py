"""Golden vectors produced by the reference's own compiled sources
(tests/golden/reference_sources.npz, written by tests/golden/make_golden_reference.py from
oracle/_ref/libecckd_ref.so).  The CPU oracle must reproduce them anywhere (no GPU, no
/root/reference needed); the CUDA path must match them on the B200.
"""
import os

import numpy as np
import pytest

from tests import synthetic as S
from tests.oracle_lib import METHODS

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sources.npz"))
FLUX_ATOL = 1e-6   # W m-2 (north_star)
COST_RTOL = 1e-11


def _avg(fn, method):
    return fn(int(G["avg_ng"]), float(G["avg_vmr"]), G["avg_pressure_fl"], G["avg_pressure_hl"], G["avg_g_point"],
              G["avg_od"], G["avg_planck"], method)


def _gold_avg(method):
    return tuple(G[f"avg_{method}__{k}"] for k in ("molar_abs", "min", "max"))


def _scale():
    return S.out_scale(dict(pressure_fl=G["avg_pressure_fl"], pressure_hl=G["avg_pressure_hl"], vmr=float(G["avg_vmr"])))


@pytest.mark.parametrize("method", METHODS)
def test_oracle_average(O, method):
    got = _avg(O.average_optical_depth_to_g_point, method)
    S.assert_avg_close(got[:3], _gold_avg(method), method=method, scale=_scale(), rtol=2e-13)


def test_oracle_planck_and_flux_solvers(O):
    np.testing.assert_allclose(O.planck_function(G["avg_temperature_fl"], G["avg_wavenumber"], G["avg_d_wavenumber"]),
                               G["avg_planck"], rtol=1e-13)
    od, pl, emis, surf, grey, p, ssi, alb = (G[k] for k in ("rt_od", "rt_planck", "rt_emis", "rt_surf", "rt_grey", "rt_p",
                                                            "rt_ssi", "rt_alb"))
    dn, up = O.radiative_transfer_lw(pl, od, emis, surf)
    np.testing.assert_allclose(dn, G["rt_lw_dn"], rtol=1e-13), np.testing.assert_allclose(up, G["rt_lw_up"], rtol=1e-13)
    np.testing.assert_allclose(O.heating_rate(p, G["rt_lw_dn"], G["rt_lw_up"]), G["rt_hr"], rtol=1e-13)
    np.testing.assert_allclose(O.heating_rate(p, G["rt_lw_dn"], None), G["rt_hr_dn_only"], rtol=1e-13)
    a = O.radiative_transfer_lw_bb(pl, od, grey, emis, surf)
    np.testing.assert_allclose(a[0], G["rt_lw_bb_dn"], rtol=1e-13), np.testing.assert_allclose(a[1], G["rt_lw_bb_up"], rtol=1e-13)
    np.testing.assert_allclose(O.radiative_transfer_direct_sw(0.37, ssi, od), G["rt_sw_direct"], rtol=1e-13)
    a = O.radiative_transfer_norayleigh_sw(0.5, ssi, od, alb)
    np.testing.assert_allclose(a[0], G["rt_sw_dn"], rtol=1e-13), np.testing.assert_allclose(a[1], G["rt_sw_up"], rtol=1e-13)
    a = O.radiative_transfer_sw_bb(0.5, ssi, od, grey * 0.1, 0.15)
    np.testing.assert_allclose(a[0], G["rt_sw_bb_dn"], rtol=1e-13), np.testing.assert_allclose(a[1], G["rt_sw_bb_up"], rtol=1e-13)
    np.testing.assert_allclose(O.rayleigh_molar_scattering_coeff(G["ray_wn"]), G["ray_coeff"], rtol=1e-13)


def _cost_lw_args():
    return (G["fc_p"], G["fc_planck"], G["fc_emis"], G["fc_surf"], G["fc_od"], G["fc_dn"], G["fc_up"], G["fc_hr"], 0.2, 0.2,
            0.8, 0.1)


def test_oracle_costs_and_g_points(O):
    e = O.calc_cost_function_lw(G["ie_p"], G["ie_planck"], G["ie_emis"], G["ie_surf"], G["ie_od"], G["ie_fit"],
                                G["ie_dn_surf"], G["ie_up_toa"], G["ie_hr"], 0.2, G["ie_lw"])
    assert abs(e - float(G["ie_cost_lw"])) <= 1e-12 * e
    for alb in (0.0, 0.15):
        e = O.calc_cost_function_sw(0.5, G["ie_p"], G["ie_ssi"], alb, G["ie_od"], G["ie_fit"] * 0.1, G["ie_sw_dn_surf"],
                                    G["ie_sw_up_toa"], G["ie_sw_hr"], 0.2, G["ie_lw"])
        assert abs(e - float(G[f"ie_cost_sw_{alb}"])) <= 1e-12 * e
    c = O.calc_cost_function_ckd_lw(*_cost_lw_args(), G["fc_lw"], spectral_flux_dn_surf=G["fc_sfd"],
                                    spectral_flux_up_toa=G["fc_sfu"], band_mapping=G["fc_bm"])
    assert abs(c - float(G["fc_cost_lw"])) <= 1e-12 * c
    pb = np.zeros(4)
    c = O.calc_cost_function_ckd_sw(0.6, G["fc_p"], G["fc_ssi"], G["fc_albedo"], G["fc_od"], G["fc_sw_dn"], G["fc_sw_up"],
                                    G["fc_sw_hr"], 0.2, 0.2, 0.8, G["fc_lw"], spectral_boundary_weights=G["fc_sbw"],
                                    spectral_flux_dn_surf=G["fc_sfd"], spectral_flux_up_toa=G["fc_sfu"],
                                    band_mapping=G["fc_bm"], cost_fn_per_band=pb)
    assert abs(c - float(G["fc_cost_sw"])) <= 1e-12 * c
    np.testing.assert_allclose(pb, G["fc_cost_sw_per_band"], rtol=1e-12)
    assert np.array_equal(O.store_g_points(G["gp_rank"], G["gp_rank1"], G["gp_rank2"]), G["gp_g_point"])
    sv, off = [], 0
    for n in G["gp_n_g"].sum(1):
        sv.append(G["gp_sv"][off:off + n])
        off += n
    ng, gmin, gmax, band = O.overlap_g_points(G["gp_n_g"], sv)
    assert ng == int(G["gp_ng"]) and np.array_equal(gmin, G["gp_gmin"]) and np.array_equal(gmax, G["gp_gmax"])
    assert np.array_equal(band, G["gp_band"])


# ------------------------------------------------------------------ CUDA path against the same vectors
@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
def test_cuda_average(ctx, method):
    import ecckd_b200 as E

    got = _avg(lambda *a: E.average_optical_depth_to_g_point(ctx, *a), method)
    S.assert_avg_close(got[:3], _gold_avg(method), method=method, scale=_scale())
    # fused Planck weights (create_look_up_table.cpp:320-331 call pair in one launch)
    got = E.planck_average_optical_depth_to_g_point(ctx, int(G["avg_ng"]), float(G["avg_vmr"]), G["avg_pressure_fl"],
                                                    G["avg_pressure_hl"], G["avg_g_point"], G["avg_od"],
                                                    G["avg_temperature_fl"], G["avg_wavenumber"], G["avg_d_wavenumber"],
                                                    method)
    S.assert_avg_close(got[:3], _gold_avg(method), method=method, scale=_scale(), planck_scale=G["avg_planck"])


@pytest.mark.gpu
def test_cuda_planck_and_flux_solvers(ctx):
    import ecckd_b200 as E

    np.testing.assert_allclose(E.planck_function(ctx, G["avg_temperature_fl"], G["avg_wavenumber"], G["avg_d_wavenumber"]),
                               G["avg_planck"], rtol=1e-12)
    od, pl, emis, surf, grey, p, ssi, alb = (G[k] for k in ("rt_od", "rt_planck", "rt_emis", "rt_surf", "rt_grey", "rt_p",
                                                            "rt_ssi", "rt_alb"))
    dn, up = E.radiative_transfer_lw(ctx, pl, od, emis, surf)
    assert np.max(np.abs(dn - G["rt_lw_dn"])) <= FLUX_ATOL and np.max(np.abs(up - G["rt_lw_up"])) <= FLUX_ATOL
    np.testing.assert_allclose(E.heating_rate(ctx, p, G["rt_lw_dn"], G["rt_lw_up"]), G["rt_hr"], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(E.heating_rate(ctx, p, G["rt_lw_dn"], None), G["rt_hr_dn_only"], rtol=1e-12, atol=1e-18)
    a = E.radiative_transfer_lw_bb(ctx, pl, od, grey, emis, surf)
    assert np.max(np.abs(a[0] - G["rt_lw_bb_dn"])) <= FLUX_ATOL and np.max(np.abs(a[1] - G["rt_lw_bb_up"])) <= FLUX_ATOL
    assert np.max(np.abs(E.radiative_transfer_direct_sw(ctx, 0.37, ssi, od) - G["rt_sw_direct"])) <= FLUX_ATOL
    a = E.radiative_transfer_norayleigh_sw(ctx, 0.5, ssi, od, alb)
    assert np.max(np.abs(a[0] - G["rt_sw_dn"])) <= FLUX_ATOL and np.max(np.abs(a[1] - G["rt_sw_up"])) <= FLUX_ATOL
    a = E.radiative_transfer_sw_bb(ctx, 0.5, ssi, od, grey * 0.1, 0.15)
    assert np.max(np.abs(a[0] - G["rt_sw_bb_dn"])) <= FLUX_ATOL and np.max(np.abs(a[1] - G["rt_sw_bb_up"])) <= FLUX_ATOL
    a = E.radiative_transfer_sw_bb(ctx, 0.5, ssi, od, grey * 0.1, 0.15, direct_only=True)
    assert np.max(np.abs(a[0] - G["rt_sw_bb_direct"])) <= FLUX_ATOL
    np.testing.assert_allclose(E.rayleigh_molar_scattering_coeff(ctx, G["ray_wn"]), G["ray_coeff"], rtol=1e-13)


@pytest.mark.gpu
def test_cuda_costs_and_g_points(ctx):
    import ecckd_b200 as E

    n = G["ie_od"].shape[1]
    # the interval error with the reference's fitted profile: transmission metric whose fit is ie_fit
    # cannot be inverted in general, so drive the flux + error stages directly through the full-band
    # interval with method "linear" on a metric equal to the fit (constant over the band)
    metric = np.broadcast_to(G["ie_fit"][:, None], G["ie_od"].shape).copy()
    err, fit = E.interval_error_batch_lw(ctx, "linear", G["ie_p"], G["ie_planck"], G["ie_emis"], G["ie_surf"], G["ie_od"],
                                         metric, G["ie_dn_surf"], G["ie_up_toa"], G["ie_hr"], 0.2, G["ie_lw"], [0], [n - 1])
    np.testing.assert_allclose(fit[0], G["ie_fit"], rtol=1e-13)
    assert abs(err[0] - float(G["ie_cost_lw"])) <= COST_RTOL * float(G["ie_cost_lw"])
    for alb in (0.0, 0.15):
        err, _ = E.interval_error_batch_sw(ctx, "linear", 0.5, G["ie_p"], G["ie_ssi"], alb, G["ie_od"], metric * 0.1,
                                           G["ie_sw_dn_surf"], G["ie_sw_up_toa"], G["ie_sw_hr"], 0.2, G["ie_lw"], [0],
                                           [n - 1])
        assert abs(err[0] - float(G[f"ie_cost_sw_{alb}"])) <= COST_RTOL * float(G[f"ie_cost_sw_{alb}"])
    one = lambda a: np.asarray(a)[None]  # noqa: E731  (a scene of one profile)
    r = E.cost_function_ckd_lw(ctx, one(G["fc_p"]), one(G["fc_planck"]), one(G["fc_emis"]), one(G["fc_surf"]), one(G["fc_od"]),
                               one(G["fc_dn"]), one(G["fc_up"]), one(G["fc_hr"]), 0.2, 0.2, 0.8, 0.1,
                               layer_weight=one(G["fc_lw"]), spectral_flux_dn_surf=one(G["fc_sfd"]),
                               spectral_flux_up_toa=one(G["fc_sfu"]), band_mapping=G["fc_bm"])
    assert abs(r["cost_profile"][0] - float(G["fc_cost_lw"])) <= COST_RTOL * float(G["fc_cost_lw"])
    r = E.cost_function_ckd_sw(ctx, 0.6, one(G["fc_p"]), G["fc_ssi"], G["fc_albedo"], one(G["fc_od"]), one(G["fc_sw_dn"]),
                               one(G["fc_sw_up"]), one(G["fc_sw_hr"]), 0.2, 0.2, 0.8, layer_weight=one(G["fc_lw"]),
                               spectral_boundary_weights=G["fc_sbw"], spectral_flux_dn_surf=one(G["fc_sfd"]),
                               band_mapping=G["fc_bm"], cost_fn_per_band=np.zeros(4))
    assert abs(r["cost_profile"][0] - float(G["fc_cost_sw"])) <= COST_RTOL * float(G["fc_cost_sw"])
    np.testing.assert_allclose(r["cost_fn_per_band"], G["fc_cost_sw_per_band"], rtol=COST_RTOL)
    # integer g-point logic, bit-exact
    nw = G["gp_rank"].size
    d_g = ctx.empty(nw, np.int32)
    E.store_g_points_dev(ctx, nw, ctx.to_device(G["gp_rank"], np.int32), G["gp_rank1"], G["gp_rank2"], d_g)
    assert np.array_equal(d_g.numpy(), G["gp_g_point"])
    sv, off = [], 0
    for k in G["gp_n_g"].sum(1):
        sv.append(G["gp_sv"][off:off + k])
        off += k
    ng, gmin, gmax, band = E.overlap_g_points(G["gp_n_g"], sv)
    assert ng == int(G["gp_ng"]) and np.array_equal(gmin, G["gp_gmin"]) and np.array_equal(gmax, G["gp_gmax"])
    assert np.array_equal(band, G["gp_band"])
