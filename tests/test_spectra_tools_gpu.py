"""SURVEY.md 8f item 4, the adjacent tools that reuse the hot-path kernels: the per-profile body of lw_spectra
(lw_spectra.cpp:208-255) and the flux sums / scaling of scale_lut (scale_lut.cpp:117-133, 180-181), against the
oracle's planck_function, radiative_transfer_lw and average_optical_depth_to_g_point (each pinned to the
reference's own translation unit, tests/test_reference_sources.py) composed as the tool composes them."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu


def _seg_sum(x, g, ng):
    out = np.zeros((x.shape[0], ng))
    for ig in range(ng):
        out[:, ig] = x[:, g == ig].sum(axis=1)
    return out


@pytest.mark.parametrize("nwav", [4096, 10007])
def test_lw_spectra_profile_matches_the_tool_composition(ctx, O, nwav):
    c = S.lw_spectrum_case(nlay=11, nwav=nwav, ng=9, nband=3, seed=8)
    g, ng, nlay = c["g_point"].copy(), c["ng"], 11
    g[::13] = -1  # wavenumbers outside every g point still count in the broadband sums
    p_hl = c["pressure_hl"]
    t_hl = 200.0 + 88.0 * (np.log(p_hl) / np.log(p_hl[-1]))
    p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
    t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
    wn, dwn, od = c["wavenumber"], c["d_wavenumber"], c["od"]
    # the tool: lw_spectra.cpp:208-255
    planck_hl = O.planck_function(t_hl, wn, dwn)
    dn, up = O.radiative_transfer_lw(planck_hl, od, np.ones(nwav), planck_hl[-1])
    ref_od = O.average_optical_depth_to_g_point(ng, 0.0, p_fl, p_hl, g, od, O.planck_function(t_fl, wn, dwn),
                                                "transmission")
    d_g = ctx.to_device(g, np.int32)
    plan = E.GPointPlan(ctx, d_g, nwav, ng)
    got = E.lw_spectra_profile(ctx, plan, ctx.to_device(wn), ctx.to_device(dwn), p_hl, t_hl, ctx.to_device(od), nwav)
    scale = np.abs(dn).sum(axis=1).max()
    np.testing.assert_allclose(got["flux_dn_lw"], dn.sum(axis=1), rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(got["flux_up_lw"], up.sum(axis=1), rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(got["spectral_flux_dn_lw"], _seg_sum(dn, g, ng), rtol=1e-11, atol=1e-12 * scale)
    np.testing.assert_allclose(got["spectral_flux_up_lw"], _seg_sum(up, g, ng), rtol=1e-11, atol=1e-12 * scale)
    S.assert_avg_close((got["optical_depth"], None, None, None), ref_od, method="transmission",
                       what="lw_spectra optical_depth: ")
    plan.close()


def test_gpoint_sum_and_row_sum_are_exact_on_integers_and_reproducible(ctx):
    rng = np.random.default_rng(3)
    nwav, ng, nrow = 30011, 17, 5
    g = rng.integers(-1, ng, size=nwav).astype(np.int32)
    x = rng.integers(-1000, 1000, size=(nrow, nwav)).astype(np.float64)  # integer sums are exact in any order
    d_x = ctx.to_device(x)
    plan = E.GPointPlan(ctx, ctx.to_device(g, np.int32), nwav, ng)
    out, out2, rs = ctx.empty((nrow, ng)), ctx.empty((nrow, ng)), ctx.empty(nrow)
    E.gpoint_sum_dev(ctx, plan, nrow, d_x, nwav, out)
    E.gpoint_sum_dev(ctx, plan, nrow, d_x, nwav, out2)
    E.row_sum_dev(ctx, nrow, nwav, d_x, nwav, rs)
    assert np.array_equal(out.numpy(), _seg_sum(x, g, ng))
    assert np.array_equal(out.numpy(), out2.numpy())
    assert np.array_equal(rs.numpy(), x.sum(axis=1))
    plan.close()
