"""The block-loop state of create_look_up_table on the device (create_look_up_table.cpp:249-591):
ecckd_lut_builder_* and the g-point plan cache of the host-array entry points.  A builder block must be
bit-identical to the one-shot ecckd_create_lut_block on the same inputs (float64 and float32 spectra),
match the oracle, and the plan cache must never serve a plan for a map that differs in a single element."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu


def _fl(p_hl, t_hl):
    p_fl = 0.5 * (p_hl[:-1] + p_hl[1:])
    t_x_p = t_hl * p_hl
    return p_fl, 0.5 * (t_x_p[:-1] + t_x_p[1:]) / p_fl


@pytest.mark.parametrize("sw", [False, True])
def test_builder_blocks_equal_one_shot_blocks_and_the_oracle(ctx, O, sw):
    c = S.sw_spectrum_case(nlay=15, nwav=20011, ng=8, nband=2, seed=8) if sw else \
        S.lw_spectrum_case(nlay=15, nwav=20011, ng=8, nband=2, seed=8)
    method = "transmission-3" if sw else "transmission"
    p_hl = c["pressure_hl"]
    rng = np.random.default_rng(4)
    builder = E.LutBuilder(ctx, c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], c["ssi"] if sw else None)
    for iblock in range(4):  # temperature / concentration blocks: other profile, other spectrum, same map and grid
        t_hl = c["temperature_hl"] + 10.0 * iblock
        od = (c["od"] * rng.uniform(0.3, 3.0)).astype(np.float32).astype(np.float64)  # float32-representable
        vmr = c["vmr"] * (1.0 + iblock)
        p_fl, t_fl = _fl(p_hl, t_hl)
        one = E.create_lut_block(ctx, c["ng"], vmr, p_hl, t_hl, c["g_point"], od, method,
                                 wavenumber_cm_1=c["wavenumber"], d_wavenumber_cm_1=c["d_wavenumber"],
                                 ssi=c["ssi"] if sw else None)
        for dtype in (np.float64, np.float32):
            got = builder.block(vmr, p_hl, t_hl, od.astype(dtype), method)
            for a, b in zip(got[:4], one[:4]):
                assert np.array_equal(a, b), (iblock, dtype)
            assert np.array_equal(got[4], t_fl)
        w = c["ssi"] if sw else O.planck_function(t_fl, c["wavenumber"], c["d_wavenumber"])
        ref = O.average_optical_depth_to_g_point(c["ng"], vmr, p_fl, p_hl, c["g_point"], od, w, method)
        S.assert_avg_close(got[:4], ref, method=method, scale=S.out_scale(dict(vmr=vmr, pressure_fl=p_fl, pressure_hl=p_hl)))
    builder.close()


def test_builder_rejects_bad_arguments(ctx):
    c = S.lw_spectrum_case(nlay=3, nwav=1000, ng=4, nband=1, seed=1)
    with pytest.raises(E.EcckdError):
        E.LutBuilder(ctx, c["ng"], c["g_point"])  # neither a grid nor a solar spectrum
    b = E.LutBuilder(ctx, c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"])
    with pytest.raises(E.EcckdError):
        b.block(c["vmr"], c["pressure_hl"], c["temperature_hl"], c["od"][:, :-1], "transmission")
    with pytest.raises(E.EcckdError):
        b.block(c["vmr"], c["pressure_hl"], c["temperature_hl"], c["od"], "no-such-method")
    b.close()


def test_plan_cache_hits_on_the_same_map_and_misses_on_any_change(ctx, O, monkeypatch):
    monkeypatch.delenv("ECCKD_PLAN_WINDOW", raising=False)
    c = S.lw_spectrum_case(nlay=4, nwav=30001, ng=10, nband=2, seed=21)
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    args = (c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"])

    def run(g, od):
        return E.average_optical_depth_to_g_point(ctx, *args, g, od, planck, "linear")

    g = c["g_point"].copy()
    g[0] = (g[0] + 1) % c["ng"]  # a map no earlier test has used
    run(g, c["od"])
    h0, m0 = ctx.plan_cache_stats()
    a = run(g, c["od"] * 2.0)
    h1, m1 = ctx.plan_cache_stats()
    assert (h1, m1) == (h0 + 1, m0)
    ref = O.average_optical_depth_to_g_point(*args, g, c["od"] * 2.0, planck, "linear")
    S.assert_avg_close(a, ref, method="linear", scale=S.out_scale(c))
    # one element changed, at the very end of the map: must be seen, and the result must be that of the new map
    g2 = g.copy()
    g2[-1] = (g2[-1] + 1) % c["ng"]
    b = run(g2, c["od"])
    h2, m2 = ctx.plan_cache_stats()
    assert (h2, m2) == (h1, m1 + 1)
    ref = O.average_optical_depth_to_g_point(*args, g2, c["od"], planck, "linear")
    S.assert_avg_close(b, ref, method="linear", scale=S.out_scale(c))
    # the plan-shaping settings are part of the key
    monkeypatch.setenv("ECCKD_PLAN_WINDOW", "0")
    run(g2, c["od"])
    assert ctx.plan_cache_stats() == (h2, m2 + 1)
