"""GPU parity for the rows of SURVEY.md section 8 that build on the flux solvers: broadband
solvers (a15), the batched candidate-interval error of find_g_points (a13), the forward cost of
optimize_lut (a21-a23), the base g-point split (a4b) and the Rayleigh coefficients (a7b).

Every CUDA result is held against the CPU oracle on the same seeded inputs; the oracle itself is
held against the reference's own compiled sources in tests/test_reference_sources.py.
Tolerances: fluxes <= 1e-6 W m-2 (north_star), errors / costs relative 1e-11 (they are sums of
squares of those fluxes, summed in a different but fixed order), integer outputs bit-exact.
"""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S
from tests.test_reference_sources import ckd_cost_case, interval_case, rt_inputs

pytestmark = pytest.mark.gpu

FLUX_ATOL = 1e-6
COST_RTOL = 1e-11


def test_radiative_transfer_bb(ctx, O):
    od, planck, emis, surf = rt_inputs(nlay=54, n=3001, seed=1)
    grey = np.exp(np.random.default_rng(2).normal(-3, 2, od.shape[0]))
    grey[3] = 0.0
    a = E.radiative_transfer_lw_bb(ctx, planck, od, grey, emis, surf)
    b = O.radiative_transfer_lw_bb(planck, od, grey, emis, surf)
    assert np.max(np.abs(a[0] - b[0])) <= FLUX_ATOL and np.max(np.abs(a[1] - b[1])) <= FLUX_ATOL
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12)
    rng = np.random.default_rng(4)
    ssi = rng.uniform(0.1, 3.0, od.shape[1])
    for direct, albedo in ((False, 0.15), (True, 0.15), (False, 0.0)):
        a = E.radiative_transfer_sw_bb(ctx, 0.5, ssi, od, grey * 0.1, albedo, direct_only=direct)
        b = O.radiative_transfer_sw_bb(0.5, ssi, od, grey * 0.1, albedo, direct_only=direct)
        np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-9)
        if not direct:
            np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-9)


def test_interval_bounds_to_indices(ctx):
    i1, i2 = E.interval_bounds_to_indices(1000, [0.0, 0.25, 0.5001, 0.3], [0.2499, 0.5, 0.5002, 1.0])
    assert list(i1) == [0, 250, 500, 300] and list(i2) == [249, 499, 500, 999]  # empty range collapses to ibound1
    for b1, b2 in (([-0.1], [0.5]), ([0.2], [1.1]), ([0.6], [0.5])):
        with pytest.raises(E.EcckdError) as ei:
            E.interval_bounds_to_indices(1000, b1, b2)
        assert ei.value.code == 148  # PROCESSING_ERROR, find_g_points.cpp:297-312


def _intervals(n, rng, k=9):
    cuts = np.sort(rng.choice(np.arange(1, n - 1), k - 1, replace=False))
    i1 = np.concatenate([[0], cuts])
    i2 = np.concatenate([cuts - 1, [n - 1]])
    # plus overlapping candidates, a single point and the whole band
    i1 = np.concatenate([i1, [5, n // 3, 0]])
    i2 = np.concatenate([i2, [5, n - 7, n - 1]])
    return i1.astype(np.int64), i2.astype(np.int64)


@pytest.mark.parametrize("method", ["linear", "transmission", "transmission-2", "square-root", "logarithmic"])
def test_interval_error_batch_lw(ctx, O, method):
    c = interval_case(nlay=20, n=3000, seed=7)
    rng = c["rng"]
    true_dn, true_up = O.radiative_transfer_lw(c["planck"], c["od"] * 1.3, c["emis"], c["surf"])
    hr = O.heating_rate(c["p"], true_dn, true_up)
    od_gas = c["od"] * rng.uniform(0.1, 2.0, c["od"].shape)
    if method in ("transmission", "transmission-2"):
        kd = 1.66 * (2.0 if method.endswith("2") else 1.0)
        metric = 1.0 - np.exp(-kd * od_gas)
    elif method == "square-root":
        metric = np.sqrt(od_gas)
    else:
        metric = od_gas
    i1, i2 = _intervals(c["od"].shape[1], rng)
    err, fit = E.interval_error_batch_lw(ctx, method, c["p"], c["planck"], c["emis"], c["surf"], c["od"], metric,
                                         true_dn[-1], true_up[0], hr, 0.2, c["lw"], i1, i2)
    for k, (a, b) in enumerate(zip(i1, i2)):
        f = O.fit_optical_depth_lw(method, a, b, c["planck"], metric)
        np.testing.assert_allclose(fit[k], f, rtol=1e-11, atol=1e-300, err_msg=f"fit {k}")
        sl = slice(a, b + 1)
        e = O.calc_cost_function_lw(c["p"], c["planck"][:, sl], c["emis"][sl], c["surf"][sl], c["od"][:, sl], f,
                                    true_dn[-1, sl], true_up[0, sl], hr[:, sl], 0.2, c["lw"])
        assert abs(err[k] - e) <= COST_RTOL * abs(e) + 1e-12, (k, err[k], e)


def test_interval_error_batch_lw_errors(ctx):
    c = interval_case(nlay=5, n=300, seed=1)
    z = np.zeros(300)
    args = (c["p"], c["planck"], c["emis"], c["surf"], c["od"], c["od"], z, z, c["od"], 0.2, c["lw"])
    with pytest.raises(E.EcckdError) as ei:
        E.interval_error_batch_lw(ctx, "transmission-3", *args, [0], [10])
    assert ei.value.code == 147  # method not understood, find_g_points.cpp:111-113
    with pytest.raises(E.EcckdError) as ei:
        E.interval_error_batch_lw(ctx, "linear", *args, [0], [300])
    assert ei.value.code == 148


@pytest.mark.parametrize("method", ["linear", "transmission", "transmission-2", "square-root", "logarithmic"])
@pytest.mark.parametrize("albedo", [0.0, 0.15])
def test_interval_error_batch_sw(ctx, O, method, albedo):
    c = interval_case(nlay=20, n=2500, seed=9)
    rng = c["rng"]
    n = c["od"].shape[1]
    ssi = rng.uniform(0.1, 2.0, n) * 1e-3
    true_dn, true_up = O.radiative_transfer_norayleigh_sw(0.5, ssi, c["od"] * 1.2, np.full(n, 0.15))
    hr = O.heating_rate(c["p"], true_dn, None)
    od_gas = c["od"] * rng.uniform(0.1, 2.0, c["od"].shape) * 0.05
    if method in ("transmission", "transmission-2"):
        metric = 1.0 - np.exp(-1.66 * (2.0 if method.endswith("2") else 1.0) * od_gas)
    elif method == "square-root":
        metric = np.sqrt(od_gas)
    else:
        metric = od_gas
    i1, i2 = _intervals(n, rng)
    err, fit = E.interval_error_batch_sw(ctx, method, 0.5, c["p"], ssi, albedo, c["od"], metric, true_dn[-1],
                                         true_up[0], hr, 0.2, c["lw"], i1, i2)
    for k, (a, b) in enumerate(zip(i1, i2)):
        f = O.fit_optical_depth_sw(method, a, b, ssi, metric)
        np.testing.assert_allclose(fit[k], f, rtol=1e-11, atol=1e-300, err_msg=f"fit {k}")
        sl = slice(a, b + 1)
        e = O.calc_cost_function_sw(0.5, c["p"], ssi[sl], albedo, c["od"][:, sl], f, true_dn[-1, sl], true_up[0, sl],
                                    hr[:, sl], 0.2, c["lw"])
        assert abs(err[k] - e) <= COST_RTOL * abs(e) + 1e-12, (k, err[k], e)


def test_interval_error_batch_sw_total_transmission(ctx, O):
    c = interval_case(nlay=20, n=2500, seed=10)
    rng = c["rng"]
    n = c["od"].shape[1]
    ssi = rng.uniform(0.1, 2.0, n) * 1e-3
    od_gas = c["od"] * rng.uniform(0.1, 2.0, c["od"].shape) * 0.3
    bg = c["od"] * 0.1
    bg[:, 40:80] += 400.0  # an interval whose direct beam underflows: the linear fallback (find_g_points.cpp:196-198)
    mins, maxs = 0.5, 2.5
    sets = []
    for scal in (mins, maxs):
        dn, up = O.radiative_transfer_norayleigh_sw(0.5, ssi, bg + scal * od_gas, np.full(n, 0.15))
        sets.append((dn[-1], up[0], O.heating_rate(c["p"], dn, None)))
    i1, i2 = _intervals(n, rng)
    i1 = np.concatenate([i1, [40]])
    i2 = np.concatenate([i2, [79]])
    err, fit = E.interval_error_batch_sw(ctx, "total-transmission", 0.5, c["p"], ssi, 0.15, bg, od_gas, None, None, None,
                                         0.2, c["lw"], i1, i2, low=sets[0], high=sets[1], min_scaling=mins,
                                         max_scaling=maxs)
    for k, (a, b) in enumerate(zip(i1, i2)):
        f = O.fit_optical_depth_sw_total_trans(a, b, ssi, bg, od_gas)
        np.testing.assert_allclose(fit[k], f, rtol=1e-10, atol=1e-14, err_msg=f"fit {k}")
        sl = slice(a, b + 1)
        e = 0.0
        for scal, st in zip((mins, maxs), sets):
            e += 0.5 * O.calc_cost_function_sw(0.5, c["p"], ssi[sl], 0.15, bg[:, sl], f * scal, st[0][sl], st[1][sl],
                                               st[2][:, sl], 0.2, c["lw"])
        assert abs(err[k] - e) <= 1e-9 * abs(e) + 1e-12, (k, err[k], e)


def _scene(nprof, nlay, ng, nband, seed, sw=False):
    rng = np.random.default_rng(seed)
    cases = [ckd_cost_case(nlay=nlay, ng=ng, nband=nband, seed=seed * 100 + i) for i in range(nprof)]
    bm = cases[0]["bm"]
    d = dict(p=np.stack([c["p"] for c in cases]), od=np.stack([c["od"] for c in cases]),
             planck=np.stack([c["planck"] for c in cases]), surf=np.stack([c["surf"] for c in cases]), bm=bm, rng=rng)
    d["od"][0, 2, 3] = -0.01  # negative optical depths: penalty + clamp (solve_adept.cpp:107-116)
    d["od"][nprof - 1, 0, 0] = -0.2
    return d


@pytest.mark.parametrize("banded", [True, False])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8, 0.1), (0.5, 0.0, 0.0, 0.0)])
@pytest.mark.parametrize("power", [0.5, 1.0, 0.3])
def test_cost_function_ckd_lw(ctx, O, banded, weights, power):
    nprof, nlay, ng, nb = 7, 54, 64, 13
    s = _scene(nprof, nlay, ng, nb, seed=3)
    rng = s["rng"]
    nband = nb if banded else ng
    bm = s["bm"] if banded else None
    emis = rng.uniform(0.9, 1.0, (nprof, nband))
    t_dn, t_up, hr, lws = [], [], [], []
    for i in range(nprof):
        eg = emis[i][bm] if banded else emis[i]
        dn, up = O.radiative_transfer_lw(s["planck"][i], np.maximum(s["od"][i], 0) * 1.1, eg, s["surf"][i])
        if banded:
            dn = np.stack([dn[:, bm == b].sum(1) for b in range(nband)], 1)
            up = np.stack([up[:, bm == b].sum(1) for b in range(nband)], 1)
        t_dn.append(dn), t_up.append(up), hr.append(O.heating_rate(s["p"][i], dn, up))
        lw = np.diff(s["p"][i] ** power) if power != 0.5 else np.diff(np.sqrt(s["p"][i]))
        lws.append(lw / lw.sum())
    t_dn, t_up, hr = np.stack(t_dn), np.stack(t_up), np.stack(hr)
    sfd, sfu = rng.uniform(0, 5, (nprof, ng)), rng.uniform(0, 5, (nprof, ng))
    rel = rng.uniform(0, 1, (2, nprof, nlay + 1, ng))
    got = E.cost_function_ckd_lw(ctx, s["p"], s["planck"], emis, s["surf"], s["od"], t_dn, t_up, hr, *weights,
                                 pressure_weight_power=power, negative_od_penalty=1.0e4, spectral_flux_dn_surf=sfd,
                                 spectral_flux_up_toa=sfu, rel_dn=rel[0], rel_up=rel[1], band_mapping=bm)
    ref = [O.calc_cost_function_ckd_lw(s["p"][i], s["planck"][i], emis[i], s["surf"][i], np.maximum(s["od"][i], 0.0),
                                       t_dn[i], t_up[i], hr[i], *weights, lws[i], spectral_flux_dn_surf=sfd[i],
                                       spectral_flux_up_toa=sfu[i], rel_dn=rel[0][i], rel_up=rel[1][i], band_mapping=bm)
           for i in range(nprof)]
    np.testing.assert_allclose(got["cost_profile"], ref, rtol=COST_RTOL)
    pen = 1.0e4 * float((np.minimum(s["od"], 0.0) ** 2).sum())
    assert abs(got["penalty"] - pen) <= 1e-12 * pen
    assert abs(got["cost_total"] - (pen + sum(ref))) <= COST_RTOL * (pen + sum(ref))
    # explicit layer weights give the same answer as the in-kernel ones
    got2 = E.cost_function_ckd_lw(ctx, s["p"], s["planck"], emis, s["surf"], s["od"], t_dn, t_up, hr, *weights,
                                  layer_weight=np.stack(lws), spectral_flux_dn_surf=sfd, spectral_flux_up_toa=sfu,
                                  rel_dn=rel[0], rel_up=rel[1], band_mapping=bm)
    np.testing.assert_allclose(got2["cost_profile"], ref, rtol=COST_RTOL)


@pytest.mark.parametrize("albedo_on", [True, False])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8), (0.5, 0.0, 0.0)])
def test_cost_function_ckd_sw(ctx, O, albedo_on, weights):
    nprof, nlay, ng, nband = 6, 54, 32, 9
    s = _scene(nprof, nlay, ng, nband, seed=5, sw=True)
    rng, bm = s["rng"], s["bm"]
    ssi = rng.uniform(1.0, 100.0, ng)
    albedo = rng.uniform(0.05, 0.5, nband) if albedo_on else np.zeros(nband)
    mu = rng.uniform(0.2, 1.0, nprof)
    t_dn, t_up, hr, lws = [], [], [], []
    for i in range(nprof):
        dn, up = O.radiative_transfer_norayleigh_sw(mu[i], ssi, np.maximum(s["od"][i], 0) * 1.1, albedo[bm])
        dn = np.stack([dn[:, bm == b].sum(1) for b in range(nband)], 1)
        up = np.stack([up[:, bm == b].sum(1) for b in range(nband)], 1)
        t_dn.append(dn), t_up.append(up), hr.append(O.heating_rate(s["p"][i], dn, None))
        lw = np.diff(np.sqrt(s["p"][i]))
        lws.append(lw / lw.sum())
    t_dn, t_up, hr = np.stack(t_dn), np.stack(t_up), np.stack(hr)
    sbw, sfd = rng.uniform(0, 1, ng), rng.uniform(0, 5, (nprof, ng))
    per_band0 = rng.uniform(0, 1, nband)
    got = E.cost_function_ckd_sw(ctx, mu, s["p"], ssi, albedo, s["od"], t_dn, t_up, hr, *weights,
                                 spectral_boundary_weights=sbw, spectral_flux_dn_surf=sfd, band_mapping=bm,
                                 cost_fn_per_band=per_band0)
    per_band = per_band0.copy()
    ref = [O.calc_cost_function_ckd_sw(mu[i], s["p"][i], ssi, albedo, np.maximum(s["od"][i], 0.0), t_dn[i], t_up[i],
                                       hr[i], *weights, lws[i], spectral_boundary_weights=sbw,
                                       spectral_flux_dn_surf=sfd[i], spectral_flux_up_toa=sfd[i], band_mapping=bm,
                                       cost_fn_per_band=per_band) for i in range(nprof)]
    np.testing.assert_allclose(got["cost_profile"], ref, rtol=COST_RTOL)
    np.testing.assert_allclose(got["cost_fn_per_band"], per_band, rtol=COST_RTOL)


def test_cost_function_scenes_batch_equals_scene_by_scene(ctx, O):
    """ecckd_cost_function_ckd_scenes_dev (the scene loop of solve_adept.cpp:96-199 in one launch) against one call
    of the per-scene entry points per scene (themselves held against the oracle above): same kernel body, so the
    profile costs are bitwise equal; the totals are the sums in scene order.  Scenes of different sizes, LW and SW."""
    nlay, ng, nband = 20, 32, 5
    weights = (0.2, 0.2, 0.8, 0.1)
    nprofs = [3, 1, 6, 4]
    # ---- longwave
    scenes, singles = [], []
    for k, nprof in enumerate(nprofs):
        s = _scene(nprof, nlay, ng, nband, seed=40 + k)
        rng = s["rng"]
        bm = s["bm"] if k == 0 else bm  # one band mapping for all scenes, as in solve_adept
        emis = rng.uniform(0.9, 1.0, (nprof, nband))
        dn, up, hr = rng.uniform(0, 300, (nprof, nlay + 1, nband)), rng.uniform(0, 300, (nprof, nlay + 1, nband)), \
            rng.uniform(-1e-4, 0, (nprof, nlay, nband))
        sfd, sfu = rng.uniform(0, 5, (nprof, ng)), rng.uniform(0, 5, (nprof, ng))
        singles.append(E.cost_function_ckd_lw(ctx, s["p"], s["planck"], emis, s["surf"], s["od"], dn, up, hr, *weights,
                                              spectral_flux_dn_surf=sfd, spectral_flux_up_toa=sfu, band_mapping=bm))
        cp = ctx.empty(nprof)
        scenes.append(dict(nprof=nprof, pressure_hl=ctx.to_device(s["p"]), optical_depth=ctx.to_device(s["od"]),
                           flux_dn=ctx.to_device(dn), flux_up=ctx.to_device(up), hr=ctx.to_device(hr),
                           spectral_flux_dn_surf=ctx.to_device(sfd), spectral_flux_up_toa=ctx.to_device(sfu),
                           planck_hl=ctx.to_device(s["planck"]), surf_emissivity=ctx.to_device(emis),
                           surf_planck=ctx.to_device(s["surf"]), cost_profile=cp))
    d_bm = ctx.to_device(bm, np.int32)
    d_cs, d_ps, d_ct = ctx.empty(len(nprofs)), ctx.empty(len(nprofs)), ctx.empty(1)
    E.cost_function_ckd_scenes_dev(ctx, False, scenes, nlay, ng, nband, *weights, d_band_mapping=d_bm,
                                   d_cost_scene=d_cs, d_penalty_scene=d_ps, d_cost_total=d_ct)
    for k, one in enumerate(singles):
        assert np.array_equal(scenes[k]["cost_profile"].numpy(), one["cost_profile"])
        assert d_cs.numpy()[k] == one["cost_total"] and d_ps.numpy()[k] == one["penalty"]
    assert abs(d_ct.numpy()[0] - sum(o["cost_total"] for o in singles)) <= 1e-13 * d_ct.numpy()[0]
    # ---- shortwave (per-band costs added to in scene order), scene costs only in scratch
    scenes, singles = [], []
    per_band = np.linspace(0.1, 0.5, nband)
    running = per_band.copy()
    for k, nprof in enumerate(nprofs):
        s = _scene(nprof, nlay, ng, nband, seed=60 + k, sw=True)
        rng = s["rng"]
        ssi, albedo, mu = rng.uniform(1.0, 100.0, ng), rng.uniform(0.05, 0.5, nband), rng.uniform(0.2, 1.0, nprof)
        dn, up, hr = rng.uniform(0, 300, (nprof, nlay + 1, nband)), rng.uniform(0, 100, (nprof, nlay + 1, nband)), \
            rng.uniform(0, 1e-4, (nprof, nlay, nband))
        sbw, sfd = rng.uniform(0, 1, ng), rng.uniform(0, 5, (nprof, ng))
        one = E.cost_function_ckd_sw(ctx, mu, s["p"], ssi, albedo, s["od"], dn, up, hr, *weights[:3],
                                     spectral_boundary_weights=sbw, spectral_flux_dn_surf=sfd, band_mapping=bm,
                                     cost_fn_per_band=running)
        running = one["cost_fn_per_band"]
        singles.append(one)
        scenes.append(dict(nprof=nprof, pressure_hl=ctx.to_device(s["p"]), optical_depth=ctx.to_device(s["od"]),
                           flux_dn=ctx.to_device(dn), flux_up=ctx.to_device(up), hr=ctx.to_device(hr),
                           spectral_flux_dn_surf=ctx.to_device(sfd), cos_sza=ctx.to_device(mu), ssi=ctx.to_device(ssi),
                           albedo=albedo, spectral_boundary_weights=ctx.to_device(sbw), cost_profile=ctx.empty(nprof)))
    d_pb = ctx.to_device(per_band)
    E.cost_function_ckd_scenes_dev(ctx, True, scenes, nlay, ng, nband, *weights[:3], d_band_mapping=d_bm,
                                   d_cost_fn_per_band=d_pb, d_cost_total=d_ct)
    for k, one in enumerate(singles):
        assert np.array_equal(scenes[k]["cost_profile"].numpy(), one["cost_profile"])
    np.testing.assert_allclose(d_pb.numpy(), running, rtol=1e-13)
    assert abs(d_ct.numpy()[0] - sum(o["cost_total"] for o in singles)) <= 1e-13 * d_ct.numpy()[0]
    with pytest.raises(E.EcckdError) as ei:
        E.cost_function_ckd_scenes_dev(ctx, False, [dict(nprof=2)], nlay, ng, nband, *weights)
    assert ei.value.code == 147


def test_cost_function_errors(ctx):
    s = _scene(2, 6, 8, 3, seed=1)
    z = np.zeros((2, 7, 3))
    with pytest.raises(E.EcckdError) as ei:  # truth arrays per band but no band_mapping
        E.cost_function_ckd_lw(ctx, s["p"], s["planck"], np.ones((2, 3)), s["surf"], s["od"], z, z, np.zeros((2, 6, 3)),
                               0.2, 0.2, 0.8, 0.1)
    assert ei.value.code == 147


def test_split_base_g_points_bit_exact(ctx, O):
    c = S.sw_spectrum_case(nlay=4, nwav=20000, ng=9, nband=3, seed=6)
    wn, g, ng = c["wavenumber"], c["g_point"], c["ng"]
    edges = c["band_edges"]
    b1, b2 = edges[:-1].copy(), edges[1:].copy()
    band_number = np.repeat(np.arange(3), S.split_g_points(9, 3)).astype(np.int32)
    solar = np.array([c["ssi"][g == k].sum() for k in range(ng)])
    # two cuts in band 0, none in band 1, one in band 2, plus values outside every band / on an edge
    bounds = np.array([edges[0] + 0.3 * (edges[1] - edges[0]), edges[0] + 0.6 * (edges[1] - edges[0]),
                       0.5 * (edges[2] + edges[3]), edges[1], 1.0e9])
    for use_ssi in (True, False):
        ref = O.split_base_g_points(b1, b2, bounds, wn, c["ssi"] if use_ssi else None, ng, band_number, g,
                                    solar if use_ssi else None)
        d_g = ctx.to_device(g, np.int32)
        got = E.split_base_g_points_dev(ctx, b1, b2, bounds, wn.size, ctx.to_device(wn),
                                        ctx.to_device(c["ssi"]) if use_ssi else None, ng, band_number, d_g,
                                        solar if use_ssi else None)
        assert got[0] == ref[0] == ng + 3
        assert np.array_equal(got[1], ref[1]) and np.array_equal(d_g.numpy(), ref[2])
        if use_ssi:
            np.testing.assert_allclose(got[2], ref[3], rtol=1e-13)
    with pytest.raises(E.EcckdError) as ei:
        E.split_base_g_points_dev(ctx, b1, b2, bounds, wn.size, ctx.to_device(wn), None, ng, band_number,
                                  ctx.to_device(g, np.int32), None, cap=ng + 1)
    assert ei.value.code == 148


def test_rayleigh_coefficients(ctx, O):
    wn = np.concatenate([np.linspace(275.0, 49975.0, 995), [19999.9, 20000.0, 20000.1]])
    np.testing.assert_allclose(E.rayleigh_molar_scattering_coeff(ctx, wn), O.rayleigh_molar_scattering_coeff(wn),
                               rtol=1e-13)
    rng = np.random.default_rng(3)
    w1 = 250.0 + 50.0 * np.arange(995)
    w2 = w1 + 50.0
    gf = rng.uniform(0, 1, (32, 995)) * (rng.uniform(size=(32, 995)) < 0.1)
    gf[:, 0] += 1e-3
    ssi_int = S.solar_spectrum(0.5 * (w1 + w2))
    np.testing.assert_allclose(E.rayleigh_molar_scat_gpoints(ctx, w1, w2, gf, ssi_int),
                               O.rayleigh_molar_scat_gpoints(w1, w2, gf, ssi_int), rtol=1e-12)


@pytest.mark.parametrize("banded", [True, False])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8, 0.1), (0.5, 0.0, 0.0, 0.0)])
def test_cost_function_ckd_lw_gradient_matches_finite_differences(ctx, O, banded, weights):
    """the hand-written reverse mode of the longwave cost (calc_cost_function_lw.cpp:115-232 +
    the penalty of solve_adept.cpp:107-116) against central differences of the ORACLE's forward cost (the function
    Adept would differentiate, pinned to the reference's own translation unit) and of the device's forward cost:
    random directions for the whole gradient, single entries for a sample of (profile, layer, g point)."""
    nprof, nlay, ng, nb = 3, 20, 16, 5
    s = _scene(nprof, nlay, ng, nb, seed=9)
    rng = s["rng"]
    nband = nb if banded else ng
    bm = s["bm"] if banded else None
    emis = rng.uniform(0.9, 1.0, (nprof, nband))
    od = np.abs(s["od"]) * rng.choice([1e-4, 1e-2, 1.0, 5.0], size=s["od"].shape)  # thin (f = e/2 branch) to thick layers
    od[0, 3, 2] = -0.01   # clamped entries: only the penalty depends on them
    od[2, 7, 5] = -0.003
    t_dn, t_up, hr, lws = [], [], [], []
    for i in range(nprof):
        eg = emis[i][bm] if banded else emis[i]
        dn, up = O.radiative_transfer_lw(s["planck"][i], np.maximum(od[i], 0) * 1.3, eg, s["surf"][i])
        if banded:
            dn = np.stack([dn[:, bm == b].sum(1) for b in range(nband)], 1)
            up = np.stack([up[:, bm == b].sum(1) for b in range(nband)], 1)
        t_dn.append(dn), t_up.append(up), hr.append(O.heating_rate(s["p"][i], dn, up))
        lw = np.diff(np.sqrt(s["p"][i]))
        lws.append(lw / lw.sum())
    t_dn, t_up, hr, lws = np.stack(t_dn), np.stack(t_up), np.stack(hr), np.stack(lws)
    sfd, sfu = rng.uniform(0, 5, (nprof, ng)), rng.uniform(0, 5, (nprof, ng))
    rel = rng.uniform(0, 1, (2, nprof, nlay + 1, ng))
    pen = 50.0
    kw = dict(layer_weight=lws, negative_od_penalty=pen, spectral_flux_dn_surf=sfd, spectral_flux_up_toa=sfu,
              rel_dn=rel[0], rel_up=rel[1], band_mapping=bm)

    def cost_oracle(x):
        c = pen * float((np.minimum(x, 0.0) ** 2).sum())
        for i in range(nprof):
            c += O.calc_cost_function_ckd_lw(s["p"][i], s["planck"][i], emis[i], s["surf"][i], np.maximum(x[i], 0.0),
                                             t_dn[i], t_up[i], hr[i], *weights, lws[i], spectral_flux_dn_surf=sfd[i],
                                             spectral_flux_up_toa=sfu[i], rel_dn=rel[0][i], rel_up=rel[1][i],
                                             band_mapping=bm)
        return c

    def cost_device(x):
        return E.cost_function_ckd_lw(ctx, s["p"], s["planck"], emis, s["surf"], x, t_dn, t_up, hr, *weights, **kw)["cost_total"]

    grad = E.cost_function_ckd_lw_grad(ctx, s["p"], s["planck"], emis, s["surf"], od, t_dn, t_up, hr, *weights, **kw)
    assert grad.shape == od.shape and np.all(np.isfinite(grad))
    assert grad[0, 3, 2] == 2 * pen * od[0, 3, 2] and grad[2, 7, 5] == 2 * pen * od[2, 7, 5]
    # directional derivatives (relative steps: the entries span four decades); clamped entries are kept out of the
    # direction, a central difference across the kink at zero is not a derivative
    # ... and so are the entries next to the reference's own branch point (emissivity = 1e-5, i.e. tau = 6.0e-6,
    # radiative_transfer_lw.cpp:41-43), where its two expressions for the Planck-gradient factor differ by ~1e-11:
    # a jump of that size over a step of 1e-6 tau is not a derivative either
    free = (od > 0.0) & ~((od > 2.0e-6) & (od < 2.0e-5))
    for k in range(4):
        v = rng.normal(size=od.shape) * od * free
        h = 1.0e-4  # relative step (v is proportional to od): truncation ~h^2, rounding ~1e-14 cost / h
        for cost in (cost_oracle, cost_device):
            fd = (cost(od + h * v) - cost(od - h * v)) / (2 * h)
            an = float((grad * v).sum())
            assert abs(fd - an) <= 5.0e-6 * max(abs(an), abs(fd)) + 2e-9 * abs(cost(od)), (k, cost.__name__, fd, an)
    # single entries (those big enough for a central difference of a cost of O(100) to resolve the step)
    for _ in range(40):
        i, l, g = rng.integers(nprof), rng.integers(nlay), rng.integers(ng)
        if not free[i, l, g] or od[i, l, g] < 1.0e-3:
            continue
        h = 1.0e-4 * od[i, l, g]
        xp, xm = od.copy(), od.copy()
        xp[i, l, g] += h
        xm[i, l, g] -= h
        fd = (cost_oracle(xp) - cost_oracle(xm)) / (2 * h)
        scale = np.abs(grad[i]).max()
        assert abs(fd - grad[i, l, g]) <= 1.0e-5 * abs(grad[i, l, g]) + 1e-7 * scale, (i, l, g, fd, grad[i, l, g])


def test_gradient_with_respect_to_the_look_up_tables_end_to_end(ctx, O):
    """SURVEY.md 8f item 2, the chain of optimize_lut for a longwave scene: tables -> calc_total_optical_depth ->
    cost, and back: ecckd_cost_function_ckd_lw_grad_dev then ecckd_calc_total_optical_depth_adjoint_dev.  Central
    differences of the device's forward chain in single table entries (models re-created with the entry moved)."""
    from tests import synthetic as S
    ncol, nlay, ng = 4, 20, 16
    c = S.ckd_model_case(ncol=ncol, nlay=nlay, ng=ng, seed=41)
    rng = np.random.default_rng(8)
    d_vmr = [None if v is None else ctx.to_device(v) for v in c["vmr_fl"]]
    d_p, d_t = ctx.to_device(c["pressure_hl"]), ctx.to_device(c["temperature_hl"])
    planck = rng.uniform(5.0, 30.0, (ncol, nlay + 1, ng)) * np.linspace(0.5, 1.5, nlay + 1)[None, :, None]
    emis, surf = rng.uniform(0.9, 1.0, (ncol, ng)), planck[:, -1, :] * 1.02
    weights = (0.2, 0.2, 0.8, 0.0)

    def model_of(gases):
        return E.CkdModel(ctx, gases, c["log_pressure"], c["temperature"], c["temperature_planck"], c["planck_lut"])

    def od_of(m):
        d_od = ctx.empty((ncol, nlay, ng))
        m.calc_total_optical_depth_dev(d_vmr, ncol, nlay, d_p, d_t, d_od)
        return d_od.numpy()

    m0 = model_of(c["gases"])
    od0 = od_of(m0)
    t_dn, t_up, hr = [], [], []
    for i in range(ncol):  # "truth": the fluxes of somewhat different optical depths
        dn, up = O.radiative_transfer_lw(planck[i], od0[i] * 1.2, emis[i], surf[i])
        t_dn.append(dn), t_up.append(up), hr.append(O.heating_rate(c["pressure_hl"][i], dn, up))
    t_dn, t_up, hr = np.stack(t_dn), np.stack(t_up), np.stack(hr)

    def cost_of(od):
        return E.cost_function_ckd_lw(ctx, c["pressure_hl"], planck, emis, surf, od, t_dn, t_up, hr, *weights)["cost_total"]

    g_od = E.cost_function_ckd_lw_grad(ctx, c["pressure_hl"], planck, emis, surf, od0, t_dn, t_up, hr, *weights)
    grads = [ctx.zeros(np.asarray(g["molar_abs"]).shape) for g in c["gases"]]
    m0.calc_total_optical_depth_adjoint_dev(d_vmr, ncol, nlay, d_p, d_t, ctx.to_device(g_od), grads)
    grads = [g.numpy() for g in grads]
    checked = 0
    for igas in (1, 2, 4):  # h2o (concentration axis), o3 (linear), ch4 (relative-linear)
        gk = grads[igas]
        for idx in np.argsort(np.abs(gk).ravel())[::-1][:2]:  # the entries that matter most
            tab = np.asarray(c["gases"][igas]["molar_abs"])
            h = 1.0e-4 * tab.ravel()[idx]
            costs = []
            for sgn in (+1, -1):
                t2 = tab.copy()
                t2.ravel()[idx] += sgn * h
                gases = [dict(g) for g in c["gases"]]
                gases[igas]["molar_abs"] = t2
                m = model_of(gases)
                costs.append(cost_of(od_of(m)))
                m.close()
            fd = (costs[0] - costs[1]) / (2 * h)
            an = gk.ravel()[idx]
            assert abs(fd - an) <= 1.0e-5 * abs(an) + 1e-9 * abs(costs[0]) / abs(h), (igas, idx, fd, an)
            checked += 1
    assert checked == 6


@pytest.mark.parametrize("albedo_on", [True, False])
@pytest.mark.parametrize("weights", [(0.2, 0.2, 0.8), (0.5, 0.0, 0.0)])
def test_cost_function_ckd_sw_gradient_matches_finite_differences(ctx, O, albedo_on, weights):
    """The hand-written reverse mode of the shortwave cost (calc_cost_function_sw.cpp:115-277 + the penalty of
    solve_adept.cpp:107-116) against central differences of the pinned oracle's forward cost and of the device's:
    with and without the Lambertian return, with and without the broadband part (no division by nband then)."""
    nprof, nlay, ng, nband = 3, 18, 16, 5
    s = _scene(nprof, nlay, ng, nband, seed=15, sw=True)
    rng, bm = s["rng"], s["bm"]
    ssi = rng.uniform(1.0, 100.0, ng)
    albedo = rng.uniform(0.05, 0.5, nband) if albedo_on else np.zeros(nband)
    mu = rng.uniform(0.2, 1.0, nprof)
    od = np.abs(s["od"]) * rng.choice([1e-3, 1e-1, 1.0], size=s["od"].shape)
    od[1, 4, 3] = -0.02  # clamped: only the penalty depends on it
    t_dn, t_up, hr, lws = [], [], [], []
    for i in range(nprof):
        dn, up = O.radiative_transfer_norayleigh_sw(mu[i], ssi, np.maximum(od[i], 0) * 1.2, albedo[bm])
        dn = np.stack([dn[:, bm == b].sum(1) for b in range(nband)], 1)
        up = np.stack([up[:, bm == b].sum(1) for b in range(nband)], 1)
        t_dn.append(dn), t_up.append(up), hr.append(O.heating_rate(s["p"][i], dn, None))
        lw = np.diff(np.sqrt(s["p"][i]))
        lws.append(lw / lw.sum())
    t_dn, t_up, hr, lws = np.stack(t_dn), np.stack(t_up), np.stack(hr), np.stack(lws)
    sbw, sfd = rng.uniform(0, 1, ng), rng.uniform(0, 5, (nprof, ng))
    pen = 30.0
    kw = dict(layer_weight=lws, negative_od_penalty=pen, spectral_boundary_weights=sbw, spectral_flux_dn_surf=sfd,
              band_mapping=bm)

    def cost_oracle(x):
        c = pen * float((np.minimum(x, 0.0) ** 2).sum())
        for i in range(nprof):
            c += O.calc_cost_function_ckd_sw(mu[i], s["p"][i], ssi, albedo, np.maximum(x[i], 0.0), t_dn[i], t_up[i], hr[i],
                                             *weights, lws[i], spectral_boundary_weights=sbw, spectral_flux_dn_surf=sfd[i],
                                             spectral_flux_up_toa=sfd[i], band_mapping=bm)
        return c

    def cost_device(x):
        return E.cost_function_ckd_sw(ctx, mu, s["p"], ssi, albedo, x, t_dn, t_up, hr, *weights, **kw)["cost_total"]

    grad = E.cost_function_ckd_sw_grad(ctx, mu, s["p"], ssi, albedo, od, t_dn, t_up, hr, *weights, **kw)
    assert grad.shape == od.shape and np.all(np.isfinite(grad))
    assert grad[1, 4, 3] == 2 * pen * od[1, 4, 3]
    free = od > 0.0
    for k in range(4):
        v = rng.normal(size=od.shape) * od * free
        h = 1.0e-4
        for cost in (cost_oracle, cost_device):
            fd = (cost(od + h * v) - cost(od - h * v)) / (2 * h)
            an = float((grad * v).sum())
            assert abs(fd - an) <= 5.0e-6 * max(abs(an), abs(fd)) + 2e-9 * abs(cost(od)), (k, cost.__name__, fd, an)


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ng", [(1, 1), (7, 5), (318, 32), (636, 200)])
def test_background_cost_matches_the_oracle(ctx, O, nx, ng):
    """CkdModel<true>::calc_background_cost_function (ckd_model.cpp:838-873) for one gas and for the Rayleigh part:
    cost and accumulated gradient against the oracle, the same bits on every call."""
    rng = np.random.default_rng(nx * 1000 + ng)
    a = rng.standard_normal((nx, nx))
    shape = a @ a.T / nx + np.eye(nx)
    err = rng.uniform(0.5, 4.0, ng)
    dx = rng.standard_normal((nx, ng))
    grad0 = rng.standard_normal((nx, ng))
    ref_grad = grad0.copy()
    ref_cost = O.background_cost_gas(shape, err, dx, ref_grad)
    d_shape, d_err, d_dx = ctx.to_device(shape), ctx.to_device(err), ctx.to_device(dx)
    d_grad = ctx.to_device(grad0)
    cost = E.background_cost_gas_dev(ctx, nx, ng, d_shape, d_err, d_dx, d_grad)
    np.testing.assert_allclose(cost, ref_cost, rtol=1e-13)
    scale = np.abs(ref_grad - grad0).max()
    np.testing.assert_allclose(d_grad.numpy(), ref_grad, rtol=1e-12, atol=1e-13 * scale)
    d_grad2 = ctx.to_device(grad0)
    cost2 = E.background_cost_gas_dev(ctx, nx, ng, d_shape, d_err, d_dx, d_grad2)
    assert cost2 == cost and np.array_equal(d_grad2.numpy(), d_grad.numpy())
    assert E.background_cost_gas_dev(ctx, nx, ng, d_shape, d_err, d_dx, None) == cost  # cost only
    inv_b = 1.0 / err ** 2
    d_inv = ctx.to_device(inv_b)
    got = E.background_cost_rayleigh_dev(ctx, ng, d_inv, d_dx)  # the first row of dx as the Rayleigh slice
    np.testing.assert_allclose(got, O.background_cost_rayleigh(inv_b, dx[0]), rtol=1e-14)
    with pytest.raises(E.EcckdError):
        E.background_cost_gas_dev(ctx, 0, ng, d_shape, d_err, d_dx, None)
