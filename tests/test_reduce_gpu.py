"""GPU parity of the spectral reduction (average_optical_depth.cpp:22-197 and
companions) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S
from tests.oracle_lib import METHODS

pytestmark = pytest.mark.gpu


def _planck(O, c):
    return O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])


@pytest.mark.parametrize("method", METHODS)
def test_avg_host_api_all_methods(ctx, O, method):
    c = S.lw_spectrum_case(nlay=13, nwav=6000, ng=10, nband=3, seed=21)
    planck = _planck(O, c)
    got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                             c["od"], planck, method)
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                             c["od"], planck, method)
    S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=method + " ")


@pytest.mark.parametrize("nwav", [1, 2, 63, 64, 65, 127, 4097])
def test_avg_ragged_sizes(ctx, O, nwav):
    """Chunk boundaries, odd lengths (8-byte copy path) and tiny spectra."""
    c = S.lw_spectrum_case(nlay=5, nwav=max(nwav, 8), ng=3, nband=1, seed=nwav)
    sl = slice(0, nwav)
    od, g = np.ascontiguousarray(c["od"][:, sl]), c["g_point"][sl]
    planck = np.ascontiguousarray(_planck(O, c)[:, sl])
    for method in ("transmission", "logarithmic"):
        got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], g, od,
                                                 planck, method)
        ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], g, od, planck,
                                                 method)
        S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=f"nwav={nwav} {method} ")


def test_avg_empty_unassigned_and_all_zero(ctx, O):
    c = S.lw_spectrum_case(nlay=7, nwav=3000, ng=6, nband=2, seed=4)
    g = c["g_point"].copy()
    g[g == 3] = 2       # empty g point
    g[::13] = -1        # wavenumbers in no g point
    g[5::97] = 99       # out-of-range index also belongs to no g point
    od = c["od"].copy()
    od[2, g == 1] = 0.0  # a (layer, g) with no positive optical depth at all
    planck = _planck(O, c)
    for method in METHODS:
        got = E.average_optical_depth_to_g_point(ctx, 6, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, od, planck,
                                                 method)
        ref = O.average_optical_depth_to_g_point(6, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, od, planck,
                                                 method)
        S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=method + " ")
        assert got[3][3] & 1 and np.all(got[0][:, 3] == 0.0)


def test_avg_raw_optical_depth_and_solar_weight(ctx, O):
    """reference_surface_vmr <= 0 returns optical depth (:186-193); SW weights are one
    row broadcast over layers (create_look_up_table.cpp:335)."""
    c = S.sw_spectrum_case(nlay=9, nwav=5000, ng=8, nband=2, seed=6)
    for vmr in (-1.0, 2.0e-3):
        got = E.average_optical_depth_to_g_point(ctx, c["ng"], vmr, c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                                 c["od"], c["ssi"], "transmission")
        ref = O.average_optical_depth_to_g_point(c["ng"], vmr, c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                                 c["od"], c["ssi"], "transmission")
        cc = dict(c, vmr=vmr)
        S.assert_avg_close(got, ref, method="transmission", scale=S.out_scale(cc))


def test_avg_bad_method_is_parameter_error(ctx):
    c = S.lw_spectrum_case(nlay=3, nwav=100, ng=2, nband=1, seed=5)
    with pytest.raises(E.EcckdError) as ei:
        E.average_optical_depth_to_g_point(ctx, 2, 1.0, c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"],
                                           c["od"], "total-transmission")
    assert ei.value.code == 147


def test_avg_known_answer_constant_od(ctx):
    c = S.lw_spectrum_case(nlay=6, nwav=1200, ng=5, nband=1, seed=2, zero_frac=0.0)
    tau_g = np.array([1e-6, 3e-3, 0.2, 1.5, 4.0])
    lay = np.linspace(0.5, 2.0, 6)
    od = np.ascontiguousarray(np.broadcast_to(tau_g[c["g_point"]], c["od"].shape) * lay[:, None])
    w = np.abs(np.sin(np.arange(1200.0)))[None, :] + 0.1 * np.arange(1, 7)[:, None]
    for m in METHODS:
        out, mn, mx, warn = E.average_optical_depth_to_g_point(ctx, 5, -1.0, c["pressure_fl"], c["pressure_hl"],
                                                               c["g_point"], od, w, m)
        np.testing.assert_allclose(out, lay[:, None] * tau_g[None, :], rtol=1e-9, err_msg=m)
        assert np.all(warn & 4)


def _device_case(ctx, c):
    d = dict(od=ctx.to_device(c["od"]), g=ctx.to_device(c["g_point"], np.int32), wn=ctx.to_device(c["wavenumber"]),
             dwn=ctx.to_device(c["d_wavenumber"]))
    d["plan"] = E.GPointPlan(ctx, d["g"], c["nwav"], c["ng"])
    return d


def test_plan_counts(ctx):
    c = S.lw_spectrum_case(nlay=2, nwav=7001, ng=9, nband=3, seed=8)
    d = _device_case(ctx, c)
    assert np.array_equal(d["plan"].counts(), np.bincount(c["g_point"], minlength=c["ng"]))


@pytest.mark.parametrize("method", ["transmission", "hybrid-logarithmic-transmission-3", "linear"])
def test_avg_fused_planck_weights(ctx, O, method):
    """ECCKD_WEIGHT_PLANCK fuses planck_function (planck_function.cpp:47-53) into the
    reduction; result must equal planck_function followed by the averaging."""
    c = S.lw_spectrum_case(nlay=40, nwav=9000, ng=12, nband=3, seed=31)
    d = _device_case(ctx, c)
    nlay, ng = c["nlay"], c["ng"]
    out, omin, omax = ctx.empty((nlay, ng)), ctx.empty((nlay, ng)), ctx.empty((nlay, ng))
    warn = ctx.zeros(ng, np.int32)
    E.average_optical_depth_to_g_point_dev(ctx, d["plan"], c["vmr"], c["pressure_fl"], c["pressure_hl"], d["od"],
                                           c["nwav"], E.WEIGHT_PLANCK, None, 0, c["temperature_fl"], d["wn"], d["dwn"],
                                           method, out, omin, omax, ng, warn)
    ref = O.average_optical_depth_to_g_point(ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"],
                                             _planck(O, c), method)
    S.assert_avg_close((out.numpy(), omin.numpy(), omax.numpy(), warn.numpy()), ref, method=method,
                       scale=S.out_scale(c))


@pytest.mark.parametrize("weight_mode", ["planck", "matrix", "vector"])
def test_avg_batch_matches_single_calls(ctx, O, weight_mode):
    """All (gas, concentration) blocks of one temperature index in one pass: identical
    to nspec separate reference calls."""
    nspec = 5
    base = S.lw_spectrum_case(nlay=11, nwav=7000, ng=9, nband=2, seed=40)
    rng = np.random.default_rng(1)
    ods = [base["od"] * np.exp(rng.normal(0, 1.0, base["od"].shape)) for _ in range(nspec)]
    vmrs = [1.0, 3e-4, 1e-6, -1.0, 2.5e-2]
    d = _device_case(ctx, base)
    nlay, ng, nwav = base["nlay"], base["ng"], base["nwav"]
    planck = _planck(O, base)
    ssi = np.abs(np.sin(np.arange(nwav) * 0.01)) + 0.2
    if weight_mode == "planck":
        mode, d_w, wstride, w_ref = E.WEIGHT_PLANCK, None, 0, planck
    elif weight_mode == "matrix":
        mode, d_w, wstride, w_ref = E.WEIGHT_MATRIX, ctx.to_device(planck), nwav, planck
    else:
        mode, d_w, wstride, w_ref = E.WEIGHT_VECTOR, ctx.to_device(ssi), 0, ssi
    d_ods = [ctx.to_device(o) for o in ods]
    outs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    mins = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    maxs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    warns = [ctx.zeros(ng, np.int32) for _ in range(nspec)]
    E.average_optical_depth_batch_dev(ctx, d["plan"], vmrs, base["pressure_fl"], base["pressure_hl"], d_ods, nwav,
                                      mode, d_w, wstride, base["temperature_fl"], d["wn"], d["dwn"], "transmission",
                                      outs, mins, maxs, ng, warns)
    for s in range(nspec):
        ref = O.average_optical_depth_to_g_point(ng, vmrs[s], base["pressure_fl"], base["pressure_hl"],
                                                 base["g_point"], ods[s], w_ref, "transmission")
        cc = dict(base, vmr=vmrs[s])
        S.assert_avg_close((outs[s].numpy(), mins[s].numpy(), maxs[s].numpy(), warns[s].numpy()), ref,
                           method="transmission", scale=S.out_scale(cc), what=f"spec {s} ")


def test_avg_incoherent_g_map(ctx, O):
    """A g map with no spectral coherence (every chunk holds most g points)."""
    c = S.lw_spectrum_case(nlay=8, nwav=8000, ng=40, nband=1, seed=12, coherent=False)
    planck = _planck(O, c)
    got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                             c["od"], planck, "transmission")
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                             c["od"], planck, "transmission")
    S.assert_avg_close(got, ref, method="transmission", scale=S.out_scale(c))


def test_avg_linearity_property_full_width(ctx):
    """Size-independent property at a larger size than the oracle is run at: the
    'linear' average of a*od1 + b*od2 equals a*avg(od1) + b*avg(od2) (raw optical depth),
    and the result does not depend on how the kernel splits the spectrum (determinism)."""
    nlay, nwav, ng = 53, 1 << 18, 64
    rng = np.random.default_rng(0)
    wn = np.linspace(0.0002, 3260.0, nwav)
    key = np.log(S.smooth_spectrum(wn, rng, ncomp=40))
    g, ng = S.g_point_map(wn, key, np.linspace(wn[0], wn[-1], 14), S.split_g_points(ng, 13), rng)
    od1 = np.exp(rng.normal(-2, 1.5, (nlay, nwav)))
    od2 = np.exp(rng.normal(-1, 1.0, (nlay, nwav)))
    w = rng.uniform(0.5, 1.5, (nlay, nwav))
    p_hl = S.pressure_grid(nlay)
    p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
    run = lambda od: E.average_optical_depth_to_g_point(ctx, ng, -1.0, p_fl, p_hl, g, od, w, "linear")
    a1, a2 = run(od1), run(od2)
    a3 = run(2.0 * od1 + 0.5 * od2)
    np.testing.assert_allclose(a3[0], 2.0 * a1[0] + 0.5 * a2[0], rtol=1e-12)
    od3 = 2.0 * od1 + 0.5 * od2
    mn_ref = np.full((nlay, ng), np.inf)
    for l in range(nlay):
        np.minimum.at(mn_ref[l], g, od3[l])
    assert np.array_equal(a3[1], mn_ref)  # minima are selections: exact
    again = run(od1)
    assert np.array_equal(again[0], a1[0])  # bitwise reproducible
    # weighted mean from NumPy directly
    sums = np.zeros((nlay, ng))
    wsum = np.zeros((nlay, ng))
    for l in range(nlay):
        sums[l] = np.bincount(g, weights=od1[l] * w[l], minlength=ng)
        wsum[l] = np.bincount(g, weights=w[l], minlength=ng)
    np.testing.assert_allclose(a1[0], sums / wsum, rtol=1e-12)


@pytest.mark.parametrize("method", ["transmission", "logarithmic"])
def test_avg_wavenumber_shards_emulated_on_one_device(ctx, O, method):
    """The multi-GPU decomposition (partial per wavenumber shard -> SUM/MIN reduction of the
    partial accumulators -> finalize) with three shards emulated on one device; the reduction
    across shards is done here in NumPy exactly as NCCL would (sum, min)."""
    from ecckd_b200 import sharding

    c = S.lw_spectrum_case(nlay=9, nwav=6007, ng=11, nband=2, seed=77)
    nlay, ng, nwav, nspec = c["nlay"], c["ng"], c["nwav"], 2
    ods = [c["od"], c["od"][::-1].copy() * 0.5]
    vmrs = [c["vmr"], -1.0]
    sums_tot, mm_tot, cnt_tot = None, None, None
    for r in range(3):
        lo, hi = sharding.shard_range(nwav, r, 3)
        d_g = ctx.to_device(c["g_point"][lo:hi], np.int32)
        plan = E.GPointPlan(ctx, d_g, hi - lo, ng)
        d_ods = [ctx.to_device(np.ascontiguousarray(o[:, lo:hi])) for o in ods]
        d_wn, d_dwn = ctx.to_device(c["wavenumber"][lo:hi]), ctx.to_device(c["d_wavenumber"][lo:hi])
        d_s, d_m = ctx.empty((nlay, nspec, ng, 3)), ctx.empty((nlay, nspec, ng, 2))
        E.average_optical_depth_partial_dev(ctx, plan, nlay, c["pressure_fl"], d_ods, hi - lo, E.WEIGHT_PLANCK, None, 0,
                                            c["temperature_fl"], d_wn, d_dwn, method, d_s, d_m)
        s_r, m_r, c_r = d_s.numpy(), d_m.numpy(), plan.counts()
        sums_tot = s_r if sums_tot is None else sums_tot + s_r
        mm_tot = m_r if mm_tot is None else np.minimum(mm_tot, m_r)
        cnt_tot = c_r if cnt_tot is None else cnt_tot + c_r
    assert np.array_equal(cnt_tot, np.bincount(c["g_point"], minlength=ng))
    d_s, d_m, d_c = ctx.to_device(sums_tot), ctx.to_device(mm_tot), ctx.to_device(cnt_tot, np.int64)
    outs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    mins = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    maxs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    warns = [ctx.zeros(ng, np.int32) for _ in range(nspec)]
    E.average_optical_depth_finalize_dev(ctx, ng, d_c, vmrs, c["pressure_fl"], c["pressure_hl"], method, d_s, d_m, outs,
                                         mins, maxs, ng, warns)
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    for s in range(nspec):
        ref = O.average_optical_depth_to_g_point(ng, vmrs[s], c["pressure_fl"], c["pressure_hl"], c["g_point"], ods[s],
                                                 planck, method)
        S.assert_avg_close((outs[s].numpy(), mins[s].numpy(), maxs[s].numpy(), warns[s].numpy()), ref, method=method,
                           scale=S.out_scale(dict(c, vmr=vmrs[s])), what=f"spec {s} ")


def test_avg_many_g_points_per_chunk(ctx, O):
    """A chunk of 480 wavenumbers holding up to 150 g points needs several plan rounds (more lanes than a
    warp); the shorter chunk lengths of the plan builder are exercised too."""
    rng = np.random.default_rng(5)
    c = S.lw_spectrum_case(nlay=6, nwav=9000, ng=8, nband=1, seed=31)
    ng = 150
    g = rng.integers(0, ng, size=c["g_point"].size).astype(np.int32)
    g[:1500] = np.repeat(np.arange(ng, dtype=np.int32), 10)  # long runs too
    planck = _planck(O, c)
    for method in ("transmission", "linear", "logarithmic"):
        got = E.average_optical_depth_to_g_point(ctx, ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck,
                                                 method)
        ref = O.average_optical_depth_to_g_point(ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck,
                                                 method)
        S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=method + " ")


def test_avg_extremes_are_exact_with_duplicates_huge_and_tiny_values(ctx, O):
    """min/max are selections of input values: the range-check fast path must never lose an extreme that
    shares its high word with the running one, repeats it, is zero, denormal, or enormous (optical depths
    beyond 708/kd saturate 1 - exp exactly)."""
    rng = np.random.default_rng(17)
    c = S.lw_spectrum_case(nlay=9, nwav=20000, ng=6, nband=2, seed=8)
    od = c["od"].copy()
    od[0, :] = 0.5                                  # constant row: min == max
    od[1, :] = 0.5 * (1.0 + 1e-9 * rng.random(od.shape[1]))  # all values share the extremes' high word
    od[2, ::7] = 0.0                                # many exact zeros
    od[3, 11::501] = 5e-324                         # denormal minimum
    od[4, 3::997] = 1.0e5                           # beyond the exponent clamp
    od[5, 123] = 3.0e12
    od[6, :] = np.sort(od[6, :])                    # new maximum in every chunk
    od[7, :] = np.sort(od[7, :])[::-1]              # new minimum in every chunk
    planck = _planck(O, c)
    for method in ("transmission", "linear"):
        got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                                 od, planck, method)
        ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                                 planck, method)
        S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=method + " ")


def test_avg_negative_optical_depths_take_the_exact_path(ctx, O):
    """Unphysical but legal input: the extremes of a g point holding negative values are still exact."""
    c = S.lw_spectrum_case(nlay=4, nwav=5000, ng=4, nband=1, seed=12)
    od = c["od"].copy()
    od[1, ::5] *= -1.0
    od[2, :] = -np.abs(od[2, :])
    planck = _planck(O, c)
    got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                             planck, "linear")
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                             planck, "linear")
    for i in (1, 2):  # min / max before the reference's clamping logic touches them: compare where both finite
        ok = np.isfinite(ref[i]) & np.isfinite(got[i])
        assert np.allclose(got[i][ok], ref[i][ok], rtol=1e-15, atol=0.0)


def test_avg_group_of_twenty_spectra_and_wavenumber_windows(ctx, O, monkeypatch):
    """Twenty spectra share each layer's weights (one full CTA group: two rows per consumer warp, the last
    warp also sums the weights) and the work is cut into several wavenumber windows."""
    monkeypatch.setenv("ECCKD_NWIN", "5")
    nspec, nlay, nwav, ng = 20, 3, 30000, 12
    c = S.lw_spectrum_case(nlay=nlay, nwav=nwav, ng=ng, nband=3, seed=44)
    rng = np.random.default_rng(3)
    ods = [np.ascontiguousarray(c["od"] * rng.uniform(0.2, 5.0)) for _ in range(nspec)]
    vmrs = [float(v) for v in rng.uniform(1e-6, 1e-3, nspec)]
    d_g = ctx.to_device(c["g_point"].astype(np.int32))
    plan = E.GPointPlan(ctx, d_g, nwav, ng)
    d_wn, d_dwn = ctx.to_device(c["wavenumber"]), ctx.to_device(c["d_wavenumber"])
    d_ods = [ctx.to_device(o) for o in ods]
    outs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    mins = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    maxs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
    E.average_optical_depth_batch_dev(ctx, plan, vmrs, c["pressure_fl"], c["pressure_hl"], d_ods, nwav, E.WEIGHT_PLANCK, None,
                                      nwav, c["temperature_fl"], d_wn, d_dwn, "transmission", outs, mins, maxs, ng)
    planck = _planck(O, c)
    for s in range(nspec):
        ref = O.average_optical_depth_to_g_point(ng, vmrs[s], c["pressure_fl"], c["pressure_hl"], c["g_point"], ods[s], planck,
                                                 "transmission")
        got = (outs[s].numpy(), mins[s].numpy(), maxs[s].numpy(), None)
        scale = (9.80665 * 0.001 * 28.970 / vmrs[s]) / (c["pressure_hl"][1:] - c["pressure_hl"][:-1])
        S.assert_avg_close(got, ref, method="transmission", scale=scale, what=f"spectrum {s} ")


@pytest.mark.parametrize("method", ["transmission", "logarithmic"])
def test_avg_host_call_float32_spectra_bit_identical(ctx, method):
    """ecckd_planck_average_optical_depth_to_g_point_f32: float32 optical depths (what the CKDMIP files hold)
    widened on the device give bit-identical results to widening on the host; odd nwav and a zero row."""
    c = S.lw_spectrum_case(nlay=23, nwav=9001, ng=10, nband=2, seed=77)
    od32 = c["od"].astype(np.float32)
    od32[5] = 0.0
    args = (c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"])
    tail = (c["temperature_fl"], c["wavenumber"], c["d_wavenumber"], method)
    a = E.planck_average_optical_depth_to_g_point(ctx, *args, od32, *tail)
    b = E.planck_average_optical_depth_to_g_point(ctx, *args, od32.astype(np.float64), *tail)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("method", ["transmission", "transmission-10"])
def test_avg_huge_optical_depths_saturate(ctx, O, method):
    """Optical depths beyond 708/kd (where exp(-kd tau) underflows) scattered through every g point, as the first
    values of a g point and as whole rows: the kernel's unclamped fast path must never see them."""
    c = S.lw_spectrum_case(nlay=9, nwav=6000, ng=8, nband=2, seed=91, zero_frac=0.0)
    od = c["od"].copy()
    rng = np.random.default_rng(5)
    od[rng.uniform(size=od.shape) < 0.03] = 1e3
    od[rng.uniform(size=od.shape) < 0.01] = 1e30
    od[2, :] = 5e4          # a whole layer opaque
    od[3, :40] = 1e300      # the first values a lane sees
    planck = _planck(O, c)
    got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                             planck, method)
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                             planck, method)
    assert np.all(np.isfinite(got[0]))
    S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c))


@pytest.mark.parametrize("sw", [False, True])
def test_create_lut_block(ctx, O, sw):
    """The body of the create_look_up_table temperature loop (create_look_up_table.cpp:293-344): full-level
    pressure and pressure-weighted temperature from the block's half-level profile, then Planck- or
    solar-weighted averaging."""
    c = S.sw_spectrum_case(nlay=15, nwav=5000, ng=8, nband=2, seed=8) if sw else \
        S.lw_spectrum_case(nlay=15, nwav=5000, ng=8, nband=2, seed=8)
    p_hl, t_hl = c["pressure_hl"], c["temperature_hl"]
    p_fl = 0.5 * (p_hl[:-1] + p_hl[1:])
    t_x_p = t_hl * p_hl
    t_fl = 0.5 * (t_x_p[:-1] + t_x_p[1:]) / p_fl
    method = "transmission-3" if sw else "transmission"
    got = E.create_lut_block(ctx, c["ng"], c["vmr"], p_hl, t_hl, c["g_point"], c["od"], method,
                             wavenumber_cm_1=c["wavenumber"], d_wavenumber_cm_1=c["d_wavenumber"],
                             ssi=c["ssi"] if sw else None)
    assert np.array_equal(got[4], t_fl)
    w = c["ssi"] if sw else O.planck_function(t_fl, c["wavenumber"], c["d_wavenumber"])
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], p_fl, p_hl, c["g_point"], c["od"], w, method)
    S.assert_avg_close(got[:4], ref, method=method, scale=S.out_scale(dict(c, pressure_fl=p_fl)))
