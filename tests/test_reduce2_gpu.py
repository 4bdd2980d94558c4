"""The generations of the streaming reduction kernel behind ecckd_average_optical_depth_*
(average_optical_depth.cpp:22-197): reduce2_kernel on windowed (balanced) plans and reduce_kernel
on free plans must both match the oracle on the same inputs, agree
with each other, and be bitwise reproducible whatever the launch geometry knobs say (stress of the
mbarrier / bulk-copy pipelines: many tiny chunks, several wavenumber windows, repeated launches,
rows that are not 16-byte aligned, odd lengths)."""
import numpy as np
import pytest

import ecckd_b200 as E
from tests import synthetic as S

pytestmark = pytest.mark.gpu

METHODS2 = ["transmission", "transmission-10", "linear", "square-root"]  # the kinds reduce2_kernel runs


def _planck(O, c):
    return O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])


# (plan kind, kernel generation): free plan + reduce_kernel, windowed plan + each generation
GENS = [("0", "1"), ("1", "1"), ("1", "2")]


def _select(monkeypatch, gen):
    monkeypatch.setenv("ECCKD_PLAN_WINDOW", gen[0])
    monkeypatch.setenv("ECCKD_REDUCE_GEN", gen[1])


def _run(ctx, c, g, ng, planck, method):
    return E.average_optical_depth_to_g_point(ctx, ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"],
                                              planck, method)


def _cases():
    rng = np.random.default_rng(77)
    out = {}
    c = S.lw_spectrum_case(nlay=9, nwav=20011, ng=22, nband=4, seed=41)
    out["banded-coherent"] = (c, c["g_point"], c["ng"])
    c = S.lw_spectrum_case(nlay=5, nwav=9000, ng=40, nband=1, seed=12, coherent=False)
    out["one-band-40g-incoherent"] = (c, c["g_point"], c["ng"])  # every chunk spans several windows
    c = S.lw_spectrum_case(nlay=4, nwav=9000, ng=8, nband=1, seed=31)
    g = rng.integers(0, 150, size=c["g_point"].size).astype(np.int32)
    g[:1500] = np.repeat(np.arange(150, dtype=np.int32), 10)
    out["150g-random"] = (c, g, 150)
    c = S.lw_spectrum_case(nlay=6, nwav=15000, ng=12, nband=2, seed=9)
    g = c["g_point"].copy()
    g[::7] = -1          # wavenumbers in no g point
    g[g == 5] = 4        # an empty g point in the middle of a window
    out["holes-and-empty-g"] = (c, g, c["ng"])
    return out


CASES = _cases()


@pytest.mark.parametrize("gen", GENS)
@pytest.mark.parametrize("case", sorted(CASES))
def test_every_kernel_generation_matches_the_oracle(ctx, O, monkeypatch, case, gen):
    _select(monkeypatch, gen)
    window = "/".join(gen)
    c, g, ng = CASES[case]
    planck = _planck(O, c)
    for method in METHODS2:
        got = _run(ctx, c, g, ng, planck, method)
        ref = O.average_optical_depth_to_g_point(ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck,
                                                 method)
        S.assert_avg_close(got, ref, method=method, scale=S.out_scale(c), what=f"{case} window={window} {method} ")


@pytest.mark.parametrize("case", sorted(CASES))
def test_generations_agree_and_extremes_are_identical(ctx, O, monkeypatch, case):
    c, g, ng = CASES[case]
    planck = _planck(O, c)
    res = {}
    for gen in GENS:
        _select(monkeypatch, gen)
        res[gen] = _run(ctx, c, g, ng, planck, "linear")
    base = res[GENS[0]]
    for gen in GENS[1:]:
        np.testing.assert_allclose(res[gen][0], base[0], rtol=1e-13, atol=0.0)
        assert np.array_equal(res[gen][1], base[1]) and np.array_equal(res[gen][2], base[2]), gen  # selections
        assert np.array_equal(res[gen][3], base[3]), gen


@pytest.mark.parametrize("gen", GENS)
def test_pipeline_stress_tiny_chunks_windows_and_repeats(ctx, O, monkeypatch, gen):
    """Many tiny chunks (every mbarrier phase flips thousands of times per CTA), 1..7 wavenumber
    windows, odd row lengths (generic-store fallback followed by bulk copies into the same stage),
    20 spectra in one group; every configuration is launched three times and must reproduce itself
    bit for bit, and all of them must match the oracle."""
    _select(monkeypatch, gen)
    nspec = 20
    c = S.lw_spectrum_case(nlay=3, nwav=30001, ng=14, nband=3, seed=5)
    rng = np.random.default_rng(3)
    ods = [c["od"] * rng.uniform(0.2, 5.0) for _ in range(nspec)]
    planck = _planck(O, c)
    refs = [O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], od,
                                               planck, "transmission") for od in ods]
    d_g = ctx.to_device(c["g_point"], np.int32)
    d_wn, d_dwn = ctx.to_device(c["wavenumber"]), ctx.to_device(c["d_wavenumber"])
    d_ods = [ctx.to_device(od) for od in ods]
    nlay, ng, nwav = c["nlay"], c["ng"], c["nwav"]
    for ch, nwin in (("32", "1"), ("64", "3"), ("96", "7"), ("480", "2")):
        monkeypatch.setenv("ECCKD_CH", ch)
        monkeypatch.setenv("ECCKD_NWIN", nwin)
        plan = E.GPointPlan(ctx, d_g, nwav, ng)
        first = None
        for rep in range(3):
            outs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
            mins = [ctx.empty((nlay, ng)) for _ in range(nspec)]
            maxs = [ctx.empty((nlay, ng)) for _ in range(nspec)]
            E.average_optical_depth_batch_dev(ctx, plan, [c["vmr"]] * nspec, c["pressure_fl"], c["pressure_hl"], d_ods,
                                              nwav, E.WEIGHT_PLANCK, None, 0, c["temperature_fl"], d_wn, d_dwn,
                                              "transmission", outs, mins, maxs, ng, None)
            got = [(o.numpy(), a.numpy(), b.numpy()) for o, a, b in zip(outs, mins, maxs)]
            if first is None:
                first = got
                for s in range(nspec):
                    S.assert_avg_close(got[s] + (refs[s][3],), refs[s], method="transmission", scale=S.out_scale(c),
                                       what=f"ch={ch} nwin={nwin} spectrum {s} ")
            else:
                for s in range(nspec):
                    for a, b in zip(got[s], first[s]):
                        assert np.array_equal(a, b), f"ch={ch} nwin={nwin} rep={rep}: not reproducible"
        plan.close()


@pytest.mark.parametrize("gen", GENS)
def test_nan_optical_depth_is_reported_not_absorbed(ctx, O, monkeypatch, gen):
    _select(monkeypatch, gen)
    c = S.lw_spectrum_case(nlay=4, nwav=5000, ng=6, nband=2, seed=2)
    od = c["od"].copy()
    w_bad = 1234
    od[2, w_bad] = np.nan
    g_bad = int(c["g_point"][w_bad])
    planck = _planck(O, c)
    for method in ("transmission", "linear", "logarithmic"):
        got = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                                 od, planck, method)
        assert np.isnan(got[0][2, g_bad]), method
        assert got[3][g_bad] & 8, method  # ECCKD_WARN_NONFINITE
        ok = np.ones_like(got[0], dtype=bool)
        ok[2, g_bad] = False
        assert np.all(np.isfinite(got[0][ok])), method
        clean = E.average_optical_depth_to_g_point(ctx, c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"],
                                                   c["g_point"], c["od"], planck, method)
        assert np.array_equal(got[0][ok], clean[0][ok]), method


def _plan_check(info, rounds, g, ng):
    """Every wavenumber of a g point sits in exactly one valid cell of a lane of that g point, padding
    cells read a wavenumber of the lane's own g point; returns the mean shared-memory wavefronts per
    64-bit gather under the measured bank model (a warp is served as two half-warps, each costing the
    fullest of the 16 eight-byte bank pairs: tools/ubench/lds64_conflicts.cu)."""
    ch = info["ch"]
    seen = np.zeros(g.size, dtype=np.int32)
    wavefronts = loads = 0
    gv = np.where((g >= 0) & (g < ng), g, -1)
    for r in rounds:
        base = int(r["chunk"]) * ch
        lane_g = (r["meta"] & 0xFFF).astype(np.int64)
        used = lane_g != 0xFFF
        if info["balanced"]:
            mask = (r["meta"] >> 12) & 0xFFFF
        else:
            mask = (1 << ((r["meta"] >> 12) & 31)) - 1
        elem = r["off"].astype(np.int64) // 8          # [slot][lane]
        assert np.all(r["off"] % 8 == 0)
        for lane in np.nonzero(used)[0]:
            e = base + elem[:, lane]
            assert np.all(e < g.size), "offset beyond the spectrum"
            assert np.all(gv[e] == lane_g[lane]), "a cell reads a wavenumber of another g point"
            valid = ((mask[lane] >> np.arange(16)) & 1).astype(bool)
            np.add.at(seen, e[valid], 1)
        for s in range(16):
            for h in (0, 1):
                a = elem[s, 16 * h:16 * h + 16]  # unused lanes load too (their weight is zero)
                loads += 1
                worst = 0
                for q in np.unique(a % 16):
                    worst = max(worst, np.unique(a[a % 16 == q]).size)
                wavefronts += worst
    assert np.array_equal(seen, (gv >= 0).astype(np.int32)), "a wavenumber is missing or visited twice"
    return 2.0 * wavefronts / loads


@pytest.mark.parametrize("kind", ["coherent", "incoherent", "150g"])
def test_balanced_plan_is_valid_and_nearly_conflict_free(ctx, monkeypatch, kind):
    monkeypatch.setenv("ECCKD_PLAN_WINDOW", "1")
    if kind == "coherent":
        c = S.lw_spectrum_case(nlay=1, nwav=48011, ng=22, nband=4, seed=41)
        g, ng = c["g_point"], c["ng"]
    elif kind == "incoherent":
        c = S.lw_spectrum_case(nlay=1, nwav=20000, ng=12, nband=2, seed=12, coherent=False)
        g, ng = c["g_point"], c["ng"]
    else:
        rng = np.random.default_rng(5)
        g, ng = rng.integers(-1, 150, size=9000).astype(np.int32), 150
    d_g = ctx.to_device(g, np.int32)
    res = {}
    for balance in ("0", "1"):
        monkeypatch.setenv("ECCKD_PLAN_BALANCE", balance)
        plan = E.GPointPlan(ctx, d_g, g.size, ng)
        info, rounds = plan.export()
        assert info["windowed"] and info["balanced"] == (balance == "1")
        res[balance] = _plan_check(info, rounds, g, ng)
        plan.close()
    # the adjacent-lane placement costs 3-4 wavefronts per gather; the balanced one is close to the ideal 2
    assert res["1"] < res["0"]
    if kind != "150g":
        assert res["1"] <= 2.35, res


@pytest.mark.parametrize("method", ["transmission", "linear"])
def test_group_of_19_spectra_with_the_floating_row(ctx, O, monkeypatch, method):
    """A LUT temperature index is a group of 19 spectra sharing one Planck weight row: 10 consumer warps, of
    which the last two pass the 19th row between them every chunk visit (reduce2_kernel's floating row, whose
    two partial results combine_kernel adds).  Every row - the floating one included - must match the oracle,
    the call must be bitwise reproducible, and the min/max must not depend on whether the row floats."""
    nspec, nlay, nwav = 19, 3, 14011  # odd length, several chunk visits per CTA piece
    c = S.lw_spectrum_case(nlay=nlay, nwav=nwav, ng=22, nband=4, seed=5)
    rng = np.random.default_rng(19)
    ods = [c["od"] * np.exp(rng.normal(0.0, 1.0, (nlay, 1))) * (0.2 + 0.15 * s) for s in range(nspec)]
    vmrs = [c["vmr"] * (1.0 + 0.1 * s) for s in range(nspec)]
    planck = _planck(O, c)
    d_g = ctx.to_device(c["g_point"], np.int32)
    d_wn, d_dwn = ctx.to_device(c["wavenumber"]), ctx.to_device(c["d_wavenumber"])
    d_ods = [ctx.to_device(od) for od in ods]

    def run():
        plan = E.GPointPlan(ctx, d_g, nwav, c["ng"])
        outs = [ctx.empty((nlay, c["ng"])) for _ in range(nspec)]
        mins = [ctx.empty((nlay, c["ng"])) for _ in range(nspec)]
        maxs = [ctx.empty((nlay, c["ng"])) for _ in range(nspec)]
        E.average_optical_depth_batch_dev(ctx, plan, vmrs, c["pressure_fl"], c["pressure_hl"], d_ods, nwav,
                                          E.WEIGHT_PLANCK, None, 0, c["temperature_fl"], d_wn, d_dwn, method, outs, mins,
                                          maxs, c["ng"])
        res = [(o.numpy(), a.numpy(), b.numpy()) for o, a, b in zip(outs, mins, maxs)]
        plan.close()
        return res

    got = run()
    again = run()
    for s in range(nspec):
        ref = O.average_optical_depth_to_g_point(c["ng"], vmrs[s], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                                 ods[s], planck, method)
        cs = dict(vmr=vmrs[s], pressure_fl=c["pressure_fl"], pressure_hl=c["pressure_hl"])
        S.assert_avg_close(got[s] + (None,), ref, method=method, scale=S.out_scale(cs), what=f"spectrum {s}: ")
        for a, b in zip(got[s], again[s]):
            assert np.array_equal(a, b), f"spectrum {s}: not reproducible"
    monkeypatch.setenv("ECCKD_NO_FLOAT_ROW", "1")
    fixed = run()
    for s in range(nspec):
        assert np.array_equal(got[s][1], fixed[s][1]) and np.array_equal(got[s][2], fixed[s][2]), s
        if s < nspec - 1:  # rows that never float are summed in the same order either way
            assert np.array_equal(got[s][0], fixed[s][0]), s
        else:
            cs = dict(vmr=vmrs[s], pressure_fl=c["pressure_fl"], pressure_hl=c["pressure_hl"])
            ref = O.average_optical_depth_to_g_point(c["ng"], vmrs[s], c["pressure_fl"], c["pressure_hl"],
                                                     c["g_point"], ods[s], planck, method)
            S.assert_avg_close(fixed[s] + (None,), ref, method=method, scale=S.out_scale(cs), what="fixed rows: ")
