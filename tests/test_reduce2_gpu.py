"""The generations of the streaming reduction kernel behind ecckd_average_optical_depth_*
(average_optical_depth.cpp:22-197): reduce3_kernel / reduce2_kernel on windowed plans and
reduce_kernel on free and on windowed plans must all match the oracle on the same inputs, agree
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
GENS = [("0", "1"), ("1", "1"), ("1", "2"), ("1", "3")]


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
