"""The oracle held against the reference's OWN tool sources (no GPU).

oracle/build_ref_tools.sh compiles run_ckd.cpp, create_look_up_table.cpp, reorder_spectrum.cpp,
find_g_points.cpp, ckd_model.cpp, read_spectrum.cpp, read_merged_spectrum.cpp, write_order.cpp,
equipartition.cpp and the hot-path units, unmodified and where they lie under /root/reference,
against the Adept API shim and an in-memory stand-in for the NetCDF layer.  These tests pin what
tests/test_reference_sources.py could not reach (the units that need DataFile / OutputDataFile):
  * CkdModel::calc_optical_depth / calc_planck_function / calc_rayleigh_optical_depth
    (ckd_model.cpp:923-1145, ckd_model.h:242-252) on a model read by CkdModel::read (:30-287);
  * fit_optical_depth_{lw,sw,sw_total_trans} and CkdEquipartition::calc_error
    (find_g_points.cpp:54-204, 291-405);
  * the tools end to end: run_ckd (run_ckd.cpp:27-371), create_look_up_table
    (create_look_up_table.cpp:28-601), reorder_spectrum (reorder_spectrum.cpp:37-311) - files as the
    tools write them (FLOAT / INT / SHORT variables), so those comparisons carry float32 rounding.
The library travels to the GPU box prebuilt; where it is absent the module is skipped."""
import numpy as np
import pytest

from tests import ref_tools as RT
from tests import synthetic as S
from tests.oracle_lib import METHODS
from tests.test_reference_sources import interval_case

pytestmark = pytest.mark.skipif(not RT.available(), reason="oracle/_ref/libecckd_ref_tools.so not built")

F32 = 2.0 ** -23  # the tools write FLOAT variables


@pytest.fixture(scope="module")
def R():
    return RT.ref_tools()


def put_ckd_definition(R, fname, case, sw=False):
    """A ckd-definition file as CkdModel::read expects it (ckd_model.cpp:30-287), from
    tests.synthetic.ckd_model_case."""
    ng = case["ng"]
    if sw:
        R.put(fname, "solar_irradiance", case["solar_irradiance"])
        R.put(fname, "rayleigh_molar_scattering_coeff", case["rayleigh_molar_scat"])
    else:
        R.put(fname, "temperature_planck", case["temperature_planck"])
        R.put(fname, "planck_function", case["planck_lut"])
    R.put(fname, "temperature", case["temperature"])
    R.put(fname, "pressure", np.exp(case["log_pressure"]))
    R.put(fname, "wavenumber1", [0.0, 10.0])
    R.put(fname, "wavenumber2", [10.0, 20.0])
    R.put(fname, "gpoint_fraction", np.full((ng, 2), 0.5))
    R.put(fname, "wavenumber1_band", [0.0])
    R.put(fname, "wavenumber2_band", [20.0])
    R.put(fname, "band_number", np.zeros(ng))
    R.put(fname, "n_gases", np.array(float(len(case["gases"]))))
    R.put_string(fname, "constituent_id", " ".join(g["name"] for g in case["gases"]))
    for k in ("history", "summary", "config", "model_id"):
        R.put_string(fname, k, "test")
    for g in case["gases"]:
        m = g["name"]
        R.put(fname, m + "_molar_absorption_coeff", g["molar_abs"])
        if g["conc_dependence"] == 2:
            R.put(fname, m + "_mole_fraction", g["vmr_lut"])
        else:
            R.put(fname, m + "_conc_dependence_code", np.array(float(g["conc_dependence"])))
            if g["conc_dependence"] == 0:
                R.put(fname, m + "_mole_fraction", np.full((2, 3), 0.5))  # composite_vmr [gas][level] (unused by the hot path)
                R.put_string(fname, m + "_constituent_id", "n2 o2")
            if g["conc_dependence"] == 3:
                R.put(fname, m + "_reference_mole_fraction", np.array(g["reference_vmr"]))


# ------------------------------------------------------------------ CkdModel (a18-a20)
@pytest.mark.parametrize("sw", [False, True])
def test_ckd_model_methods_match_the_reference(R, O, sw):
    case = S.ckd_model_case(ncol=4, nlay=21, ng=7, seed=17, sw=sw)
    R.reset()
    put_ckd_definition(R, "ckd.nc", case, sw)
    t_fl = O.temperature_fl(case["temperature_hl"], case["pressure_hl"])
    for igas, g in enumerate(case["gases"]):
        ref = R.ckd_calc_optical_depth("ckd.nc", igas, case["pressure_hl"], t_fl, case["vmr_fl"][igas], case["ng"])
        got = O.ckd_calc_optical_depth(g, case["log_pressure"], case["temperature"], case["pressure_hl"], t_fl,
                                       case["vmr_fl"][igas])
        # (the interpolation weights are differences of rounded quotients: a few 1e-14 either way)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-300, err_msg=g["name"])
    if sw:
        ref = R.ckd_calc_rayleigh_optical_depth("ckd.nc", case["pressure_hl"], case["ng"])
        dp = case["pressure_hl"][:, 1:] - case["pressure_hl"][:, :-1]
        got = dp[:, :, None] * (1.0 / (9.80665 * 0.001 * 28.970)) * case["rayleigh_molar_scat"][None, None, :]
        np.testing.assert_allclose(got, ref, rtol=1e-14)
    else:
        t = np.concatenate([case["temperature_hl"].ravel(), [90.0, 119.999, 120.0, 349.5, 350.0, 351.0, 420.0]])
        ref = R.ckd_calc_planck_function("ckd.nc", t, case["ng"])
        got = O.ckd_calc_planck_function(case["temperature_planck"], case["planck_lut"], t)
        np.testing.assert_allclose(got, ref, rtol=1e-14)


# ------------------------------------------------------------------ interval fits and calc_error (a13)
@pytest.mark.parametrize("method", ["linear", "transmission", "transmission-2", "square-root", "logarithmic"])
def test_fit_optical_depth_lw_matches_the_reference(R, O, method):
    c = interval_case(nlay=15, n=700, seed=3)
    metric = c["od"] * c["rng"].uniform(0.1, 2.0, c["od"].shape)
    if method == "logarithmic":
        metric[:, ::17] = 0.0  # the reference's zero handling (find_g_points.cpp:76-100)
        metric[4, 100:140] = 0.0
    for i1, i2 in ((0, 699), (5, 5), (100, 139), (350, 698)):
        ref = R.fit_optical_depth_lw(method, i1, i2, c["planck"], metric)
        got = O.fit_optical_depth_lw(method, i1, i2, c["planck"], metric)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-300, err_msg=f"{method} [{i1},{i2}]")


@pytest.mark.parametrize("method", ["linear", "transmission", "transmission-2", "square-root"])
def test_fit_optical_depth_sw_matches_the_reference(R, O, method):
    c = interval_case(nlay=15, n=700, seed=4)
    ssi = c["rng"].uniform(0.1, 3.0, 700)
    metric = c["od"] * c["rng"].uniform(0.1, 2.0, c["od"].shape)
    for i1, i2 in ((0, 699), (7, 7), (300, 420)):
        ref = R.fit_optical_depth_sw(method, i1, i2, ssi, metric)
        got = O.fit_optical_depth_sw(method, i1, i2, ssi, metric)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-300, err_msg=f"{method} [{i1},{i2}]")


def test_fit_optical_depth_sw_total_trans_matches_the_reference(R, O):
    c = interval_case(nlay=15, n=500, seed=5)
    ssi = c["rng"].uniform(0.1, 3.0, 500)
    bg = c["od"] * 0.3
    od = c["od"] * c["rng"].uniform(0.1, 2.0, c["od"].shape)
    for i1, i2 in ((0, 499), (10, 60), (3, 3)):
        ref = R.fit_optical_depth_sw_total_trans(i1, i2, ssi, bg, od)
        got = O.fit_optical_depth_sw_total_trans(i1, i2, ssi, bg, od)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)


def _bounds(rng, n, k=7):
    cuts = np.sort(rng.uniform(0.0, 1.0, k - 1))
    b1 = np.concatenate([[0.0], cuts, [0.0, 0.3]])
    b2 = np.concatenate([cuts, [1.0], [1.0, 0.3 + 0.5 / n]])
    return b1, b2


@pytest.mark.parametrize("method", ["linear", "transmission", "logarithmic"])
def test_calc_error_lw_matches_the_reference(R, O, method):
    """CkdEquipartition::calc_error (find_g_points.cpp:291-405) = bounds -> indices (:282-317), fit, then
    calc_cost_function_lw on the interval's slab."""
    c = interval_case(nlay=12, n=900, seed=9)
    rng = c["rng"]
    true_dn, true_up = O.radiative_transfer_lw(c["planck"], c["od"] * 1.3, c["emis"], c["surf"])
    hr = O.heating_rate(c["p"], true_dn, true_up)
    od_gas = c["od"] * rng.uniform(0.1, 2.0, c["od"].shape)
    metric = 1.0 - np.exp(-1.66 * od_gas) if method == "transmission" else od_gas
    band = R.band_lw(method, 0.2, c["lw"], c["p"], c["emis"], c["surf"], true_dn[-1], true_up[0], c["planck"], c["od"],
                     metric, hr)
    b1, b2 = _bounds(rng, 900)
    ref = R.calc_error_lw(band, b1, b2)
    n = 900
    for k in range(b1.size):
        i1, i2 = int(np.ceil(b1[k] * (n - 1))), int(np.floor(b2[k] * (n - 1)))
        if i2 < i1:
            i2 = i1
        f = O.fit_optical_depth_lw(method, i1, i2, c["planck"], metric)
        sl = slice(i1, i2 + 1)
        e = O.calc_cost_function_lw(c["p"], c["planck"][:, sl], c["emis"][sl], c["surf"][sl], c["od"][:, sl], f,
                                    true_dn[-1, sl], true_up[0, sl], hr[:, sl], 0.2, c["lw"])
        assert abs(e - ref[k]) <= 1e-12 * abs(ref[k]) + 1e-15, (k, e, ref[k])


@pytest.mark.parametrize("albedo", [0.0, 0.15])
def test_calc_error_sw_matches_the_reference(R, O, albedo):
    c = interval_case(nlay=12, n=800, seed=10)
    rng = c["rng"]
    ssi = rng.uniform(0.1, 3.0, 800)
    if albedo > 0.0:
        true_dn, true_up = O.radiative_transfer_norayleigh_sw(0.5, ssi, c["od"] * 1.2, np.full(800, albedo))
    else:
        true_dn = O.radiative_transfer_direct_sw(0.5, ssi, c["od"] * 1.2)
        true_up = np.zeros_like(true_dn)
    hr = O.heating_rate(c["p"], true_dn)
    metric = c["od"] * rng.uniform(0.1, 2.0, c["od"].shape)
    band = R.band_sw("linear", 0.2, 0.5, albedo, c["lw"], c["p"], ssi, true_dn[-1], true_up[0], c["od"], metric, hr)
    b1, b2 = _bounds(rng, 800)
    ref = R.calc_error_sw(band, b1, b2)
    n = 800
    for k in range(b1.size):
        i1, i2 = int(np.ceil(b1[k] * (n - 1))), int(np.floor(b2[k] * (n - 1)))
        if i2 < i1:
            i2 = i1
        f = O.fit_optical_depth_sw("linear", i1, i2, ssi, metric)
        sl = slice(i1, i2 + 1)
        e = O.calc_cost_function_sw(0.5, c["p"], ssi[sl], albedo, c["od"][:, sl], f, true_dn[-1, sl], true_up[0, sl],
                                    hr[:, sl], 0.2, c["lw"])
        assert abs(e - ref[k]) <= 1e-12 * abs(ref[k]) + 1e-15, (k, e, ref[k])


# ------------------------------------------------------------------ run_ckd, the tool (config A shape)
def test_run_ckd_tool_matches_the_oracle(R, O, tmp_path):
    case = S.ckd_model_case(ncol=6, nlay=24, ng=8, seed=23)
    R.reset()
    put_ckd_definition(R, "ckd.nc", case)
    R.put("input.nc", "temperature_hl", case["temperature_hl"])
    R.put("input.nc", "pressure_hl", case["pressure_hl"])
    for k in ("experiment", "experiment_id", "sub_experiment", "sub_experiment_id"):
        R.put_string("input.nc", k, "x")
    for g, v in zip(case["gases"], case["vmr_fl"]):
        if v is not None:
            R.put("input.nc", g["name"] + "_mole_fraction_fl", v)
    cfg = tmp_path / "run.cfg"
    cfg.write_text("ckd_model ckd.nc\ninput input.nc\noutput out.nc\nlog_level error\n")
    assert R.run("run_ckd", [str(cfg)]) == 0
    ref = O.run_ckd(case, is_sw=False)
    for name, key in (("optical_depth", "optical_depth"), ("spectral_flux_dn_lw", "flux_dn"), ("spectral_flux_up_lw", "flux_up"),
                      ("planck_hl", "planck_hl")):
        got = R.get("out.nc", name)
        np.testing.assert_allclose(got, ref[key], rtol=4 * F32, atol=1e-30, err_msg=name)
    np.testing.assert_allclose(R.get("out.nc", "flux_dn_lw"), ref["flux_dn"].sum(axis=2), rtol=4 * F32)
    np.testing.assert_allclose(R.get("out.nc", "flux_up_lw"), ref["flux_up"].sum(axis=2), rtol=4 * F32)


# ------------------------------------------------------------------ create_look_up_table, the tool (config B shape)
def _put_spectrum(R, fname, c, od, vmr):
    R.put(fname, "pressure_hl", c["pressure_hl"][None, :])
    R.put(fname, "temperature_hl", c["temperature_hl"][None, :])
    R.put(fname, "wavenumber", c["wavenumber"])
    R.put(fname, "d_wavenumber", c["d_wavenumber"])
    R.put(fname, "optical_depth", od[None, :, :])
    R.put(fname, "mole_fraction_fl", np.full((1, c["nlay"]), vmr))
    R.put(fname, "reference_surface_mole_fraction", np.array(vmr))
    R.put_string(fname, "constituent_id", "o3")


@pytest.mark.parametrize("method", ["transmission", "logarithmic", "linear"])
def test_create_look_up_table_tool_matches_the_oracle(R, O, tmp_path, method):
    c = S.lw_spectrum_case(nlay=9, nwav=2500, ng=7, nband=2, seed=33, vmr=3.0e-6)
    R.reset()
    _put_spectrum(R, "o3_spectrum.h5", c, c["od"], c["vmr"])
    band_number = np.array([int(np.searchsorted(c["band_edges"][1:-1], c["wavenumber"][c["g_point"] == g][0],
                                                side="right")) for g in range(c["ng"])], dtype=float)
    R.put("gpoints.h5", "g_point", c["g_point"].astype(float))
    R.put("gpoints.h5", "wavenumber1_band", c["band_edges"][:-1])
    R.put("gpoints.h5", "wavenumber2_band", c["band_edges"][1:])
    R.put("gpoints.h5", "band_number", band_number)
    R.put_string("gpoints.h5", "history", "test")
    cfg = tmp_path / "lut.cfg"
    cfg.write_text(f"input gpoints.h5\noutput lut.nc\naveraging_method {method}\nlog_level error\ngases o3\n"
                   "\\begin o3\n  conc_dependence linear\n  input o3_spectrum.h5\n\\end o3\n")
    rc = R.run("create_look_up_table", [str(cfg)])
    assert rc == 0, (rc, R.L.ref_tool_last_exception())
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    ref = O.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"],
                                             c["od"], planck, method)
    got = R.get("lut.nc", "o3_molar_absorption_coeff")  # [temperature=1][pressure][g_point], FLOAT
    np.testing.assert_allclose(got[0], ref[0], rtol=4 * F32, atol=1e-30)
    np.testing.assert_allclose(R.get("lut.nc", "temperature")[0], c["temperature_fl"], rtol=4 * F32)
    # the Planck look-up table of the file is the row "a6" of the hot path (create_look_up_table.cpp:579-591)
    t_lut = np.arange(120.0, 351.0)
    np.testing.assert_allclose(R.get("lut.nc", "planck_function"),
                               O.planck_lut(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], t_lut),
                               rtol=4 * F32)
    w1, w2 = O.gpoint_fraction_grid(False, c["band_edges"][:-1], c["band_edges"][1:])
    np.testing.assert_allclose(R.get("lut.nc", "gpoint_fraction"),
                               O.gpoint_fraction(c["ng"], c["g_point"], c["wavenumber"], c["d_wavenumber"], w1, w2),
                               rtol=4 * F32, atol=1e-30)


# ------------------------------------------------------------------ reorder_spectrum, the tool (config D shape)
def test_reorder_spectrum_tool_rank_matches_the_oracle(R, O, tmp_path):
    c = S.lw_spectrum_case(nlay=14, nwav=4000, ng=4, nband=1, seed=44, vmr=4.0e-4, zero_frac=0.0)
    R.reset()
    od = c["od"] * 40.0
    _put_spectrum(R, "spec.h5", c, od, c["vmr"])
    b1, b2 = np.array([10.0, 900.0, 2000.0]), np.array([900.0, 2000.0, 3250.0])
    cfg = tmp_path / "reorder.cfg"
    cfg.write_text("input spec.h5\noutput order.h5\niprofile 0\nlog_level error\n"
                   "wavenumber1 10 900 2000\nwavenumber2 900 2000 3250\n")
    rc = R.run("reorder_spectrum", [str(cfg)])
    assert rc == 0, (rc, R.L.ref_tool_last_exception())
    key, col = O.sort_key_lw(c["pressure_hl"], c["wavenumber"], c["d_wavenumber"], od, threshold=0.5)
    np.testing.assert_allclose(R.get("order.h5", "sorting_variable"), key, rtol=4 * F32, atol=1e-6)
    band, rank, _order = O.rank_bands(c["wavenumber"], key, b1, b2)
    assert np.array_equal(R.get("order.h5", "rank").astype(np.int64), rank.astype(np.int64))
    assert np.array_equal(R.get("order.h5", "band_number").astype(np.int64), band.astype(np.int64))
