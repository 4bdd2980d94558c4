"""A compiled C++ consumer of the C ABI: the reference's OWN tool mains (create_look_up_table.cpp, run_ckd.cpp,
reorder_spectrum.cpp - unmodified object code) linked once with the reference's hot-path units
(oracle/_ref/libecckd_ref_tools.so) and once with integration/ecckd_gpu_bindings.cpp in their place
(libecckd_ref_tools_gpu.so: average_optical_depth_to_g_point, planck_function and
radiative_transfer_lw<false> become calls into libecckd_b200.so, as INTEGRATION.md prescribes).
Both builds run the same configuration and input files; what they write must agree.

The two libraries are built by oracle/build_ref_tools.sh where /root/reference exists and travel to the
GPU box prebuilt; where they are absent the module is skipped."""
import subprocess

import numpy as np
import pytest

from tests import ref_tools as RT
from tests import synthetic as S
from tests.test_reference_tools import _put_spectrum, put_ckd_definition

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (RT.available() and RT.gpu_available()),
                                 reason="oracle/_ref/libecckd_ref_tools{,_gpu}.so not built")]

F32 = 2.0 ** -23  # the tools write FLOAT variables


@pytest.fixture(scope="module")
def builds():
    return {"stock": RT.RefTools(RT.LIB), "gpu": RT.RefTools(RT.GPU_LIB)}


def test_the_gpu_build_binds_the_c_abi():
    und = subprocess.run(["nm", "-D", "--undefined-only", RT.GPU_LIB], capture_output=True, text=True).stdout
    for sym in ("ecckd_ctx_create", "ecckd_average_optical_depth_to_g_point", "ecckd_planck_function",
                "ecckd_radiative_transfer_lw_dev", "ecckd_last_error"):
        assert f"U {sym}" in und, sym
    stock = subprocess.run(["nm", "-D", "--undefined-only", RT.LIB], capture_output=True, text=True).stdout
    assert "ecckd_" not in stock


@pytest.mark.parametrize("method", ["transmission", "logarithmic", "linear"])
def test_create_look_up_table_main_on_the_c_abi_writes_what_the_stock_build_writes(builds, tmp_path, method):
    c = S.lw_spectrum_case(nlay=9, nwav=20011, ng=7, nband=2, seed=33, vmr=3.0e-6)
    band_number = np.array([int(np.searchsorted(c["band_edges"][1:-1], c["wavenumber"][c["g_point"] == g][0],
                                                side="right")) for g in range(c["ng"])], dtype=float)
    cfg = tmp_path / "lut.cfg"
    cfg.write_text(f"input gpoints.h5\noutput lut.nc\naveraging_method {method}\nlog_level error\ngases o3\n"
                   "\\begin o3\n  conc_dependence linear\n  input o3_spectrum.h5\n\\end o3\n")
    out = {}
    for name, R in builds.items():
        R.reset()
        _put_spectrum(R, "o3_spectrum.h5", c, c["od"], c["vmr"])
        R.put("gpoints.h5", "g_point", c["g_point"].astype(float))
        R.put("gpoints.h5", "wavenumber1_band", c["band_edges"][:-1])
        R.put("gpoints.h5", "wavenumber2_band", c["band_edges"][1:])
        R.put("gpoints.h5", "band_number", band_number)
        R.put_string("gpoints.h5", "history", "test")
        rc = R.run("create_look_up_table", [str(cfg)])
        assert rc == 0, (name, rc, R.L.ref_tool_last_exception())
        out[name] = {v: R.get("lut.nc", v) for v in ("o3_molar_absorption_coeff", "o3_molar_absorption_coeff_min",
                                                     "o3_molar_absorption_coeff_max", "temperature", "planck_function",
                                                     "gpoint_fraction", "pressure")}
    for v, a in out["stock"].items():
        np.testing.assert_allclose(out["gpu"][v], a, rtol=2 * F32, atol=1e-30, err_msg=v)
    assert np.all(out["gpu"]["o3_molar_absorption_coeff"] > 0)


def test_run_ckd_main_on_the_c_abi_writes_what_the_stock_build_writes(builds, tmp_path):
    case = S.ckd_model_case(ncol=6, nlay=24, ng=32, seed=23)
    cfg = tmp_path / "run.cfg"
    cfg.write_text("ckd_model ckd.nc\ninput input.nc\noutput out.nc\nlog_level error\n")
    out = {}
    for name, R in builds.items():
        R.reset()
        put_ckd_definition(R, "ckd.nc", case)
        R.put("input.nc", "temperature_hl", case["temperature_hl"])
        R.put("input.nc", "pressure_hl", case["pressure_hl"])
        for k in ("experiment", "experiment_id", "sub_experiment", "sub_experiment_id"):
            R.put_string("input.nc", k, "x")
        for g, v in zip(case["gases"], case["vmr_fl"]):
            if v is not None:
                R.put("input.nc", g["name"] + "_mole_fraction_fl", v)
        assert R.run("run_ckd", [str(cfg)]) == 0, name
        out[name] = {v: R.get("out.nc", v) for v in ("optical_depth", "spectral_flux_dn_lw", "spectral_flux_up_lw",
                                                     "flux_dn_lw", "flux_up_lw", "planck_hl")}
    for v, a in out["stock"].items():
        np.testing.assert_allclose(out["gpu"][v], a, rtol=2 * F32, atol=1e-6, err_msg=v)
    assert out["gpu"]["flux_up_lw"].max() > 100.0


def test_reorder_spectrum_main_on_the_c_abi_ranks_like_the_stock_build(builds, tmp_path):
    """reorder_spectrum.cpp:121-190 calls planck_function and radiative_transfer_lw<false>; the rank it writes
    is std::stable_sort of keys that carry the two builds' rounding, so it may differ where keys nearly tie."""
    c = S.lw_spectrum_case(nlay=14, nwav=4000, ng=4, nband=1, seed=44, vmr=4.0e-4, zero_frac=0.0)
    cfg = tmp_path / "reorder.cfg"
    cfg.write_text("input spec.h5\noutput order.h5\niprofile 0\nlog_level error\n"
                   "wavenumber1 10 900 2000\nwavenumber2 900 2000 3250\n")
    out = {}
    for name, R in builds.items():
        R.reset()
        _put_spectrum(R, "spec.h5", c, c["od"] * 40.0, c["vmr"])
        rc = R.run("reorder_spectrum", [str(cfg)])
        assert rc == 0, (name, rc, R.L.ref_tool_last_exception())
        out[name] = {v: R.get("order.h5", v) for v in ("rank", "band_number", "sorting_variable")}
    np.testing.assert_allclose(out["gpu"]["sorting_variable"], out["stock"]["sorting_variable"], rtol=4 * F32, atol=1e-6)
    assert np.array_equal(out["gpu"]["band_number"], out["stock"]["band_number"])
    # keys that differ in the last bits can swap neighbours: the ranks agree except where keys tie to float32
    differ = np.mean(out["gpu"]["rank"] != out["stock"]["rank"])
    assert differ < 0.01, differ
