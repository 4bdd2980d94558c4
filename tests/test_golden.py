"""Committed golden vectors (tests/golden/, produced by tests/golden/make_golden.py): the CPU
oracle and the NumPy restatement must reproduce them on any machine (this file runs without a
GPU), and the CUDA path must match them on the B200 (the gpu-marked tests)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as N
from tests import synthetic as S
from tests.oracle_lib import METHODS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _avg_case():
    d = np.load(os.path.join(GOLD, "average_optical_depth.npz"))
    c = dict(pressure_fl=d["pressure_fl"], pressure_hl=d["pressure_hl"], vmr=float(d["vmr"]))
    return d, c


def _ckd_case(name):
    d = np.load(os.path.join(GOLD, name))
    gases, vmr = [], []
    for i in range(int(d["ngas"])):
        g = dict(conc_dependence=int(d[f"gas{i}_conc_dependence"]), molar_abs=d[f"gas{i}_molar_abs"],
                 reference_vmr=float(d[f"gas{i}_reference_vmr"]))
        if f"gas{i}_vmr_lut" in d:
            g["vmr_lut"] = d[f"gas{i}_vmr_lut"]
        gases.append(g)
        vmr.append(d[f"gas{i}_vmr_fl"] if f"gas{i}_vmr_fl" in d else None)
    case = dict(gases=gases, vmr_fl=vmr, log_pressure=d["log_pressure"], temperature=d["temperature"],
                pressure_hl=d["pressure_hl"], temperature_hl=d["temperature_hl"],
                temperature_planck=d["temperature_planck"], planck_lut=d["planck_lut"])
    if "solar_irradiance" in d:
        case.update(solar_irradiance=d["solar_irradiance"], rayleigh_molar_scat=d["rayleigh_molar_scat"])
    return d, case


@pytest.mark.parametrize("method", METHODS)
def test_oracles_reproduce_golden_average(O, method):
    d, c = _avg_case()
    gold = tuple(d[f"{method}__{k}"] for k in ("molar_abs", "min", "max", "warn"))
    a = O.average_optical_depth_to_g_point(6, c["vmr"], c["pressure_fl"], c["pressure_hl"], d["g_point"], d["od"],
                                           d["planck"], method)
    S.assert_avg_close(a, gold, method=method, scale=S.out_scale(c), rtol=1e-13)
    b = N.average_optical_depth_to_g_point(6, c["vmr"], c["pressure_fl"], c["pressure_hl"], d["g_point"], d["od"],
                                           d["planck"], method)
    S.assert_avg_close(b, gold, method=method, scale=S.out_scale(c), rtol=2e-13)


def test_oracle_reproduces_golden_run_ckd_and_ranking(O):
    for name, sw in (("run_ckd_lw.npz", False), ("run_ckd_sw.npz", True)):
        d, case = _ckd_case(name)
        r = O.run_ckd(case, is_sw=sw)
        for k in ("optical_depth", "flux_dn", "bb_flux_dn") + (() if sw else ("flux_up", "planck_hl")):
            np.testing.assert_allclose(r[k], d[f"out_{k}"], rtol=1e-12, atol=1e-9, err_msg=f"{name} {k}")
    d = np.load(os.path.join(GOLD, "ranking.npz"))
    iband, rank, order = O.rank_bands(d["wavenumber"], d["key"], d["b1"], d["b2"])
    assert np.array_equal(rank, d["rank"]) and np.array_equal(iband, d["iband"])
    assert np.array_equal(N.rank_bands(d["wavenumber"], d["key"], d["b1"], d["b2"])[1], d["rank"])
    g, n_un = O.assign_g_points(d["gas_g"], d["g_min"], d["g_max"])
    assert np.array_equal(g, d["g_point"]) and n_un == int(d["n_unassigned"])


@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
def test_cuda_matches_golden_average(ctx, method):
    import ecckd_b200 as E

    d, c = _avg_case()
    gold = tuple(d[f"{method}__{k}"] for k in ("molar_abs", "min", "max", "warn"))
    got = E.average_optical_depth_to_g_point(ctx, 6, c["vmr"], c["pressure_fl"], c["pressure_hl"], d["g_point"],
                                             d["od"], d["planck"], method)
    S.assert_avg_close(got, gold, method=method, scale=S.out_scale(c))


@pytest.mark.gpu
def test_cuda_matches_golden_run_ckd_and_ranking(ctx):
    import ecckd_b200 as E

    for name, sw in (("run_ckd_lw.npz", False), ("run_ckd_sw.npz", True)):
        d, case = _ckd_case(name)
        if sw:
            m = E.CkdModel(ctx, case["gases"], case["log_pressure"], case["temperature"],
                           solar_irradiance=case["solar_irradiance"], rayleigh_molar_scat=case["rayleigh_molar_scat"])
        else:
            m = E.CkdModel(ctx, case["gases"], case["log_pressure"], case["temperature"], case["temperature_planck"],
                           case["planck_lut"])
        r = m.run_ckd(case["vmr_fl"], case["pressure_hl"], case["temperature_hl"])
        for k in ("flux_dn", "bb_flux_dn") + (() if sw else ("flux_up", "bb_flux_up")):
            assert np.max(np.abs(r[k] - d[f"out_{k}"])) <= S.FLUX_ATOL, (name, k)
    d = np.load(os.path.join(GOLD, "ranking.npz"))
    n = d["key"].size
    d_rank, d_ib = ctx.empty(n, np.int32), ctx.empty(n, np.int32)
    E.rank_bands_dev(ctx, n, ctx.to_device(d["wavenumber"]), ctx.to_device(d["key"]), d["b1"], d["b2"], d_ib, d_rank)
    assert np.array_equal(d_rank.numpy(), d["rank"]) and np.array_equal(d_ib.numpy(), d["iband"])
    d_gg, d_out = ctx.to_device(d["gas_g"], np.int32), ctx.empty(n, np.int32)
    n_un = E.assign_g_points_dev(ctx, n, d_gg, d["g_min"], d["g_max"], d_out)
    assert np.array_equal(d_out.numpy(), d["g_point"]) and n_un == int(d["n_unassigned"])
