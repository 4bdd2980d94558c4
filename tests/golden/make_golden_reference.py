"""Regenerates tests/golden/reference_sources.npz: small seeded inputs and the outputs of the
REFERENCE'S OWN CODE for them - oracle/_ref/libecckd_ref.so, i.e. the reference's translation units
compiled in place (oracle/build_ref_shim.sh) - so that the vectors can be checked where
/root/reference does not exist (the GPU box) and by the CUDA path.

Needs /root/reference (to build oracle/_ref).  Run from the repo root:
    python tests/golden/make_golden_reference.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import synthetic as S  # noqa: E402
from tests.oracle_lib import METHODS, reference_oracle  # noqa: E402
from tests.test_reference_sources import ckd_cost_case, interval_case, rt_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    R = reference_oracle()
    assert R is not None, "build oracle/_ref first: python -c 'import __graft_entry__ as g; g.build()'"
    out = {}
    # average_optical_depth_to_g_point + planck_function
    c = S.lw_spectrum_case(nlay=6, nwav=700, ng=6, nband=2, seed=111)
    g = c["g_point"].copy()
    g[::31] = -1
    planck = R.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    out.update(avg_pressure_fl=c["pressure_fl"], avg_pressure_hl=c["pressure_hl"], avg_temperature_fl=c["temperature_fl"],
               avg_wavenumber=c["wavenumber"], avg_d_wavenumber=c["d_wavenumber"], avg_od=c["od"], avg_g_point=g,
               avg_planck=planck, avg_vmr=np.array(c["vmr"]), avg_ng=np.array(c["ng"]))
    for m in METHODS:
        r = R.average_optical_depth_to_g_point(c["ng"], c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        for name, a in zip(("molar_abs", "min", "max"), r[:3]):
            out[f"avg_{m}__{name}"] = a
    # flux solvers
    od, pl, emis, surf = rt_inputs(nlay=9, n=40, seed=21)
    grey = np.exp(np.random.default_rng(22).normal(-3, 2, 9))
    p = np.linspace(1.0, 1.0e5, 10)
    ssi = np.random.default_rng(23).uniform(0.1, 3.0, 40)
    alb = np.random.default_rng(24).uniform(0.0, 0.6, 40)
    dn, up = R.radiative_transfer_lw(pl, od, emis, surf)
    out.update(rt_od=od, rt_planck=pl, rt_emis=emis, rt_surf=surf, rt_grey=grey, rt_p=p, rt_ssi=ssi, rt_alb=alb,
               rt_lw_dn=dn, rt_lw_up=up, rt_hr=R.heating_rate(p, dn, up), rt_hr_dn_only=R.heating_rate(p, dn, None))
    out["rt_lw_bb_dn"], out["rt_lw_bb_up"] = R.radiative_transfer_lw_bb(pl, od, grey, emis, surf)
    out["rt_sw_direct"] = R.radiative_transfer_direct_sw(0.37, ssi, od)
    out["rt_sw_dn"], out["rt_sw_up"] = R.radiative_transfer_norayleigh_sw(0.5, ssi, od, alb)
    out["rt_sw_bb_dn"], out["rt_sw_bb_up"] = R.radiative_transfer_sw_bb(0.5, ssi, od, grey * 0.1, 0.15)
    out["rt_sw_bb_direct"] = R.radiative_transfer_sw_bb(0.5, ssi, od, grey * 0.1, 0.15, direct_only=True)[0]
    wn = np.concatenate([np.linspace(275.0, 49975.0, 200), [19999.9, 20000.0, 20000.1]])
    out.update(ray_wn=wn, ray_coeff=R.rayleigh_molar_scattering_coeff(wn))
    # interval error
    ic = interval_case(nlay=8, n=90, seed=31)
    t_dn, t_up = R.radiative_transfer_lw(ic["planck"], ic["od"] * 1.3, ic["emis"], ic["surf"])
    hr = R.heating_rate(ic["p"], t_dn, t_up)
    out.update(ie_p=ic["p"], ie_planck=ic["planck"], ie_emis=ic["emis"], ie_surf=ic["surf"], ie_od=ic["od"], ie_fit=ic["fit"],
               ie_dn_surf=t_dn[-1], ie_up_toa=t_up[0], ie_hr=hr, ie_lw=ic["lw"])
    out["ie_cost_lw"] = np.array(R.calc_cost_function_lw(ic["p"], ic["planck"], ic["emis"], ic["surf"], ic["od"], ic["fit"],
                                                        t_dn[-1], t_up[0], hr, 0.2, ic["lw"]))
    issi = ic["rng"].uniform(0.1, 2.0, 90)
    s_dn, s_up = R.radiative_transfer_norayleigh_sw(0.5, issi, ic["od"] * 1.2, np.full(90, 0.15))
    s_hr = R.heating_rate(ic["p"], s_dn, None)
    out.update(ie_ssi=issi, ie_sw_dn_surf=s_dn[-1], ie_sw_up_toa=s_up[0], ie_sw_hr=s_hr)
    for alb_s in (0.0, 0.15):
        out[f"ie_cost_sw_{alb_s}"] = np.array(R.calc_cost_function_sw(0.5, ic["p"], issi, alb_s, ic["od"], ic["fit"] * 0.1,
                                                                      s_dn[-1], s_up[0], s_hr, 0.2, ic["lw"]))
    # forward cost
    cc = ckd_cost_case(nlay=10, ng=12, nband=4, seed=41)
    rng, bm = cc["rng"], cc["bm"]
    em = rng.uniform(0.9, 1.0, 4)
    f_dn, f_up = R.radiative_transfer_lw(cc["planck"], cc["od"] * 1.1, em[bm], cc["surf"])
    f_dn = np.stack([f_dn[:, bm == b].sum(1) for b in range(4)], 1)
    f_up = np.stack([f_up[:, bm == b].sum(1) for b in range(4)], 1)
    f_hr = R.heating_rate(cc["p"], f_dn, f_up)
    sfd, sfu = rng.uniform(0, 5, 12), rng.uniform(0, 5, 12)
    out.update(fc_p=cc["p"], fc_od=cc["od"], fc_planck=cc["planck"], fc_surf=cc["surf"], fc_bm=bm, fc_lw=cc["lw"], fc_emis=em,
               fc_dn=f_dn, fc_up=f_up, fc_hr=f_hr, fc_sfd=sfd, fc_sfu=sfu)
    out["fc_cost_lw"] = np.array(R.calc_cost_function_ckd_lw(cc["p"], cc["planck"], em, cc["surf"], cc["od"], f_dn, f_up, f_hr,
                                                            0.2, 0.2, 0.8, 0.1, cc["lw"], spectral_flux_dn_surf=sfd,
                                                            spectral_flux_up_toa=sfu, band_mapping=bm))
    fssi, falb = rng.uniform(1.0, 100.0, 12), rng.uniform(0.05, 0.5, 4)
    w_dn, w_up = R.radiative_transfer_norayleigh_sw(0.6, fssi, cc["od"] * 1.1, falb[bm])
    w_dn = np.stack([w_dn[:, bm == b].sum(1) for b in range(4)], 1)
    w_up = np.stack([w_up[:, bm == b].sum(1) for b in range(4)], 1)
    w_hr = R.heating_rate(cc["p"], w_dn, None)
    sbw = rng.uniform(0, 1, 12)
    per_band = np.zeros(4)
    out.update(fc_ssi=fssi, fc_albedo=falb, fc_sw_dn=w_dn, fc_sw_up=w_up, fc_sw_hr=w_hr, fc_sbw=sbw)
    out["fc_cost_sw"] = np.array(R.calc_cost_function_ckd_sw(0.6, cc["p"], fssi, falb, cc["od"], w_dn, w_up, w_hr, 0.2, 0.2, 0.8,
                                                            cc["lw"], spectral_boundary_weights=sbw,
                                                            spectral_flux_dn_surf=sfd, spectral_flux_up_toa=sfu,
                                                            band_mapping=bm, cost_fn_per_band=per_band))
    out["fc_cost_sw_per_band"] = per_band
    # g-point assignment (integer)
    rr = np.random.default_rng(51)
    rank = rr.permutation(3000).astype(np.int32)
    cuts = np.sort(rr.choice(np.arange(1, 2999), 5, replace=False))
    r1 = np.concatenate([[0], cuts]).astype(np.int32)
    r2 = np.concatenate([cuts - 1, [2999]]).astype(np.int32)
    r2[1] -= 25
    n_g = rr.integers(1, 5, (4, 3)).astype(np.int32)
    sv = [rr.uniform(0, 1, int(n_g[i].sum())) for i in range(4)]
    ng, gmin, gmax, band = R.overlap_g_points(n_g, sv)
    out.update(gp_rank=rank, gp_rank1=r1, gp_rank2=r2, gp_g_point=R.store_g_points(rank, r1, r2), gp_n_g=n_g,
               gp_sv=np.concatenate(sv), gp_ng=np.array(ng), gp_gmin=gmin, gp_gmax=gmax, gp_band=band)
    np.savez_compressed(os.path.join(HERE, "reference_sources.npz"), **out)
    print("wrote reference_sources.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
