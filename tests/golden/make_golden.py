"""Regenerates tests/golden/*.npz: small seeded inputs plus the outputs of the CPU oracle
(oracle/ecckd_oracle.cpp).  The reference ships no golden vectors for this path (SURVEY.md
section 4) and cannot be built here (no Adept / NetCDF), so these vectors freeze the oracle that
tests/test_oracle_crosscheck.py pins analytically and against the NumPy restatement.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import synthetic as S  # noqa: E402
from tests.oracle_lib import METHODS, oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O = oracle()
    # --- spectral reduction, every averaging method
    c = S.lw_spectrum_case(nlay=6, nwav=700, ng=6, nband=2, seed=101)
    g = c["g_point"].copy()
    g[g == 4] = 3      # an empty g point
    g[::29] = -1       # wavenumbers in no g point
    planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
    out = dict(pressure_fl=c["pressure_fl"], pressure_hl=c["pressure_hl"], temperature_fl=c["temperature_fl"],
               wavenumber=c["wavenumber"], d_wavenumber=c["d_wavenumber"], od=c["od"], g_point=g, planck=planck,
               vmr=np.array(c["vmr"]), ng=np.array(6))
    for m in METHODS:
        r = O.average_optical_depth_to_g_point(6, c["vmr"], c["pressure_fl"], c["pressure_hl"], g, c["od"], planck, m)
        for name, a in zip(("molar_abs", "min", "max", "warn"), r):
            out[f"{m}__{name}"] = a
    np.savez_compressed(os.path.join(HERE, "average_optical_depth.npz"), **out)

    # --- run_ckd LW + SW
    for sw in (False, True):
        ck = S.ckd_model_case(ncol=4, nlay=12, ng=6, seed=202 + int(sw), nconc=5, np_lut=53, sw=sw)
        r = O.run_ckd(ck, is_sw=sw)
        d = dict(log_pressure=ck["log_pressure"], temperature=ck["temperature"], pressure_hl=ck["pressure_hl"],
                 temperature_hl=ck["temperature_hl"], temperature_planck=ck["temperature_planck"],
                 planck_lut=ck["planck_lut"], ngas=np.array(len(ck["gases"])))
        if sw:
            d.update(solar_irradiance=ck["solar_irradiance"], rayleigh_molar_scat=ck["rayleigh_molar_scat"])
        for i, (gas, v) in enumerate(zip(ck["gases"], ck["vmr_fl"])):
            d[f"gas{i}_conc_dependence"] = np.array(gas["conc_dependence"])
            d[f"gas{i}_molar_abs"] = gas["molar_abs"]
            d[f"gas{i}_reference_vmr"] = np.array(gas.get("reference_vmr", 0.0))
            if gas.get("vmr_lut") is not None:
                d[f"gas{i}_vmr_lut"] = gas["vmr_lut"]
            if v is not None:
                d[f"gas{i}_vmr_fl"] = v
        for k, a in r.items():
            d[f"out_{k}"] = a
        np.savez_compressed(os.path.join(HERE, f"run_ckd_{'sw' if sw else 'lw'}.npz"), **d)

    # --- ranking with ties and signed zeros, g-point assignment
    rng = np.random.default_rng(303)
    n = 3000
    wn = np.linspace(1.0, 100.0, n)
    key = np.round(rng.normal(size=n), 1)
    key[::40] = 0.0
    key[20::40] = -0.0
    b1, b2 = np.array([1.0, 30.0, 70.0]), np.array([30.0, 70.0, 100.0])
    iband, rank, order = O.rank_bands(wn, key, b1, b2)
    n_g = np.array([[2, 3, 1], [3, 1, 2]], dtype=np.int32)
    gas_g, svs = [], []
    for a in range(2):
        perm = rng.permutation(n).astype(np.int32)
        ng_a = int(n_g[a].sum())
        cuts = np.sort(rng.choice(np.arange(1, n - 1), ng_a - 1, replace=False))
        r1 = np.concatenate([[0], cuts]).astype(np.int32)
        r2 = np.concatenate([cuts - 1, [n - 1]]).astype(np.int32)
        gas_g.append(O.store_g_points(perm, r1, r2))
        svs.append(np.sort(rng.uniform(0, 5, ng_a)))
    ng, gmin, gmax, band = O.overlap_g_points(n_g, svs)
    gp, n_un = O.assign_g_points(np.stack(gas_g), gmin, gmax)
    np.savez_compressed(os.path.join(HERE, "ranking.npz"), wavenumber=wn, key=key, b1=b1, b2=b2, iband=iband,
                        rank=rank, ordered_index=order, n_g=n_g, sv0=svs[0], sv1=svs[1], gas_g=np.stack(gas_g),
                        g_min=gmin, g_max=gmax, band_number=band, g_point=gp, n_unassigned=np.array(n_un))


if __name__ == "__main__":
    main()
