"""Seeded synthetic CKDMIP-shaped inputs (the 700 GB CKDMIP database is not
available; SURVEY.md section 8d defines the shapes) and the parity assertions
with their stated tolerances."""
from __future__ import annotations

import numpy as np

ACCEL_GRAVITY = 9.80665
MOLAR_MASS_DRY_AIR = 28.970


def pressure_grid(nlay, p_top=1.0, p_surf=1.0e5):
    """Half-level pressures, log-spaced from p_top (index 0 = TOA) to the surface."""
    return np.exp(np.linspace(np.log(p_top), np.log(p_surf), nlay + 1))


def smooth_spectrum(wavenumber, rng, ncomp=24, amp=2.5):
    """A smooth positive function of wavenumber with structure on many scales
    (stand-in for a line spectrum: adjacent wavenumbers have similar absorption)."""
    span = wavenumber[-1] - wavenumber[0]
    x = (wavenumber - wavenumber[0]) / span
    f = np.zeros_like(x)
    for j in range(ncomp):
        freq = 2.0 ** rng.uniform(0, 9)
        f += rng.normal() * np.sin(2 * np.pi * freq * x + rng.uniform(0, 2 * np.pi)) / np.sqrt(ncomp)
    return np.exp(amp * f)


def g_point_map(wavenumber, key, band_edges, ng_per_band, rng):
    """Rank `key` inside each band and cut the rank axis into g points at random
    fractions - how find_g_points produces a map (contiguous rank intervals)."""
    g = np.full(wavenumber.size, -1, dtype=np.int32)
    ig = 0
    nband = len(band_edges) - 1
    for b in range(nband):
        if b < nband - 1:
            m = (wavenumber >= band_edges[b]) & (wavenumber < band_edges[b + 1])
        else:
            m = (wavenumber >= band_edges[b]) & (wavenumber <= band_edges[b + 1])
        idx = np.nonzero(m)[0]
        order = np.argsort(key[idx], kind="stable")
        cuts = np.sort(rng.uniform(0.05, 0.95, ng_per_band[b] - 1))
        bounds = np.concatenate([[0], np.round(cuts * idx.size).astype(int), [idx.size]])
        for k in range(ng_per_band[b]):
            g[idx[order[bounds[k]:bounds[k + 1]]]] = ig
            ig += 1
    return g, ig


def split_g_points(ng, nband):
    base = [ng // nband] * nband
    for i in range(ng - sum(base)):
        base[i] += 1
    return base


def lw_spectrum_case(nlay=12, nwav=3000, ng=8, nband=2, seed=0, zero_frac=0.02, coherent=True, vmr=4.0e-4):
    """One (gas, concentration, temperature) block of a longwave LUT build."""
    rng = np.random.default_rng(seed)
    p_hl = pressure_grid(nlay)
    p_fl = 0.5 * (p_hl[:-1] + p_hl[1:])
    t_hl = 200.0 + 90.0 * (np.log(p_hl) / np.log(1.0e5)) + rng.normal(0, 2.0, nlay + 1)
    t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
    wn = np.linspace(10.0, 3250.0, nwav) + 0.0
    dwn = np.full(nwav, wn[1] - wn[0])
    base = smooth_spectrum(wn, rng)
    mass = (p_hl[1:] - p_hl[:-1]) * vmr / (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR)
    od = mass[:, None] * 3.0e-4 * base[None, :] * np.exp(0.3 * rng.normal(size=(nlay, nwav)))
    if zero_frac > 0:
        od[rng.uniform(size=od.shape) < zero_frac] = 0.0
    key = np.log(base) if coherent else rng.uniform(size=nwav)
    edges = np.linspace(wn[0], wn[-1], nband + 1)
    g, ng_out = g_point_map(wn, key, edges, split_g_points(ng, nband), rng)
    return dict(nlay=nlay, nwav=nwav, ng=ng_out, vmr=vmr, pressure_hl=p_hl, pressure_fl=p_fl, temperature_hl=t_hl,
                temperature_fl=t_fl, wavenumber=wn, d_wavenumber=dwn, od=od, g_point=g, band_edges=edges, key=key)


def solar_spectrum(wn, tsi=1361.0):
    """5777 K Planck shape normalised to tsi."""
    x = 1.438776877 * wn / 5777.0
    s = wn ** 3 / np.expm1(x)
    return s * (tsi / s.sum())


def sw_spectrum_case(nlay=12, nwav=3000, ng=8, nband=2, seed=1, zero_frac=0.02):
    c = lw_spectrum_case(nlay, nwav, ng, nband, seed, zero_frac)
    rng = np.random.default_rng(seed + 1000)
    wn = np.linspace(260.0, 49990.0, nwav)
    c["wavenumber"] = wn
    c["d_wavenumber"] = np.full(nwav, wn[1] - wn[0])
    c["ssi"] = solar_spectrum(wn)
    edges = np.linspace(wn[0], wn[-1], nband + 1)
    c["band_edges"] = edges
    c["g_point"], c["ng"] = g_point_map(wn, c["key"], edges, split_g_points(ng, nband), rng)
    return c


def ckd_model_case(ncol=5, nlay=16, ng=8, seed=0, nt=6, np_lut=53, nconc=12, sw=False):
    """A CKD definition (SURVEY.md 8d config A gas list) plus evaluation profiles."""
    rng = np.random.default_rng(seed)
    log_p = np.log(1.0) + np.arange(np_lut) * (np.log(10.0) / 10.0)  # 10 per decade from 1 Pa
    t_ref = 200.0 + 90.0 * (log_p / np.log(1.0e5)).clip(0, 1.1)
    temperature = t_ref[None, :] + (np.arange(nt) * 20.0 - 40.0)[:, None]

    def table(scale, *lead):
        return scale * np.exp(rng.normal(0.0, 1.5, size=lead + (nt, np_lut, ng)))

    gases = [
        dict(name="composite", conc_dependence=0, molar_abs=table(1e-6)),
        dict(name="h2o", conc_dependence=2, molar_abs=table(1e-2, nconc), vmr_lut=1e-7 * (10.0 ** (np.arange(nconc) * 0.5))),
        dict(name="o3", conc_dependence=1, molar_abs=table(1.0)),
        dict(name="co2", conc_dependence=1, molar_abs=table(1e-1)),
        dict(name="ch4", conc_dependence=3, molar_abs=table(1e-1), reference_vmr=1921e-9),
        dict(name="n2o", conc_dependence=3, molar_abs=table(1e-1), reference_vmr=332e-9),
        dict(name="cfc11", conc_dependence=1, molar_abs=table(10.0)),
        dict(name="cfc12", conc_dependence=1, molar_abs=table(10.0)),
    ]
    p_surf = rng.uniform(0.95e5, 1.03e5, ncol)
    p_hl = np.exp(np.linspace(np.log(1.0), np.log(p_surf), nlay + 1)).T.copy()
    t_hl = 190.0 + 105.0 * (np.log(p_hl) / np.log(1.0e5)) + rng.normal(0, 3.0, (ncol, nlay + 1))
    t_hl[:, 0] = rng.uniform(100.0, 400.0, ncol)  # exercise the Planck extrapolation branches
    t_hl[0, 0] = 101.5
    t_hl[-1, 0] = 383.25
    p_fl = 0.5 * (p_hl[:, 1:] + p_hl[:, :-1])
    shape = (ncol, nlay)
    vmr = [
        None,
        np.exp(rng.uniform(np.log(5e-8), np.log(5e-2), shape)),  # beyond both ends of the LUT axis
        np.exp(rng.uniform(np.log(1e-8), np.log(1e-5), shape)),
        np.full(shape, 415e-6),
        np.full(shape, 1921e-9) * rng.uniform(0.5, 2.0, shape),
        np.full(shape, 332e-9) * rng.uniform(0.5, 2.0, shape),
        np.full(shape, 2.3e-10),
        np.full(shape, 5e-10),
    ]
    case = dict(gases=gases, log_pressure=log_p, temperature=temperature, pressure_hl=p_hl, temperature_hl=t_hl,
                pressure_fl=p_fl, vmr_fl=vmr, ncol=ncol, nlay=nlay, ng=ng)
    t_planck = np.arange(120.0, 351.0)
    case["temperature_planck"] = t_planck
    case["planck_lut"] = (5.670374419e-8 * t_planck[:, None] ** 4) * rng.dirichlet(np.ones(ng))[None, :]
    if sw:
        case["solar_irradiance"] = 1361.0 * rng.dirichlet(np.ones(ng)) * 0.97
        case["rayleigh_molar_scat"] = np.exp(rng.uniform(np.log(1e-7), np.log(1e-3), ng))
    return case


# ------------------------------------------------------------------ tolerances
# north_star: averaged optical depths / LUT coefficients relative 1e-12 (FP64),
# fluxes <= 1e-6 W m-2, integer outputs bit-exact.
AVG_RTOL = 1.0e-12
FLUX_ATOL = 1.0e-6
EPS = 2.0 ** -53

_KD = {"transmission": 1.66, "transmission-2": 1.66 * 2.0, "transmission-3": 1.66 * 3.0,
       "transmission-10": 1.66 * 10.0, "hybrid-logarithmic-transmission-3": 1.66 * 3.0}


def out_scale(case):
    """Factor that turns fitted optical depth into the output (average_optical_depth.cpp:170-193)."""
    nlay = case["pressure_fl"].size
    if case["vmr"] > 0.0:
        dp = case["pressure_hl"][1:] - case["pressure_hl"][:-1]
        return ((ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR) / case["vmr"]) / dp
    return np.ones(nlay)


def avg_tolerance(ref_molar_abs, method="linear", scale=None, rtol=AVG_RTOL):
    """The tolerance array assert_avg_close applies to the averaged value (see there)."""
    b = ref_molar_abs
    tol = rtol * np.abs(b)
    kd = _KD.get(method)
    if kd is not None:
        sc = np.ones(b.shape[0]) if scale is None else scale
        tau = np.abs(b) / sc[:, None]
        tol = tol + 16 * EPS * np.exp(np.minimum(kd * tau, 700.0)) / kd * sc[:, None]
    return tol


def assert_avg_close(got, ref, method="linear", scale=None, rtol=AVG_RTOL, what="", planck_scale=None):
    """Compare (molar_abs, min, max, warn) tuples against the oracle.

    * min/max are selections of input values times the same factor: 4 ulp.
    * the averaged value: |got-ref| <= rtol*|ref| (1e-12, the north_star
      tolerance) plus, for the transmission methods only, the conditioning term
      16*eps*exp(kd*tau)/kd in optical-depth units.  Reason: the reference forms
      x = sum((1-exp(-kd*tau))*w)/sum(w) and returns -log(1-x)/kd; the absolute
      rounding error of x is a few eps whatever the implementation (each
      1-exp() is rounded at ulp(1) and the sum order is unspecified in the
      reference), and d(tau)/dx = exp(kd*tau)/kd.  For optically thin or
      moderately thick g points the term is far below rtol*tau; it only matters
      where the reference's own answer is rounding noise (tau -> 0 or x -> 1).
    """
    names = ("molar_abs", "min_molar_abs", "max_molar_abs")
    kd = _KD.get(method)
    for i, name in enumerate(names):
        if got[i] is None:
            continue
        a, b = got[i], ref[i]
        assert a.shape == b.shape, (name, a.shape, b.shape)
        assert np.all(np.isfinite(a) == np.isfinite(b)), f"{what}{name}: finiteness differs"
        fin = np.isfinite(b)
        if i > 0:
            tol = 4 * 2.3e-16 * np.abs(b)
        else:
            tol = rtol * np.abs(b)
            if kd is not None:
                sc = np.ones(b.shape[0]) if scale is None else scale
                tau = np.abs(b) / sc[:, None]
                tol = tol + 16 * EPS * np.exp(np.minimum(kd * tau, 700.0)) / kd * sc[:, None]
        err = np.abs(a - b)
        bad = fin & (err > tol)
        rel = err[fin] / np.maximum(np.abs(b[fin]), 1e-300)
        assert not bad.any(), (f"{what}{name}: max rel err {np.max(rel):.3e}, "
                               f"{bad.sum()} of {bad.size} out of tolerance")
    if len(got) > 3 and len(ref) > 3 and got[3] is not None and ref[3] is not None:
        assert np.array_equal(got[3] & 1, ref[3] & 1), f"{what}empty-g warning flags differ"
