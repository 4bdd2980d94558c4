"""oracle_np.py - independent NumPy restatement of ecCKD's hot path.

TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).  This second
restatement is written array-style, straight from the formula sheet of
SURVEY.md appendix B and the reference sources cited per function, without
looking at ecckd_oracle.cpp's loops.  tests/test_oracle_crosscheck.py requires
the two to agree to ~1e-13, which is the strongest pin available while the
reference (Adept + NetCDF) cannot be built here.  Paths are relative to
/root/reference.
"""
from __future__ import annotations

import numpy as np

ACCEL_GRAVITY = 9.80665  # src/ecckd/constants.h:22
SPECIFIC_HEAT_AIR = 1004.0
LW_DIFFUSIVITY = 1.66
MOLAR_MASS_DRY_AIR = 28.970

METHODS = (
    "linear",
    "transmission",
    "transmission-2",
    "transmission-3",
    "transmission-10",
    "square-root",
    "logarithmic",
    "hybrid-logarithmic-transmission-3",
)


def planck_function(temperature, wavenumber_cm_1, d_wavenumber_cm_1):
    """src/ecckd/planck_function.cpp:22-54 -> [nt][nwav]."""
    h = 6.62606896e-34
    c = 2.99792458e8
    k = 1.3806504e-23
    inv_cm_2_hz = 100.0 * c
    pi = 3.14159265358979323846
    freq = wavenumber_cm_1 * inv_cm_2_hz
    prefactor = (d_wavenumber_cm_1 * 2.0 * h * inv_cm_2_hz * pi / (c * c)) * (freq * freq * freq)
    t = np.asarray(temperature, dtype=np.float64).reshape(-1, 1)
    return prefactor[None, :] / (np.exp((h / k) * (freq[None, :] / t)) - 1.0)


def _log_average(od_row, w_row, idx):
    """average_optical_depth.cpp:79-96 for one layer."""
    sel = idx[od_row[idx] > 0.0]
    if sel.size == idx.size:
        return np.exp(np.sum(np.log(od_row[sel]) * w_row[idx]) / np.sum(w_row[idx]))
    if sel.size == 0:
        return 0.0
    return np.exp(np.sum(np.log(od_row[sel]) * w_row[sel]) / np.sum(w_row[sel])) * (sel.size / idx.size)


def average_optical_depth_to_g_point(ng, reference_surface_vmr, pressure_fl, pressure_hl, g_point, od, weight, method):
    """src/ecckd/average_optical_depth.cpp:22-197.

    Returns (molar_abs, min_molar_abs, max_molar_abs, warn) with the first three
    [nlay][ng] and warn[ng] bit flags (1 empty g, 2 min>fit, 4 min>=max).
    """
    nlay = od.shape[0]
    if weight.ndim == 1:
        weight = np.broadcast_to(weight, od.shape)  # spread<0>(ssi, nlay)
    out = np.zeros((nlay, ng))
    out_min = np.zeros((nlay, ng))
    out_max = np.zeros((nlay, ng))
    warn = np.zeros(ng, dtype=np.int32)
    kd_of = {"transmission": 1.66 * 1.0, "transmission-2": 1.66 * 1.0 * 2.0,
             "transmission-3": 1.66 * 1.0 * 3.0, "transmission-10": 1.66 * 1.0 * 10.0}
    with np.errstate(all="ignore"):
        for ig in range(ng):
            idx = np.nonzero(g_point == ig)[0]
            o = od[:, idx]
            w = weight[:, idx]
            if method == "linear":
                fit = np.sum(o * w, axis=1) / np.sum(w, axis=1)
            elif method in kd_of:
                kd = kd_of[method]
                x = np.minimum(0.9999999999999999, np.sum((1.0 - np.exp(-o * kd)) * w, axis=1) / np.sum(w, axis=1))
                fit = np.abs(-np.log(1.0 - x) / kd)
            elif method == "square-root":
                fit = np.sum(np.sqrt(o) * w, axis=1) / np.sum(w, axis=1)
                fit = fit * fit
            elif method == "logarithmic":
                fit = np.array([_log_average(od[iz], weight[iz], idx) for iz in range(nlay)])
            elif method == "hybrid-logarithmic-transmission-3":
                fit = np.zeros(nlay)
                kd = kd_of["transmission-3"]
                for iz in range(nlay):
                    if pressure_fl[iz] > 100.0e2:
                        fit[iz] = _log_average(od[iz], weight[iz], idx)
                    else:
                        x = min(0.9999999999999999, np.sum((1.0 - np.exp(-o[iz] * kd)) * w[iz]) / np.sum(w[iz]))
                        fit[iz] = abs(-np.log(1.0 - x) / kd)
            else:
                raise ValueError(147)
            if idx.size == 0:
                warn[ig] |= 1
                fit = np.zeros(nlay)
                mn = np.zeros(nlay)
                mx = np.zeros(nlay)
            else:
                mn = o.min(axis=1)
                mx = o.max(axis=1)
                fit = np.maximum(mn, np.minimum(fit, mx))
                m = mn > fit
                if m.any():
                    warn[ig] |= 2
                    mn = np.where(m, fit, mn)
                m = (mn > 0.0) & (mn >= mx)
                if m.any():
                    warn[ig] |= 4
                    mn = np.where(m, mn * 0.99, mn)
                    mx = np.where(m, mx * 1.01, mx)
            if reference_surface_vmr > 0.0:
                f = (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR) / reference_surface_vmr
                dp = pressure_hl[1:] - pressure_hl[:-1]
                out[:, ig] = f * fit / dp
                out_min[:, ig] = f * mn / dp
                out_max[:, ig] = f * mx / dp
            else:
                out[:, ig] = fit
                out_min[:, ig] = mn
                out_max[:, ig] = mx
    return out, out_min, out_max, warn


def radiative_transfer_lw(planck, od, surf_emissivity, surf_planck):
    """src/ecckd/radiative_transfer_lw.cpp:25-60 -> (flux_dn, flux_up) [nlay+1][n]."""
    nlay, n = od.shape
    with np.errstate(all="ignore"):
        emis = 1.0 - np.exp(-LW_DIFFUSIVITY * od)
        factor = np.where(emis > 1.0e-5, 1.0 - emis * (1.0 / LW_DIFFUSIVITY) / od, 0.5 * emis)
    dn = np.zeros((nlay + 1, n))
    up = np.zeros((nlay + 1, n))
    for i in range(nlay):
        dn[i + 1] = dn[i] * (1.0 - emis[i]) + planck[i] * (emis[i] - factor[i]) + planck[i + 1] * factor[i]
    up[nlay] = surf_planck * surf_emissivity + (1.0 - surf_emissivity) * dn[nlay]
    for i in range(nlay - 1, -1, -1):
        up[i] = up[i + 1] * (1.0 - emis[i]) + planck[i + 1] * (emis[i] - factor[i]) + planck[i] * factor[i]
    return dn, up


def radiative_transfer_direct_sw(cos_sza, ssi, od):
    """src/ecckd/radiative_transfer_sw.cpp:24-43."""
    nlay, n = od.shape
    dn = np.zeros((nlay + 1, n))
    dn[0] = cos_sza * ssi
    for i in range(nlay):
        dn[i + 1] = dn[i] * np.exp((-1.0 / cos_sza) * od[i])
    return dn


def radiative_transfer_norayleigh_sw(cos_sza, ssi, od, albedo):
    """src/ecckd/radiative_transfer_sw.cpp:47-77."""
    nlay, n = od.shape
    dn = radiative_transfer_direct_sw(cos_sza, ssi, od)
    up = np.zeros((nlay + 1, n))
    up[nlay] = dn[nlay] * albedo
    for i in range(nlay - 1, -1, -1):
        up[i] = up[i + 1] * np.exp(-2.0 * od[i])
    return dn, up


def heating_rate(pressure_hl, flux_dn, flux_up=None):
    """src/ecckd/heating_rate.h:25-49."""
    conv = -(ACCEL_GRAVITY / SPECIFIC_HEAT_AIR) / (pressure_hl[1:] - pressure_hl[:-1])
    if flux_up is None:
        return conv[:, None] * (flux_dn[1:] - flux_dn[:-1])
    return conv[:, None] * (flux_dn[1:] - flux_dn[:-1] - flux_up[1:] + flux_up[:-1])


def ckd_calc_optical_depth(conc_dependence, vmr_lut, reference_vmr, log_pressure_lut, temperature_lut, molar_abs,
                           pressure_hl, temperature_fl, vmr_fl):
    """src/ecckd/ckd_model.cpp:923-1102 (linear interpolation) -> [ncol][nlay][ng]."""
    nt, np_lut = temperature_lut.shape
    ng = molar_abs.shape[-1]
    ncol, nhl = pressure_hl.shape
    nlay = nhl - 1
    p0 = pressure_hl[:, :-1]
    p1 = pressure_hl[:, 1:]
    lp = np.log(0.5 * (p1 + p0))
    pidx = (lp - log_pressure_lut[0]) / (log_pressure_lut[1] - log_pressure_lut[0])
    pidx = np.fmax(0.0, np.fmin(pidx, np_lut - 1.0001))
    ip0 = pidx.astype(np.int64)
    pw1 = pidx - ip0
    pw0 = 1.0 - pw1
    t_0 = pw0 * temperature_lut[0][ip0] + pw1 * temperature_lut[0][ip0 + 1]
    d_t = temperature_lut[1, 0] - temperature_lut[0, 0]
    tidx = (temperature_fl - t_0) / d_t
    tidx = np.fmax(0.0, np.fmin(tidx, nt - 1.0001))
    it0 = tidx.astype(np.int64)
    tw1 = tidx - it0
    tw0 = 1.0 - tw1
    sw = (1.0 / (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR)) * (p1 - p0)
    if conc_dependence == 0:
        wgt = sw
    elif conc_dependence == 3:
        wgt = sw * (vmr_fl - reference_vmr)
    else:
        wgt = sw * vmr_fl

    def bil(k):
        return (tw0[..., None] * (pw0[..., None] * k[it0, ip0] + pw1[..., None] * k[it0, ip0 + 1])
                + tw1[..., None] * (pw0[..., None] * k[it0 + 1, ip0] + pw1[..., None] * k[it0 + 1, ip0 + 1]))

    if conc_dependence == 2:
        nconc = len(vmr_lut)
        cidx = (np.log(vmr_fl) - np.log(vmr_lut[0])) / np.log(vmr_lut[1] / vmr_lut[0])
        cidx = np.fmax(0.0, np.fmin(cidx, nconc - 1.0001))
        ic0 = cidx.astype(np.int64)
        cw1 = cidx - ic0
        cw0 = 1.0 - cw1
        out = np.zeros((ncol, nlay, ng))
        for c in range(ncol):
            for l in range(nlay):
                a = molar_abs[ic0[c, l]]
                b = molar_abs[ic0[c, l] + 1]
                i, j = it0[c, l], ip0[c, l]
                fa = tw0[c, l] * (pw0[c, l] * a[i, j] + pw1[c, l] * a[i, j + 1]) + tw1[c, l] * (pw0[c, l] * a[i + 1, j] + pw1[c, l] * a[i + 1, j + 1])
                fb = tw0[c, l] * (pw0[c, l] * b[i, j] + pw1[c, l] * b[i, j + 1]) + tw1[c, l] * (pw0[c, l] * b[i + 1, j] + pw1[c, l] * b[i + 1, j + 1])
                out[c, l] = wgt[c, l] * (cw0[c, l] * fa + cw1[c, l] * fb)
        return out
    return wgt[..., None] * bil(molar_abs)


def ckd_calc_planck_function(temperature_planck, planck_lut, temperature):
    """src/ecckd/ckd_model.cpp:1120-1145 -> temperature.shape + (ng,)."""
    d_t = temperature_planck[1] - temperature_planck[0]
    t0 = temperature_planck[0]
    t = np.asarray(temperature, dtype=np.float64)
    tidx = (t - t0) / d_t
    it0 = np.minimum(np.maximum(tidx, 0.0).astype(np.int64), len(temperature_planck) - 2)
    w1 = tidx - it0
    w0 = 1.0 - w1
    normal = w0[..., None] * planck_lut[it0] + w1[..., None] * planck_lut[it0 + 1]
    low = (t / t0)[..., None] * planck_lut[0]
    return np.where((tidx >= 0)[..., None], normal, low)


def rank_bands(wavenumber_cm_1, key, band_bound1, band_bound2):
    """src/ecckd/reorder_spectrum.cpp:262-300 -> (iband, rank, ordered_index).

    np.argsort(kind='stable') on the key is equivalent to std::stable_sort with
    `key[a] < key[b]` (no NaNs; -0.0 == +0.0 in both).
    """
    nwav = len(key)
    nband = len(band_bound1)
    order = np.arange(nwav, dtype=np.int64)
    iband = np.full(nwav, -1, dtype=np.int32)
    for jb in range(nband):
        if jb < nband - 1:
            m = (wavenumber_cm_1 >= band_bound1[jb]) & (wavenumber_cm_1 < band_bound2[jb])
        else:
            m = (wavenumber_cm_1 >= band_bound1[jb]) & (wavenumber_cm_1 <= band_bound2[jb])
        idx = np.nonzero(m)[0]
        iband[idx] = jb
        i1, i2 = idx[0], idx[-1]
        seg = order[i1:i2 + 1]
        order[i1:i2 + 1] = seg[np.argsort(key[seg], kind="stable")]
    rank = np.empty(nwav, dtype=np.int32)
    rank[order] = np.arange(nwav, dtype=np.int32)
    return iband, rank, order.astype(np.int32)


def store_g_points(rank, rank1, rank2):
    """src/ecckd/single_gas_data.h:56-62."""
    g = np.full(len(rank), -1, dtype=np.int32)
    for ig in range(len(rank1)):
        g[(rank >= rank1[ig]) & (rank <= rank2[ig])] = ig
    return g


def assign_g_points(gas_g_point, g_min, g_max):
    """src/ecckd/find_g_points.cpp:1459-1475; gas_g_point[ngas][nwav], g_min/g_max[ngas][ng]."""
    ngas, nwav = gas_g_point.shape
    ng = g_min.shape[1]
    g = np.full(nwav, -1, dtype=np.int32)
    for ig in range(ng):
        found = np.ones(nwav, dtype=bool)
        for igas in range(ngas):
            found &= ~((gas_g_point[igas] < g_min[igas, ig]) | (gas_g_point[igas] > g_max[igas, ig]))
        g[found] = ig
    return g


def gpoint_fraction(ng, g_point, wavenumber_cm_1, d_wavenumber_cm_1, wavenumber1, wavenumber2):
    """src/ecckd/create_look_up_table.cpp:537-548 -> [ng][nbin]."""
    out = np.zeros((ng, len(wavenumber1)))
    with np.errstate(all="ignore"):
        for ig in range(ng):
            m = g_point == ig
            tot = np.sum(d_wavenumber_cm_1[m])
            for ib in range(len(wavenumber1)):
                mm = m & (wavenumber_cm_1 > wavenumber1[ib]) & (wavenumber_cm_1 <= wavenumber2[ib])
                out[ig, ib] = np.sum(d_wavenumber_cm_1[mm]) / tot
    return out


def planck_lut(ng, g_point, wavenumber_cm_1, d_wavenumber_cm_1, temperature_lut):
    """src/ecckd/create_look_up_table.cpp:579-591 -> [nlut][ng]."""
    out = np.zeros((len(temperature_lut), ng))
    for ig in range(ng):
        m = g_point == ig
        out[:, ig] = planck_function(temperature_lut, wavenumber_cm_1[m], d_wavenumber_cm_1[m]).sum(axis=1)
    return out
