"""ctypes binding of the CPU oracle (oracle/libecckd_oracle.so).

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libecckd_oracle.so")
# the reference's own translation units, compiled in place against oracle/adept_shim (oracle/build_ref_shim.sh)
REFERENCE_LIB = os.path.join(ORACLE_DIR, "_ref", "libecckd_ref.so")

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


METHODS_FIT = METHODS + ("total-transmission",)


class OrcGasTable(C.Structure):
    _fields_ = [
        ("conc_dependence", C.c_int32),
        ("nconc", C.c_int32),
        ("vmr_lut", C.c_void_p),
        ("reference_vmr", C.c_double),
        ("molar_abs", C.c_void_p),
    ]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data


class Oracle:
    """`Oracle()` binds the restated oracle; `Oracle(REFERENCE_LIB)` binds the same entry points
    backed by the reference's own compiled sources (only the subset listed in oracle/ref_glue.cpp)."""

    def __init__(self, lib=None):
        if lib is None:
            srcs = [os.path.join(ORACLE_DIR, "ecckd_oracle.cpp")]
            stale = (not os.path.exists(ORACLE_LIB)) or any(
                os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in srcs)
            if stale:
                env = dict(os.environ)
                env.pop("CXX", None)
                env.pop("CC", None)
                subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, env=env, stdout=subprocess.DEVNULL)
            lib = ORACLE_LIB
        self.path = lib
        self.L = C.CDLL(lib)

    def has(self, name):
        return hasattr(self.L, "orc_" + name)

    # -- stage 1
    def planck_function(self, temperature, wavenumber, d_wavenumber):
        t, wn, dwn = _f64(np.atleast_1d(temperature)), _f64(wavenumber), _f64(d_wavenumber)
        out = np.empty((t.size, wn.size))
        rc = self.L.orc_planck_function(C.c_int(t.size), C.c_void_p(_p(t)), C.c_int64(wn.size), C.c_void_p(_p(wn)),
                                        C.c_void_p(_p(dwn)), C.c_void_p(_p(out)), C.c_int64(wn.size))
        assert rc == 0
        return out

    def average_optical_depth_to_g_point(self, ng, vmr, pressure_fl, pressure_hl, g_point, od, weight, method):
        od = _f64(od)
        nlay, nwav = od.shape
        w = _f64(weight)
        w_stride = 0 if w.ndim == 1 else nwav
        gp = _i32(g_point)
        pf, ph = _f64(pressure_fl), _f64(pressure_hl)
        out, omin, omax = np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng))
        warn = np.zeros(ng, dtype=np.int32)
        mid = METHODS.index(method) if isinstance(method, str) else int(method)
        rc = self.L.orc_average_optical_depth_to_g_point(
            C.c_int(ng), C.c_double(vmr), C.c_int(nlay), C.c_int64(nwav), C.c_void_p(_p(pf)), C.c_void_p(_p(ph)),
            C.c_void_p(_p(gp)), C.c_void_p(_p(od)), C.c_int64(nwav), C.c_void_p(_p(w)), C.c_int64(w_stride),
            C.c_int(mid), C.c_void_p(_p(out)), C.c_void_p(_p(omin)), C.c_void_p(_p(omax)), C.c_int64(ng),
            C.c_void_p(_p(warn)))
        if rc != 0:
            raise RuntimeError(rc)
        return out, omin, omax, warn

    def merge_axpy(self, optical_depth, od, scaling=1.0, scaling_profile=None):
        dst = _f64(optical_depth).copy()
        src = _f64(od)
        nlay, nwav = dst.shape
        sp = None if scaling_profile is None else _f64(scaling_profile)
        self.L.orc_merge_axpy(C.c_int(nlay), C.c_int64(nwav), C.c_void_p(_p(dst)), C.c_int64(nwav),
                              C.c_void_p(_p(src)), C.c_int64(nwav), C.c_double(scaling), C.c_void_p(_p(sp)))
        return dst

    def gpoint_fraction_grid(self, do_sw, band_wn1, band_wn2):
        b1, b2 = _f64(band_wn1), _f64(band_wn2)
        n = self.L.orc_gpoint_fraction_grid(C.c_int(int(do_sw)), C.c_int(b1.size), C.c_void_p(_p(b1)),
                                            C.c_void_p(_p(b2)), None, None)
        w1, w2 = np.zeros(n), np.zeros(n)
        self.L.orc_gpoint_fraction_grid(C.c_int(int(do_sw)), C.c_int(b1.size), C.c_void_p(_p(b1)), C.c_void_p(_p(b2)),
                                        C.c_void_p(_p(w1)), C.c_void_p(_p(w2)))
        return w1, w2

    def gpoint_fraction(self, ng, g_point, wavenumber, d_wavenumber, w1, w2):
        gp, wn, dwn, w1, w2 = _i32(g_point), _f64(wavenumber), _f64(d_wavenumber), _f64(w1), _f64(w2)
        out = np.zeros((ng, w1.size))
        self.L.orc_gpoint_fraction(C.c_int(ng), C.c_int64(wn.size), C.c_void_p(_p(gp)), C.c_void_p(_p(wn)),
                                   C.c_void_p(_p(dwn)), C.c_int(w1.size), C.c_void_p(_p(w1)), C.c_void_p(_p(w2)),
                                   C.c_void_p(_p(out)))
        return out

    def ssi_intervals(self, wavenumber, ssi, w1, w2):
        wn, s, w1, w2 = _f64(wavenumber), _f64(ssi), _f64(w1), _f64(w2)
        out = np.zeros(w1.size)
        self.L.orc_ssi_intervals(C.c_int64(wn.size), C.c_void_p(_p(wn)), C.c_void_p(_p(s)), C.c_int(w1.size),
                                 C.c_void_p(_p(w1)), C.c_void_p(_p(w2)), C.c_void_p(_p(out)))
        return out

    def planck_lut(self, ng, g_point, wavenumber, d_wavenumber, temperature_lut):
        gp, wn, dwn, t = _i32(g_point), _f64(wavenumber), _f64(d_wavenumber), _f64(temperature_lut)
        out = np.zeros((t.size, ng))
        self.L.orc_planck_lut(C.c_int(ng), C.c_int64(wn.size), C.c_void_p(_p(gp)), C.c_void_p(_p(wn)),
                              C.c_void_p(_p(dwn)), C.c_int(t.size), C.c_void_p(_p(t)), C.c_void_p(_p(out)))
        return out

    def compact_g_points(self, ng, g_point):
        gp = _i32(g_point).copy()
        gmap = np.zeros(ng, dtype=np.int32)
        n = self.L.orc_compact_g_points(C.c_int(ng), C.c_int64(gp.size), C.c_void_p(_p(gp)), C.c_void_p(_p(gmap)))
        return n, gp, gmap[: max(n, 0)]

    # -- radiative transfer
    def radiative_transfer_lw(self, planck, od, surf_emissivity, surf_planck):
        pl, od, se, sp = _f64(planck), _f64(od), _f64(surf_emissivity), _f64(surf_planck)
        nlay, n = od.shape
        dn, up = np.zeros((nlay + 1, n)), np.zeros((nlay + 1, n))
        self.L.orc_radiative_transfer_lw(C.c_int(nlay), C.c_int64(n), C.c_void_p(_p(pl)), C.c_void_p(_p(od)),
                                         C.c_void_p(_p(se)), C.c_void_p(_p(sp)), C.c_void_p(_p(dn)),
                                         C.c_void_p(_p(up)))
        return dn, up

    def radiative_transfer_lw_bb(self, planck, spectral_od, grey_od, surf_emissivity, surf_planck):
        pl, od, go, se, sp = (_f64(a) for a in (planck, spectral_od, grey_od, surf_emissivity, surf_planck))
        nlay, n = od.shape
        dn, up = np.zeros(nlay + 1), np.zeros(nlay + 1)
        self.L.orc_radiative_transfer_lw_bb(C.c_int(nlay), C.c_int64(n), C.c_void_p(_p(pl)), C.c_void_p(_p(od)),
                                            C.c_void_p(_p(go)), C.c_void_p(_p(se)), C.c_void_p(_p(sp)),
                                            C.c_void_p(_p(dn)), C.c_void_p(_p(up)))
        return dn, up

    def radiative_transfer_direct_sw(self, cos_sza, ssi, od):
        s, od = _f64(ssi), _f64(od)
        nlay, n = od.shape
        dn = np.zeros((nlay + 1, n))
        self.L.orc_radiative_transfer_direct_sw(C.c_int(nlay), C.c_int64(n), C.c_double(cos_sza), C.c_void_p(_p(s)),
                                                C.c_void_p(_p(od)), C.c_void_p(_p(dn)))
        return dn

    def radiative_transfer_norayleigh_sw(self, cos_sza, ssi, od, albedo):
        s, od, al = _f64(ssi), _f64(od), _f64(albedo)
        nlay, n = od.shape
        dn, up = np.zeros((nlay + 1, n)), np.zeros((nlay + 1, n))
        self.L.orc_radiative_transfer_norayleigh_sw(C.c_int(nlay), C.c_int64(n), C.c_double(cos_sza),
                                                    C.c_void_p(_p(s)), C.c_void_p(_p(od)), C.c_void_p(_p(al)),
                                                    C.c_void_p(_p(dn)), C.c_void_p(_p(up)))
        return dn, up

    def heating_rate(self, pressure_hl, flux_dn, flux_up=None):
        p, dn = _f64(pressure_hl), _f64(flux_dn)
        up = None if flux_up is None else _f64(flux_up)
        nhl, n = dn.shape
        hr = np.zeros((nhl - 1, n))
        self.L.orc_heating_rate(C.c_int(nhl - 1), C.c_int64(n), C.c_void_p(_p(p)), C.c_void_p(_p(dn)),
                                C.c_void_p(_p(up)), C.c_void_p(_p(hr)))
        return hr

    def radiative_transfer_sw_bb(self, cos_sza, ssi, spectral_od, grey_od, albedo, direct_only=False):
        s, od, go = _f64(ssi), _f64(spectral_od), _f64(grey_od)
        nlay, n = od.shape
        dn = np.zeros(nlay + 1)
        up = None if direct_only else np.zeros(nlay + 1)
        rc = self.L.orc_radiative_transfer_sw_bb(C.c_int(nlay), C.c_int64(n), C.c_double(cos_sza), C.c_void_p(_p(s)),
                                                 C.c_void_p(_p(od)), C.c_void_p(_p(go)), C.c_double(albedo),
                                                 C.c_void_p(_p(dn)), C.c_void_p(_p(up)))
        assert rc == 0, rc
        return dn, up

    def heating_rate_single(self, pressure_hl, flux_dn, flux_up=None):
        p, dn = _f64(pressure_hl), _f64(flux_dn)
        up = None if flux_up is None else _f64(flux_up)
        hr = np.zeros(dn.size - 1)
        rc = self.L.orc_heating_rate_single(C.c_int(dn.size - 1), C.c_void_p(_p(p)), C.c_void_p(_p(dn)),
                                            C.c_void_p(_p(up)), C.c_void_p(_p(hr)))
        assert rc == 0, rc
        return hr

    def rayleigh_molar_scattering_coeff(self, wavenumber):
        wn = _f64(wavenumber)
        out = np.zeros(wn.size)
        rc = self.L.orc_rayleigh_molar_scattering_coeff(C.c_int64(wn.size), C.c_void_p(_p(wn)), C.c_void_p(_p(out)))
        assert rc == 0, rc
        return out

    def rayleigh_molar_scat_gpoints(self, wavenumber1, wavenumber2, gpoint_fraction, ssi_intervals):
        w1, w2, gf, si = _f64(wavenumber1), _f64(wavenumber2), _f64(gpoint_fraction), _f64(ssi_intervals)
        ng, nbin = gf.shape
        out = np.zeros(ng)
        rc = self.L.orc_rayleigh_molar_scat_gpoints(C.c_int(ng), C.c_int(nbin), C.c_void_p(_p(w1)), C.c_void_p(_p(w2)),
                                                    C.c_void_p(_p(gf)), C.c_void_p(_p(si)), C.c_void_p(_p(out)))
        assert rc == 0, rc
        return out

    def split_base_g_points(self, band_wn1, band_wn2, boundaries, wavenumber, ssi, ng, band_number, g_point,
                            solar_irradiance=None):
        b1, b2, bb, wn = _f64(band_wn1), _f64(band_wn2), _f64(np.atleast_1d(boundaries)), _f64(wavenumber)
        s = None if ssi is None else _f64(ssi)
        cap = int(ng) + bb.size
        bn = np.zeros(cap, dtype=np.int32)
        bn[:ng] = band_number
        g = _i32(g_point).copy()
        sol = None
        if solar_irradiance is not None:
            sol = np.zeros(cap)
            sol[:ng] = solar_irradiance
        ng_io = C.c_int(int(ng))
        rc = self.L.orc_split_base_g_points(C.c_int(b1.size), C.c_void_p(_p(b1)), C.c_void_p(_p(b2)), C.c_int(bb.size),
                                            C.c_void_p(_p(bb)), C.c_int64(wn.size), C.c_void_p(_p(wn)),
                                            C.c_void_p(_p(s)), C.c_int(cap), C.byref(ng_io), C.c_void_p(_p(bn)),
                                            C.c_void_p(_p(g)), C.c_void_p(_p(sol)))
        if rc != 0:
            raise RuntimeError(rc)
        n = ng_io.value
        return n, bn[:n].copy(), g, (None if sol is None else sol[:n].copy())

    # -- interval fits of find_g_points (find_g_points.cpp:54-204)
    def fit_optical_depth_lw(self, method, i1, i2, planck_hl, metric):
        pl, m = _f64(planck_hl), _f64(metric)
        nlay, n = m.shape
        out = np.zeros(nlay)
        mid = METHODS_FIT.index(method) if isinstance(method, str) else int(method)
        rc = self.L.orc_fit_optical_depth_lw(C.c_int(mid), C.c_int(nlay), C.c_int64(n), C.c_int64(i1), C.c_int64(i2),
                                             C.c_void_p(_p(pl)), C.c_void_p(_p(m)), C.c_void_p(_p(out)))
        if rc != 0:
            raise RuntimeError(rc)
        return out

    def fit_optical_depth_sw(self, method, i1, i2, ssi, metric):
        s, m = _f64(ssi), _f64(metric)
        nlay, n = m.shape
        out = np.zeros(nlay)
        mid = METHODS_FIT.index(method) if isinstance(method, str) else int(method)
        rc = self.L.orc_fit_optical_depth_sw(C.c_int(mid), C.c_int(nlay), C.c_int64(n), C.c_int64(i1), C.c_int64(i2),
                                             C.c_void_p(_p(s)), C.c_void_p(_p(m)), C.c_void_p(_p(out)))
        if rc != 0:
            raise RuntimeError(rc)
        return out

    def fit_optical_depth_sw_total_trans(self, i1, i2, ssi, bg_od, od):
        s, b, o = _f64(ssi), _f64(bg_od), _f64(od)
        nlay, n = o.shape
        out = np.zeros(nlay)
        rc = self.L.orc_fit_optical_depth_sw_total_trans(C.c_int(nlay), C.c_int64(n), C.c_int64(i1), C.c_int64(i2),
                                                         C.c_void_p(_p(s)), C.c_void_p(_p(b)), C.c_void_p(_p(o)),
                                                         C.c_void_p(_p(out)))
        assert rc == 0, rc
        return out

    # -- cost functions (interval error a13, forward cost a22)
    def calc_cost_function_lw(self, pressure_hl, planck_hl, surf_emissivity, surf_planck, bg_od, od_fit, flux_dn_surf,
                              flux_up_toa, hr, flux_weight, layer_weight):
        ph, pl, se, sp, bg, fit, fds, fut, hr, lw = (_f64(a) for a in (
            pressure_hl, planck_hl, surf_emissivity, surf_planck, bg_od, od_fit, flux_dn_surf, flux_up_toa, hr,
            layer_weight))
        nlay, n = bg.shape
        cost = C.c_double(0.0)
        rc = self.L.orc_calc_cost_function_lw(
            C.c_int(nlay), C.c_int64(n), C.c_void_p(_p(ph)), C.c_void_p(_p(pl)), C.c_void_p(_p(se)), C.c_void_p(_p(sp)),
            C.c_void_p(_p(bg)), C.c_void_p(_p(fit)), C.c_void_p(_p(fds)), C.c_void_p(_p(fut)), C.c_void_p(_p(hr)),
            C.c_double(flux_weight), C.c_void_p(_p(lw)), C.byref(cost))
        assert rc == 0, rc
        return cost.value

    def calc_cost_function_sw(self, cos_sza, pressure_hl, ssi, albedo, bg_od, od_fit, flux_dn_surf, flux_up_toa, hr,
                              flux_weight, layer_weight):
        ph, s, bg, fit, fds, fut, hr, lw = (_f64(a) for a in (
            pressure_hl, ssi, bg_od, od_fit, flux_dn_surf, flux_up_toa, hr, layer_weight))
        nlay, n = bg.shape
        cost = C.c_double(0.0)
        rc = self.L.orc_calc_cost_function_sw(
            C.c_int(nlay), C.c_int64(n), C.c_double(cos_sza), C.c_void_p(_p(ph)), C.c_void_p(_p(s)), C.c_double(albedo),
            C.c_void_p(_p(bg)), C.c_void_p(_p(fit)), C.c_void_p(_p(fds)), C.c_void_p(_p(fut)), C.c_void_p(_p(hr)),
            C.c_double(flux_weight), C.c_void_p(_p(lw)), C.byref(cost))
        assert rc == 0, rc
        return cost.value

    def calc_cost_function_ckd_lw(self, pressure_hl, planck_hl, surf_emissivity, surf_planck, od, flux_dn, flux_up, hr,
                                  flux_weight, flux_profile_weight, broadband_weight, spectral_boundary_weight,
                                  layer_weight, spectral_flux_dn_surf=None, spectral_flux_up_toa=None, rel_dn=None,
                                  rel_up=None, band_mapping=None):
        ph, pl, se, sp, od, fd, fu, hr, lw = (_f64(a) for a in (
            pressure_hl, planck_hl, surf_emissivity, surf_planck, od, flux_dn, flux_up, hr, layer_weight))
        nlay, ng = od.shape
        nband = fd.shape[1]
        opt = [None if a is None else _f64(a) for a in (spectral_flux_dn_surf, spectral_flux_up_toa, rel_dn, rel_up)]
        bm = None if band_mapping is None else _i32(band_mapping)
        cost = C.c_double(0.0)
        rc = self.L.orc_calc_cost_function_ckd_lw(
            C.c_int(nlay), C.c_int(ng), C.c_int(nband), C.c_void_p(_p(ph)), C.c_void_p(_p(pl)), C.c_void_p(_p(se)),
            C.c_void_p(_p(sp)), C.c_void_p(_p(od)), C.c_void_p(_p(fd)), C.c_void_p(_p(fu)), C.c_void_p(_p(hr)),
            C.c_void_p(_p(opt[0])), C.c_void_p(_p(opt[1])), C.c_double(flux_weight), C.c_double(flux_profile_weight),
            C.c_double(broadband_weight), C.c_double(spectral_boundary_weight), C.c_void_p(_p(lw)),
            C.c_void_p(_p(opt[2])), C.c_void_p(_p(opt[3])), C.c_void_p(_p(bm)), C.byref(cost))
        if rc != 0:
            raise RuntimeError(rc)
        return cost.value

    def calc_cost_function_ckd_sw(self, cos_sza, pressure_hl, ssi, albedo, od, flux_dn, flux_up, hr, flux_weight,
                                  flux_profile_weight, broadband_weight, layer_weight, spectral_boundary_weights=None,
                                  spectral_flux_dn_surf=None, spectral_flux_up_toa=None, rel_dn=None, rel_up=None,
                                  band_mapping=None, cost_fn_per_band=None):
        ph, s, al, od, fd, fu, hr, lw = (_f64(a) for a in (
            pressure_hl, ssi, np.atleast_1d(albedo), od, flux_dn, flux_up, hr, layer_weight))
        nlay, ng = od.shape
        nband = fd.shape[1]
        opt = [None if a is None else _f64(a) for a in (
            spectral_flux_dn_surf, spectral_flux_up_toa, spectral_boundary_weights, rel_dn, rel_up)]
        bm = None if band_mapping is None else _i32(band_mapping)
        cost = C.c_double(0.0)
        if cost_fn_per_band is not None:
            assert cost_fn_per_band.dtype == np.float64 and cost_fn_per_band.flags.c_contiguous
        rc = self.L.orc_calc_cost_function_ckd_sw(
            C.c_int(nlay), C.c_int(ng), C.c_int(nband), C.c_double(cos_sza), C.c_void_p(_p(ph)), C.c_void_p(_p(s)),
            C.c_int(al.size), C.c_void_p(_p(al)), C.c_void_p(_p(od)), C.c_void_p(_p(fd)), C.c_void_p(_p(fu)),
            C.c_void_p(_p(hr)), C.c_void_p(_p(opt[0])), C.c_void_p(_p(opt[1])), C.c_double(flux_weight),
            C.c_double(flux_profile_weight), C.c_double(broadband_weight), C.c_void_p(_p(opt[2])), C.c_void_p(_p(lw)),
            C.c_void_p(_p(opt[3])), C.c_void_p(_p(opt[4])), C.c_void_p(_p(bm)), C.c_void_p(_p(cost_fn_per_band)),
            C.byref(cost))
        if rc != 0:
            raise RuntimeError(rc)
        return cost.value

    # -- CKD model
    def ckd_calc_optical_depth(self, gas, log_pressure, temperature, pressure_hl, temperature_fl, vmr_fl):
        k = _f64(gas["molar_abs"])
        v = None if gas.get("vmr_lut") is None else _f64(gas["vmr_lut"])
        lp, t = _f64(log_pressure), _f64(temperature)
        ph, tf = _f64(pressure_hl), _f64(temperature_fl)
        vm = None if vmr_fl is None else _f64(vmr_fl)
        ncol, nhl = ph.shape
        nt, npl = t.shape
        ng = k.shape[-1]
        od = np.zeros((ncol, nhl - 1, ng))
        rc = self.L.orc_ckd_calc_optical_depth(
            C.c_int(int(gas["conc_dependence"])), C.c_int(0 if v is None else v.size), C.c_void_p(_p(v)),
            C.c_double(float(gas.get("reference_vmr", 0.0))), C.c_int(nt), C.c_int(npl), C.c_int(ng),
            C.c_void_p(_p(lp)), C.c_void_p(_p(t)), C.c_void_p(_p(k)), C.c_int(ncol), C.c_int(nhl - 1),
            C.c_void_p(_p(ph)), C.c_void_p(_p(tf)), C.c_void_p(_p(vm)), C.c_void_p(_p(od)))
        if rc != 0:
            raise RuntimeError(rc)
        return od

    def background_cost_gas(self, inv_background_shape, background_error, delta_x, gradient):
        """ckd_model.cpp:838-866 for one gas: returns the cost, adds the gas's term to `gradient` [nx][ng] in place."""
        sh, err, dx = _f64(inv_background_shape), _f64(background_error), _f64(delta_x)
        nx, ng = dx.shape
        assert sh.shape == (nx, nx) and err.shape == (ng,)
        assert gradient.dtype == np.float64 and gradient.flags.c_contiguous and gradient.shape == (nx, ng)
        cost = C.c_double(0.0)
        rc = self.L.orc_background_cost_gas(C.c_int(nx), C.c_int(ng), C.c_void_p(_p(sh)), C.c_void_p(_p(err)),
                                            C.c_void_p(_p(dx)), C.c_void_p(_p(gradient)), C.byref(cost))
        if rc != 0:
            raise RuntimeError(rc)
        return cost.value

    def background_cost_rayleigh(self, rayleigh_inv_background, delta_x):
        ib, dx = _f64(rayleigh_inv_background), _f64(delta_x)
        cost = C.c_double(0.0)
        rc = self.L.orc_background_cost_rayleigh(C.c_int(dx.size), C.c_void_p(_p(ib)), C.c_void_p(_p(dx)), C.byref(cost))
        if rc != 0:
            raise RuntimeError(rc)
        return cost.value

    def ckd_calc_planck_function(self, temperature_planck, planck_lut, temperature):
        tp, pl, t = _f64(temperature_planck), _f64(planck_lut), _f64(temperature)
        out = np.zeros(t.shape + (pl.shape[1],))
        self.L.orc_ckd_calc_planck_function(C.c_int(tp.size), C.c_int(pl.shape[1]), C.c_void_p(_p(tp)),
                                            C.c_void_p(_p(pl)), C.c_int64(t.size), C.c_void_p(_p(t)),
                                            C.c_void_p(_p(out)))
        return out

    def temperature_fl(self, temperature_hl, pressure_hl):
        t, p = _f64(temperature_hl), _f64(pressure_hl)
        ncol, nhl = t.shape
        out = np.zeros((ncol, nhl - 1))
        self.L.orc_temperature_fl(C.c_int(ncol), C.c_int(nhl - 1), C.c_void_p(_p(t)), C.c_void_p(_p(p)),
                                  C.c_void_p(_p(out)))
        return out

    def run_ckd(self, case, is_sw=False, tsi=1361.0):
        """case: dict from tests.synthetic.ckd_model_case."""
        gases = case["gases"]
        keep = []
        tabs = (OrcGasTable * len(gases))()
        for i, g in enumerate(gases):
            k = _f64(g["molar_abs"])
            v = None if g.get("vmr_lut") is None else _f64(g["vmr_lut"])
            keep += [k, v]
            tabs[i].conc_dependence = int(g["conc_dependence"])
            tabs[i].nconc = 0 if v is None else v.size
            tabs[i].vmr_lut = _p(v)
            tabs[i].reference_vmr = float(g.get("reference_vmr", 0.0))
            tabs[i].molar_abs = _p(k)
        vm = [None if v is None else _f64(v) for v in case["vmr_fl"]]
        vptrs = (C.c_void_p * len(vm))(*[_p(v) for v in vm])
        lp, t = _f64(case["log_pressure"]), _f64(case["temperature"])
        ph, th = _f64(case["pressure_hl"]), _f64(case["temperature_hl"])
        ncol, nhl = ph.shape
        nlay = nhl - 1
        nt, npl = t.shape
        ng = keep[0].shape[-1]
        tp = pl = si = ra = None
        if not is_sw:
            tp, pl = _f64(case["temperature_planck"]), _f64(case["planck_lut"])
        else:
            si, ra = _f64(case["solar_irradiance"]), _f64(case["rayleigh_molar_scat"])
        out = {"optical_depth": np.zeros((ncol, nlay, ng)), "flux_dn": np.zeros((ncol, nhl, ng)),
               "bb_flux_dn": np.zeros((ncol, nhl)), "planck_hl": np.zeros((ncol, nhl, ng)),
               "planck_surf": np.zeros((ncol, ng)), "flux_up": np.zeros((ncol, nhl, ng)),
               "bb_flux_up": np.zeros((ncol, nhl))}
        rc = self.L.orc_run_ckd(
            C.c_int(int(is_sw)), C.c_int(len(gases)), tabs, vptrs, C.c_int(nt), C.c_int(npl), C.c_int(ng),
            C.c_void_p(_p(lp)), C.c_void_p(_p(t)), C.c_int(0 if tp is None else tp.size), C.c_void_p(_p(tp)),
            C.c_void_p(_p(pl)), C.c_void_p(_p(si)), C.c_void_p(_p(ra)), C.c_double(tsi), C.c_int(ncol), C.c_int(nlay),
            C.c_void_p(_p(ph)), C.c_void_p(_p(th)), C.c_void_p(_p(out["optical_depth"])),
            C.c_void_p(_p(out["planck_hl"])), C.c_void_p(_p(out["planck_surf"])), C.c_void_p(_p(out["flux_dn"])),
            C.c_void_p(_p(out["flux_up"])), C.c_void_p(_p(out["bb_flux_dn"])), C.c_void_p(_p(out["bb_flux_up"])))
        if rc != 0:
            raise RuntimeError(rc)
        return out

    # -- ranking
    def rank_bands(self, wavenumber, key, b1, b2):
        wn, k, b1, b2 = _f64(wavenumber), _f64(key), _f64(np.atleast_1d(b1)), _f64(np.atleast_1d(b2))
        n = wn.size
        iband, rank, order = (np.zeros(n, dtype=np.int32) for _ in range(3))
        rc = self.L.orc_rank_bands(C.c_int64(n), C.c_void_p(_p(wn)), C.c_void_p(_p(k)), C.c_int(b1.size),
                                   C.c_void_p(_p(b1)), C.c_void_p(_p(b2)), C.c_void_p(_p(iband)),
                                   C.c_void_p(_p(rank)), C.c_void_p(_p(order)))
        if rc != 0:
            raise RuntimeError(rc)
        return iband, rank, order

    def sort_key_lw(self, pressure_hl, wavenumber, d_wavenumber, od, threshold=0.5):
        p, wn, dwn, od = _f64(pressure_hl), _f64(wavenumber), _f64(d_wavenumber), _f64(od)
        nlay, n = od.shape
        key, col = np.zeros(n), np.zeros(n)
        self.L.orc_sort_key_lw(C.c_int(nlay), C.c_int64(n), C.c_void_p(_p(p)), C.c_void_p(_p(wn)),
                               C.c_void_p(_p(dwn)), C.c_void_p(_p(od)), C.c_double(threshold), C.c_void_p(_p(key)),
                               C.c_void_p(_p(col)))
        return key, col

    def sort_key_sw(self, pressure_hl, od, threshold=0.25):
        p, od = _f64(pressure_hl), _f64(od)
        nlay, n = od.shape
        key, col = np.zeros(n), np.zeros(n)
        rc = self.L.orc_sort_key_sw(C.c_int(nlay), C.c_int64(n), C.c_void_p(_p(p)), C.c_void_p(_p(od)),
                                    C.c_double(threshold), C.c_void_p(_p(key)), C.c_void_p(_p(col)))
        return key, col, rc

    def store_g_points(self, rank, rank1, rank2):
        r, r1, r2 = _i32(rank), _i32(rank1), _i32(rank2)
        g = np.zeros(r.size, dtype=np.int32)
        self.L.orc_store_g_points(C.c_int64(r.size), C.c_void_p(_p(r)), C.c_int(r1.size), C.c_void_p(_p(r1)),
                                  C.c_void_p(_p(r2)), C.c_void_p(_p(g)))
        return g

    def overlap_g_points(self, n_g_points, sorting_variables):
        n_g = _i32(n_g_points)
        ngas, nband = n_g.shape
        sv = _f64(np.concatenate([np.asarray(s, dtype=np.float64) for s in sorting_variables]))
        offs = np.zeros(ngas, dtype=np.int32)
        offs[1:] = np.cumsum([len(s) for s in sorting_variables])[:-1]
        cap = int(n_g.sum()) + 1
        gmin, gmax = np.zeros((ngas, cap), dtype=np.int32), np.zeros((ngas, cap), dtype=np.int32)
        band = np.zeros(cap, dtype=np.int32)
        ng = self.L.orc_overlap_g_points(C.c_int(ngas), C.c_int(nband), C.c_void_p(_p(n_g)), C.c_void_p(_p(sv)),
                                         C.c_void_p(_p(offs)), C.c_int(cap), C.c_void_p(_p(gmin)),
                                         C.c_void_p(_p(gmax)), C.c_void_p(_p(band)))
        assert ng >= 0, ng
        return ng, gmin[:, :ng].copy(), gmax[:, :ng].copy(), band[:ng].copy()

    def assign_g_points(self, gas_g_point, g_min, g_max):
        gg, gmin, gmax = _i32(gas_g_point), _i32(g_min), _i32(g_max)
        ngas, nwav = gg.shape
        ng = gmin.shape[1]
        g = np.zeros(nwav, dtype=np.int32)
        n_un = C.c_int64(0)
        self.L.orc_assign_g_points(C.c_int(ngas), C.c_int64(nwav), C.c_void_p(_p(gg)), C.c_int(ng), C.c_int(ng),
                                   C.c_void_p(_p(gmin)), C.c_void_p(_p(gmax)), C.c_void_p(_p(g)), C.byref(n_un))
        return g, int(n_un.value)


_reference = None


def reference_oracle():
    """The reference's own sources behind the oracle's entry points, or None when oracle/_ref has not
    been built (it is built by __graft_entry__.build() wherever /root/reference exists and travels
    to the GPU box as a prebuilt file)."""
    global _reference
    if _reference is None and os.path.exists(REFERENCE_LIB):
        _reference = Oracle(REFERENCE_LIB)
    return _reference


_oracle = None


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle
