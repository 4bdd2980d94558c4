"""ctypes harness over oracle/_ref/libecckd_ref_tools.so: the reference's OWN tools (run_ckd,
create_look_up_table, reorder_spectrum, find_g_points internals) compiled in place against the Adept
shim and the in-memory file layer (oracle/build_ref_tools.sh).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("REF_TOOLS_LIB", os.path.join(ROOT, "oracle", "_ref", "libecckd_ref_tools.so"))


GPU_LIB = os.path.join(ROOT, "oracle", "_ref", "libecckd_ref_tools_gpu.so")  # the same tools on the C ABI


def available():
    return os.path.exists(LIB)


def gpu_available():
    return os.path.exists(GPU_LIB)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class BandLw(C.Structure):
    _fields_ = [("method", C.c_char_p), ("flux_weight", C.c_double), ("nlay", C.c_int), ("n", C.c_int64),
                ("ibegin", C.c_int64), ("iend", C.c_int64)] + [(k, C.c_void_p) for k in (
                    "layer_weight", "pressure_hl", "surf_emissivity", "surf_planck", "flux_dn_surf", "flux_up_toa",
                    "planck_hl", "bg_od", "metric", "hr")]


class BandSw(C.Structure):
    _fields_ = [("method", C.c_char_p), ("flux_weight", C.c_double), ("cos_sza", C.c_double), ("albedo", C.c_double),
                ("nlay", C.c_int), ("n", C.c_int64), ("ibegin", C.c_int64), ("iend", C.c_int64)] + [
                    (k, C.c_void_p) for k in ("layer_weight", "pressure_hl", "ssi", "flux_dn_surf", "flux_up_toa",
                                              "bg_od", "metric", "hr", "flux_dn_surf_low", "flux_up_toa_low",
                                              "flux_dn_surf_high", "flux_up_toa_high", "hr_low", "hr_high")] + [
                    ("min_scaling", C.c_double), ("max_scaling", C.c_double)]


class RefEpOptions(C.Structure):
    _fields_ = [("partition_tolerance", C.c_double), ("resolution", C.c_double), ("partition_max_iterations", C.c_int),
                ("line_search_max_iterations", C.c_int), ("minimize_frac_range", C.c_int),
                ("cubic_interpolation", C.c_int)]


REF_ERROR_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double, C.c_double)


class RefTools:
    def __init__(self, lib=None):
        self.L = C.CDLL(lib or LIB)
        self.L.ref_tools_build_info.restype = C.c_char_p
        self.L.ref_tool_last_exception.restype = C.c_char_p
        self._keep = []

    # ---- mock files
    def reset(self):
        self.L.ref_mock_reset()

    def put(self, file, var, array):
        a = _f64(array)
        dims = (C.c_int * max(a.ndim, 1))(*a.shape)
        self.L.ref_mock_put(file.encode(), var.encode(), C.c_int(a.ndim), dims, _p(a))

    def put_string(self, file, name, value, scope=""):
        self.L.ref_mock_put_string(file.encode(), scope.encode(), name.encode(), value.encode())

    def put_number(self, file, name, value, scope=""):
        self.L.ref_mock_put_number(file.encode(), scope.encode(), name.encode(), C.c_double(value))

    def get(self, file, var):
        dims = (C.c_int * 4)()
        nd = self.L.ref_mock_dims(file.encode(), var.encode(), dims)
        if nd < 0:
            raise KeyError(f"{file}:{var}")
        shape = tuple(dims[d] for d in range(nd))
        out = np.zeros(shape if nd else (1,))
        rc = self.L.ref_mock_get(file.encode(), var.encode(), _p(out), C.c_long(out.size))
        assert rc == 0
        return out if nd else out[0]

    def list(self, file):
        buf = C.create_string_buffer(1 << 16)
        if self.L.ref_mock_list(file.encode(), buf, C.c_int(len(buf))) != 0:
            raise KeyError(file)
        return [s for s in buf.value.decode().split("\n") if s]

    # ---- tools
    def run(self, tool, args):
        argv = [tool.encode()] + [a.encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        rc = self.L.ref_tool_run(tool.encode(), C.c_int(len(argv)), arr)
        return rc

    # ---- find_g_points internals
    def fit_optical_depth_lw(self, method, i1, i2, planck_hl, metric):
        pl, m = _f64(planck_hl), _f64(metric)
        out = np.zeros(m.shape[0])
        rc = self.L.ref_fit_optical_depth_lw(method.encode(), C.c_int(m.shape[0]), C.c_int64(m.shape[1]), C.c_int64(i1),
                                             C.c_int64(i2), _p(pl), _p(m), _p(out))
        assert rc == 0, rc
        return out

    def fit_optical_depth_sw(self, method, i1, i2, ssi, metric):
        s, m = _f64(ssi), _f64(metric)
        out = np.zeros(m.shape[0])
        rc = self.L.ref_fit_optical_depth_sw(method.encode(), C.c_int(m.shape[0]), C.c_int64(m.shape[1]), C.c_int64(i1),
                                             C.c_int64(i2), _p(s), _p(m), _p(out))
        assert rc == 0, rc
        return out

    def fit_optical_depth_sw_total_trans(self, i1, i2, ssi, bg_od, od):
        s, b, o = _f64(ssi), _f64(bg_od), _f64(od)
        out = np.zeros(o.shape[0])
        rc = self.L.ref_fit_optical_depth_sw_total_trans(C.c_int(o.shape[0]), C.c_int64(o.shape[1]), C.c_int64(i1),
                                                         C.c_int64(i2), _p(s), _p(b), _p(o), _p(out))
        assert rc == 0, rc
        return out

    def band_lw(self, method, flux_weight, layer_weight, pressure_hl, surf_emissivity, surf_planck, flux_dn_surf,
                flux_up_toa, planck_hl, bg_od, metric, hr, ibegin=0, iend=None):
        arrs = [_f64(a) for a in (layer_weight, pressure_hl, surf_emissivity, surf_planck, flux_dn_surf, flux_up_toa,
                                  planck_hl, bg_od, metric, hr)]
        self._keep = arrs
        nlay, n = arrs[8].shape
        b = BandLw(method.encode(), flux_weight, nlay, n, ibegin, n - 1 if iend is None else iend,
                   *[a.ctypes.data for a in arrs])
        return b

    def calc_error_lw(self, band, bound1, bound2):
        b1, b2 = _f64(bound1), _f64(bound2)
        err = np.zeros(b1.size)
        rc = self.L.ref_calc_error_lw(C.byref(band), C.c_int(b1.size), _p(b1), _p(b2), _p(err))
        assert rc == 0, rc
        return err

    def equipartition_e_lw(self, band, tolerance, max_iterations=60, tolerance_tolerance=0.02, parallel=False,
                           max_ng=512):
        ng, status = C.c_int(0), C.c_int(0)
        bounds, error = np.zeros(max_ng + 1), np.zeros(max_ng)
        rc = self.L.ref_equipartition_e_lw(C.byref(band), C.c_double(tolerance), C.c_int(max_iterations),
                                           C.c_double(tolerance_tolerance), C.c_int(1 if parallel else 0),
                                           C.c_int(max_ng), C.byref(ng), _p(bounds), _p(error), C.byref(status))
        assert rc == 0, rc
        return status.value, bounds[:ng.value + 1].copy(), error[:ng.value].copy()

    def equipartition_n_lw(self, band, ng, max_iterations=60, tolerance_tolerance=0.02, parallel=False):
        bounds = np.sqrt(np.arange(ng + 1) / ng)  # find_g_points.cpp:1239-1242
        error = np.zeros(ng)
        status = C.c_int(0)
        rc = self.L.ref_equipartition_n_lw(C.byref(band), C.c_int(max_iterations), C.c_double(tolerance_tolerance),
                                           C.c_int(1 if parallel else 0), C.c_int(ng), _p(bounds), _p(error),
                                           C.byref(status))
        assert rc == 0, rc
        return status.value, bounds, error

    def band_sw(self, method, flux_weight, cos_sza, albedo, layer_weight, pressure_hl, ssi, flux_dn_surf, flux_up_toa,
                bg_od, metric, hr, extras=None, min_scaling=0.0, max_scaling=0.0):
        arrs = [_f64(a) for a in (layer_weight, pressure_hl, ssi, flux_dn_surf, flux_up_toa, bg_od, metric, hr)]
        ex = [None] * 6 if extras is None else [_f64(a) for a in extras]
        self._keep = arrs + ex
        nlay, n = arrs[6].shape
        return BandSw(method.encode(), flux_weight, cos_sza, albedo, nlay, n, 0, n - 1,
                      *[a.ctypes.data for a in arrs], *[None if a is None else a.ctypes.data for a in ex],
                      min_scaling, max_scaling)

    def calc_error_sw(self, band, bound1, bound2):
        b1, b2 = _f64(bound1), _f64(bound2)
        err = np.zeros(b1.size)
        rc = self.L.ref_calc_error_sw(C.byref(band), C.c_int(b1.size), _p(b1), _p(b2), _p(err))
        assert rc == 0, rc
        return err

    # ---- the reference's Equipartition over a Python error function (oracle of ecckd_partition_search_*)
    def _ep_args(self, fn, options):
        cb = REF_ERROR_FN(lambda _u, b1, b2: float(fn(b1, b2)))
        o = RefEpOptions(options.get("partition_tolerance", 0.05), options.get("resolution", 0.0),
                         options.get("partition_max_iterations", 20), options.get("line_search_max_iterations", 10),
                         options.get("minimize_frac_range", 1), options.get("cubic_interpolation", 0))
        return cb, o

    def equipartition_n_callback(self, fn, bounds, **options):
        cb, o = self._ep_args(fn, options)
        b = np.array(bounds, dtype=np.float64)
        e = np.zeros(b.size - 1)
        status, calls = C.c_int(0), C.c_long(0)
        rc = self.L.ref_equipartition_n_callback(cb, None, C.byref(o), C.c_int(b.size - 1), _p(b), _p(e), C.byref(status),
                                                 C.byref(calls))
        assert rc == 0, rc
        return status.value, b, e, calls.value

    def equipartition_e_callback(self, fn, target_error, bound0, boundn, max_ni=512, **options):
        cb, o = self._ep_args(fn, options)
        b, e = np.zeros(max_ni + 1), np.zeros(max_ni)
        ni, status, calls = C.c_int(0), C.c_int(0), C.c_long(0)
        rc = self.L.ref_equipartition_e_callback(cb, None, C.byref(o), C.c_double(target_error), C.c_double(bound0),
                                                 C.c_double(boundn), C.c_int(max_ni), C.byref(ni), _p(b), _p(e),
                                                 C.byref(status), C.byref(calls))
        assert rc == 0, rc
        return status.value, b[:ni.value + 1].copy(), e[:ni.value].copy(), calls.value

    # ---- CkdModel
    def ckd_calc_optical_depth(self, file, igas, pressure_hl, temperature_fl, vmr_fl, ng):
        ph, tf = _f64(pressure_hl), _f64(temperature_fl)
        vm = None if vmr_fl is None else _f64(vmr_fl)
        ncol, nhl = ph.shape
        out = np.zeros((ncol, nhl - 1, ng))
        ngo = C.c_int(0)
        rc = self.L.ref_ckd_calc_optical_depth(file.encode(), C.c_int(igas), C.c_int(ncol), C.c_int(nhl - 1), _p(ph),
                                               _p(tf), _p(vm), C.byref(ngo), _p(out), C.c_int64(out.size))
        assert rc == 0 and ngo.value == ng, (rc, ngo.value, self.L.ref_tool_last_exception())
        return out

    def ckd_calc_planck_function(self, file, temperature, ng):
        t = _f64(temperature).ravel()
        out = np.zeros((t.size, ng))
        rc = self.L.ref_ckd_calc_planck_function(file.encode(), C.c_int64(t.size), _p(t), _p(out), C.c_int64(out.size))
        assert rc == 0, rc
        return out

    def ckd_calc_rayleigh_optical_depth(self, file, pressure_hl, ng):
        ph = _f64(pressure_hl)
        ncol, nhl = ph.shape
        out = np.zeros((ncol, nhl - 1, ng))
        rc = self.L.ref_ckd_calc_rayleigh_optical_depth(file.encode(), C.c_int(ncol), C.c_int(nhl - 1), _p(ph), _p(out),
                                                        C.c_int64(out.size))
        assert rc == 0, rc
        return out


_cached = None


def ref_tools():
    global _cached
    if _cached is None:
        _cached = RefTools()
    return _cached
