"""Host-side Python mirror of the reference's operator interface.

Function names and argument meaning follow the reference's C++ free functions
(average_optical_depth_to_g_point, planck_function, radiative_transfer_lw, ...),
so tests read like calls into the reference.  Everything here only marshals
pointers into the C ABI (include/ecckd_b200.h); all arithmetic runs in the
sm_100a kernels.  Arrays on the device are ``DeviceArray`` (a raw device pointer
plus shape); any object exposing ``data_ptr()`` (a torch CUDA tensor) is accepted
wherever a device array is expected.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (INTERVAL_ERROR_FN, EcckdError, GasTable, IntervalErrorLwArgs, IntervalErrorSwArgs, PartitionOptions,
                   check)

AVERAGING_METHODS = (
    "linear",
    "transmission",
    "transmission-2",
    "transmission-3",
    "transmission-10",
    "square-root",
    "logarithmic",
    "hybrid-logarithmic-transmission-3",
    "total-transmission",  # interval fits of find_g_points only
)
WEIGHT_MATRIX, WEIGHT_VECTOR, WEIGHT_PLANCK = 0, 1, 2
CONC_NONE, CONC_LINEAR, CONC_LUT, CONC_RELATIVE_LINEAR = 0, 1, 2, 3


def _ptr(x) -> Optional[int]:
    """Device pointer of a DeviceArray / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, DeviceArray):
        return x.ptr
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


def _h(x, dtype=np.float64):
    """Contiguous host array + its pointer (keeps the array alive via return)."""
    if x is None:
        return None, None
    a = np.ascontiguousarray(x, dtype=dtype)
    return a, a.ctypes.data


class DeviceArray:
    """A typed view of device memory owned by a Context."""

    def __init__(self, ctx: "Context", shape, dtype=np.float64, ptr: Optional[int] = None):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            check(_lib.lib().ecckd_device_malloc(ctx.handle, max(self.nbytes, 1), C.byref(p)))
            ptr = p.value
        self.ptr = int(ptr)

    def copy_from(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        check(_lib.lib().ecckd_memcpy_h2d(self.ctx.handle, self.ptr, a.ctypes.data, a.nbytes))
        self.ctx.synchronize()  # `a` may be a temporary
        return self

    def offset(self, n: int, shape=None) -> "DeviceArray":
        """Unowned view starting n elements into the (flattened) array; keeps nothing alive."""
        total = self.nbytes // self.dtype.itemsize
        if not 0 <= n <= total:
            raise IndexError("offset outside the array")
        return DeviceArray(self.ctx, (total - n,) if shape is None else shape, self.dtype,
                           ptr=self.ptr + int(n) * self.dtype.itemsize)

    def row(self, i: int) -> "DeviceArray":
        """Unowned view of self[i] (C order)."""
        per = int(np.prod(self.shape[1:], dtype=np.int64))
        if not 0 <= i < self.shape[0]:
            raise IndexError("row outside the array")
        return self.offset(i * per, self.shape[1:] or (1,))

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        check(_lib.lib().ecckd_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def free(self):
        if self._owned and self.ptr and self.ctx.handle:
            _lib.lib().ecckd_device_free(self.ctx.handle, self.ptr)
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """ecckd_ctx: one device + one stream."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.handle = None
        h = C.c_void_p()
        check(_lib.lib().ecckd_ctx_create(device, stream, C.byref(h)))
        self.handle = h.value
        self.device = device
        # objects created on this context (plans, models, device arrays) hold a pointer to it in the
        # library: they are closed before the context is, whatever order Python drops them in
        self._children = weakref.WeakSet()

    def set_stream(self, stream: Optional[int]):
        check(_lib.lib().ecckd_ctx_set_stream(self.handle, stream))

    def synchronize(self):
        check(_lib.lib().ecckd_synchronize(self.handle))

    def empty(self, shape, dtype=np.float64) -> DeviceArray:
        return DeviceArray(self, shape, dtype)

    def zeros(self, shape, dtype=np.float64) -> DeviceArray:
        d = DeviceArray(self, shape, dtype)
        return d.copy_from(np.zeros(d.shape, dtype=dtype))

    def to_device(self, a, dtype=None) -> DeviceArray:
        a = np.ascontiguousarray(a, dtype=dtype)
        return DeviceArray(self, a.shape, a.dtype).copy_from(a)

    def set_timing(self, on: bool):
        check(_lib.lib().ecckd_ctx_set_timing(self.handle, 1 if on else 0))

    def reduce_kernel_time(self):
        """(total ms, launches) of the streaming reduction kernel since the last call."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(_lib.lib().ecckd_ctx_reduce_kernel_time(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def kernel_launches(self) -> int:
        return int(_lib.lib().ecckd_kernel_launch_count(self.handle))

    def plan_cache_stats(self):
        """(hits, misses) of the host-array entry points' g-point plan cache."""
        h, m = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().ecckd_ctx_plan_cache_stats(self.handle, C.byref(h), C.byref(m)))
        return int(h.value), int(m.value)

    def close(self):
        if self.handle:
            for child in list(self._children):
                try:
                    child.close()
                except Exception:
                    pass
            _lib.lib().ecckd_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count() -> int:
    return int(_lib.lib().ecckd_device_count())


# ----------------------------------------------------------------- stage 1
def planck_function(ctx: Context, temperature, wavenumber_cm_1, d_wavenumber_cm_1) -> np.ndarray:
    """planck_function (planck_function.cpp:22-54), host arrays in and out."""
    t, tp = _h(temperature)
    wn, wnp = _h(wavenumber_cm_1)
    dwn, dwnp = _h(d_wavenumber_cm_1)
    if wn.size != dwn.size:  # planck_function.cpp:35-38
        raise EcckdError(147, "wavenumber_cm_1 and d_wavenumber_cm_1 must have the same length")
    out = np.empty((t.size, wn.size))
    check(_lib.lib().ecckd_planck_function(ctx.handle, t.size, tp, wn.size, wnp, dwnp, out.ctypes.data, wn.size))
    return out


def debug_eval(ctx: Context, which: int, kd: float, x) -> np.ndarray:
    """ecckd_debug_eval: the kernels' 1 - exp(-kd*x) (which=0) or Planck weight at T=kd, nu=x, dnu=1 (which=1)."""
    xa, xp = _h(x)
    out = np.empty_like(xa)
    check(_lib.lib().ecckd_debug_eval(ctx.handle, int(which), float(kd), xa.size, xp, out.ctypes.data))
    return out


def planck_function_dev(ctx: Context, temperature, d_wavenumber, d_d_wavenumber, nwav: int, d_ans, ans_stride=None):
    t, tp = _h(temperature)
    check(_lib.lib().ecckd_planck_function_dev(ctx.handle, t.size, tp, nwav, _ptr(d_wavenumber), _ptr(d_d_wavenumber),
                                               _ptr(d_ans), ans_stride or nwav))


class GPointPlan:
    """ecckd_gpoint_plan: per-chunk g-sorted visiting order for one g-point map."""

    def __init__(self, ctx: Context, d_g_point, nwav: int, ng: int):
        self.ctx, self.nwav, self.ng = ctx, int(nwav), int(ng)
        self.handle = None
        h = C.c_void_p()
        check(_lib.lib().ecckd_gpoint_plan_create(ctx.handle, _ptr(d_g_point), self.nwav, self.ng, C.byref(h)))
        self.handle = h.value
        ctx._children.add(self)

    def counts(self) -> np.ndarray:
        out = np.zeros(self.ng, dtype=np.int64)
        check(_lib.lib().ecckd_gpoint_plan_counts(self.handle, out.ctypes.data))
        return out

    def counts_dev(self) -> int:
        """Device pointer to int64[ng] counts (summed across ranks when sharded)."""
        p = C.c_void_p()
        check(_lib.lib().ecckd_gpoint_plan_counts_dev(self.handle, C.byref(p)))
        return int(p.value)

    def export(self):
        """ecckd_gpoint_plan_export: (info dict, structured array of the plan's rounds)."""
        info = np.zeros(4, dtype=np.int64)
        check(_lib.lib().ecckd_gpoint_plan_export(self.handle, None, 0, info.ctypes.data))
        dt = np.dtype([("off", np.uint16, (16, 32)), ("meta", np.uint32, (32,)), ("steps", np.uint32),
                       ("chunk", np.uint32), ("flags", np.uint32), ("win", np.uint32)])
        rounds = np.zeros(int(info[0]), dtype=dt)
        check(_lib.lib().ecckd_gpoint_plan_export(self.handle, rounds.ctypes.data, rounds.size, info.ctypes.data))
        return dict(nrounds=int(info[0]), ch=int(info[1]), windowed=bool(info[2]), balanced=bool(info[3])), rounds

    def close(self):
        if self.handle:
            _lib.lib().ecckd_gpoint_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _method_id(averaging_method) -> int:
    if isinstance(averaging_method, str):
        mid = _lib.lib().ecckd_averaging_method_from_string(averaging_method.encode())
        if mid < 0:  # average_optical_depth.cpp:130-133
            raise EcckdError(147, f'averaging_method "{averaging_method}" not understood')
        return mid
    return int(averaging_method)


def average_optical_depth_to_g_point(ctx: Context, ng, reference_surface_vmr, pressure_fl, pressure_hl, g_point,
                                     optical_depth, planck_fl, averaging_method, want_minmax=True):
    """Host-array form with the reference's argument list (average_optical_depth.h:19-30).

    planck_fl may be [nlay][nwav] or a single [nwav] row (spread<0>(ssi, nlay)).
    Returns (molar_abs, min_molar_abs, max_molar_abs, warn).
    """
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, nwav = od.shape
    w = np.ascontiguousarray(planck_fl, dtype=np.float64)
    w_stride = 0 if w.ndim == 1 else nwav
    gp = np.ascontiguousarray(g_point, dtype=np.int32)
    pf, pfp = _h(pressure_fl)
    ph, php = _h(pressure_hl)
    out = np.zeros((nlay, ng))
    omin = np.zeros((nlay, ng)) if want_minmax else None
    omax = np.zeros((nlay, ng)) if want_minmax else None
    warn = np.zeros(ng, dtype=np.int32)
    if not isinstance(averaging_method, str):
        averaging_method = AVERAGING_METHODS[int(averaging_method)]
    check(_lib.lib().ecckd_average_optical_depth_to_g_point(
        ctx.handle, ng, float(reference_surface_vmr), nlay, nwav, pfp, php, gp.ctypes.data, od.ctypes.data, nwav,
        w.ctypes.data, w_stride, averaging_method.encode(), out.ctypes.data,
        omin.ctypes.data if want_minmax else None, omax.ctypes.data if want_minmax else None, ng, warn.ctypes.data))
    return out, omin, omax, warn


def planck_average_optical_depth_to_g_point(ctx: Context, ng, reference_surface_vmr, pressure_fl, pressure_hl, g_point,
                                            optical_depth, temperature_fl, wavenumber_cm_1, d_wavenumber_cm_1,
                                            averaging_method, out=None):
    """planck_function + average_optical_depth_to_g_point (create_look_up_table.cpp:320-331) on host
    arrays in one call; `out` may hold preallocated (molar_abs, min, max, warn) arrays.  A float32
    `optical_depth` (the precision the CKDMIP files hold) is uploaded as such and widened on the device."""
    f32 = isinstance(optical_depth, np.ndarray) and optical_depth.dtype == np.float32
    if f32:
        od = np.ascontiguousarray(optical_depth)
    else:
        od = optical_depth if isinstance(optical_depth, np.ndarray) and optical_depth.flags.c_contiguous and \
            optical_depth.dtype == np.float64 else np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, nwav = od.shape
    gp = np.ascontiguousarray(g_point, dtype=np.int32)
    pf, pfp = _h(pressure_fl)
    ph, php = _h(pressure_hl)
    tf, tfp = _h(temperature_fl)
    wn, wnp = _h(wavenumber_cm_1)
    dwn, dwnp = _h(d_wavenumber_cm_1)
    if out is None:
        out = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32))
    fn = _lib.lib().ecckd_planck_average_optical_depth_to_g_point_f32 if f32 else \
        _lib.lib().ecckd_planck_average_optical_depth_to_g_point
    check(fn(
        ctx.handle, ng, float(reference_surface_vmr), nlay, nwav, pfp, php, gp.ctypes.data, od.ctypes.data, nwav, tfp,
        wnp, dwnp, averaging_method.encode(), out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, ng,
        out[3].ctypes.data))
    return out


def create_lut_block(ctx: Context, ng, reference_surface_vmr, pressure_hl, temperature_hl, g_point, optical_depth,
                     averaging_method, wavenumber_cm_1=None, d_wavenumber_cm_1=None, ssi=None):
    """One block of the create_look_up_table driver loop (create_look_up_table.cpp:293-344)
    -> (molar_abs, min, max, warn, temperature_fl)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, nwav = od.shape
    gp = np.ascontiguousarray(g_point, dtype=np.int32)
    ph, php = _h(pressure_hl)
    th, thp = _h(temperature_hl)
    wn, wnp = _h(wavenumber_cm_1)
    dwn, dwnp = _h(d_wavenumber_cm_1)
    s, sp = _h(ssi)
    out = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32))
    t_fl = np.zeros(nlay)
    check(_lib.lib().ecckd_create_lut_block(
        ctx.handle, ng, float(reference_surface_vmr), nlay, nwav, php, thp, gp.ctypes.data, od.ctypes.data, nwav, wnp, dwnp,
        sp, averaging_method.encode(), t_fl.ctypes.data, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, ng,
        out[3].ctypes.data))
    return out + (t_fl,)


EP_STATUS = ("success", "max-iterations-reached", "failed-to-converge", "resolution-limit-reached", "no-progress",
             "failure", "input-error")  # EpStatus, equipartition.h:31-39


class PartitionSearch:
    """ecckd_partition_search: the Equipartition driver of find_g_points (equipartition.cpp:160-805) over a batched
    evaluator.  `evaluator` is either a Python callable (bound1[n], bound2[n]) -> error[n], or the pair
    ("lw" | "sw", filled IntervalErrorLwArgs / IntervalErrorSwArgs) selecting the device evaluators
    ecckd_interval_error_{lw,sw}_eval (no Python in the loop).  Options: fields of ecckd_partition_options."""

    def __init__(self, evaluator, **options):
        L = _lib.lib()
        self.handle = None
        self.opt = PartitionOptions()
        L.ecckd_partition_options_default(C.byref(self.opt))
        for k, v in options.items():
            if not hasattr(self.opt, k):
                raise EcckdError(147, f"unknown partition option {k}")
            setattr(self.opt, k, v)
        self.error = None
        if callable(evaluator):
            def trampoline(_user, n, b1, b2, err):
                try:
                    out = evaluator(np.ctypeslib.as_array(b1, (n,)).copy(), np.ctypeslib.as_array(b2, (n,)).copy())
                    np.ctypeslib.as_array(err, (n,))[:] = np.asarray(out, dtype=np.float64)
                    return 0
                except EcckdError as e:
                    self.error = e
                    return e.code if getattr(e, "code", 0) else 131
                except Exception as e:  # noqa: BLE001 - reported to the caller of the search
                    self.error = e
                    return 131
            self._cb = INTERVAL_ERROR_FN(trampoline)
            fn, user = C.cast(self._cb, C.c_void_p).value, None
        else:
            kind, self._args = evaluator
            f = L.ecckd_interval_error_lw_eval if kind == "lw" else L.ecckd_interval_error_sw_eval
            fn, user = C.cast(f, C.c_void_p).value, C.addressof(self._args)
        h = C.c_void_p()
        check(L.ecckd_partition_search_create(C.addressof(self.opt), fn, user, C.byref(h)))
        self.handle = h.value

    def _done(self, rc):
        if rc != 0 and self.error is not None:
            e, self.error = self.error, None
            raise e
        check(rc)

    def equipartition_n(self, bounds, error=None):
        """-> (status string, bounds[ni+1], error[ni])"""
        b = np.array(bounds, dtype=np.float64)
        ni = b.size - 1
        e = np.zeros(ni) if error is None else np.array(error, dtype=np.float64)
        st = C.c_int(0)
        self._done(_lib.lib().ecckd_partition_search_n(self.handle, ni, b.ctypes.data, e.ctypes.data,
                                                      0 if error is None else 1, C.addressof(st)))
        return EP_STATUS[st.value], b, e

    def equipartition_e(self, target_error, bound0, boundn, capacity=256):
        b, e = np.zeros(capacity + 1), np.zeros(capacity)
        ni, st = C.c_int(0), C.c_int(0)
        self._done(_lib.lib().ecckd_partition_search_e(self.handle, float(target_error), float(bound0), float(boundn),
                                                      capacity, C.addressof(ni), b.ctypes.data, e.ctypes.data,
                                                      C.addressof(st)))
        return EP_STATUS[st.value], b[:ni.value + 1].copy(), e[:ni.value].copy()

    def band(self, npoints, target_error, min_g_points=1, max_g_points=256, capacity=256):
        """One band of find_g_points (find_g_points.cpp:1229-1256, 1396-1400) -> (status, bounds, error, rank1, rank2)"""
        b, e = np.zeros(capacity + 1), np.zeros(capacity)
        r1, r2 = np.zeros(capacity, dtype=np.int64), np.zeros(capacity, dtype=np.int64)
        ng, st = C.c_int(0), C.c_int(0)
        self._done(_lib.lib().ecckd_partition_search_band(self.handle, int(npoints), float(target_error), min_g_points,
                                                         max_g_points, capacity, C.addressof(ng), b.ctypes.data,
                                                         e.ctypes.data, r1.ctypes.data, r2.ctypes.data, C.addressof(st)))
        n = ng.value
        return EP_STATUS[st.value], b[:n + 1].copy(), e[:n].copy(), r1[:n].copy(), r2[:n].copy()

    def stats(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(_lib.lib().ecckd_partition_search_stats(self.handle, C.addressof(a), C.addressof(b), C.addressof(c)))
        return dict(batches=a.value, intervals=b.value, speculative_unused=c.value)

    def close(self):
        if self.handle:
            _lib.lib().ecckd_partition_search_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interval_error_lw_evaluator(ctx, averaging_method, nlay, n, stride, d_pressure_hl, d_planck_hl, d_surf_emissivity,
                                d_surf_planck, d_bg_od, d_metric, d_flux_dn_surf, d_flux_up_toa, d_hr, flux_weight,
                                d_layer_weight, offset=0):
    """The evaluator argument of PartitionSearch for one longwave band on the device: the spectral arrays are in
    rank order, the band is the `n` points from `offset` (find_g_points.cpp:1158-1166: ibegin..iend)."""
    o = int(offset) * 8

    def sp(d):
        return _ptr(d) + o
    a = IntervalErrorLwArgs(ctx.handle, _method_id(averaging_method), nlay, n, stride, _ptr(d_pressure_hl),
                            sp(d_planck_hl), sp(d_surf_emissivity), sp(d_surf_planck), sp(d_bg_od), sp(d_metric),
                            sp(d_flux_dn_surf), sp(d_flux_up_toa), sp(d_hr), float(flux_weight), _ptr(d_layer_weight))
    return ("lw", a)


def interval_error_sw_evaluator(ctx, averaging_method, nlay, n, stride, cos_sza, d_pressure_hl, d_ssi, albedo, d_bg_od,
                                d_metric, d_flux_dn_surf, d_flux_up_toa, d_hr, flux_weight, d_layer_weight, low=None,
                                high=None, min_scaling=1.0, max_scaling=1.0, offset=0):
    """As interval_error_lw_evaluator, shortwave (low/high: the total-transmission extras, find_g_points.cpp:1175-1180)."""
    o = int(offset) * 8

    def sp(d):
        return None if d is None else _ptr(d) + o
    low = low or (None, None, None)
    high = high or (None, None, None)
    a = IntervalErrorSwArgs(ctx.handle, _method_id(averaging_method), nlay, n, stride, float(cos_sza),
                            _ptr(d_pressure_hl), sp(d_ssi), float(albedo), sp(d_bg_od), sp(d_metric),
                            sp(d_flux_dn_surf), sp(d_flux_up_toa), sp(d_hr), float(flux_weight), _ptr(d_layer_weight),
                            sp(low[0]), sp(low[1]), sp(low[2]), sp(high[0]), sp(high[1]), sp(high[2]),
                            float(min_scaling), float(max_scaling))
    return ("sw", a)


class Comm:
    """ecckd_comm: the NCCL communicator of the sharded reduction (one rank per GPU).  `exchange` is a
    callable that hands rank 0's 128-byte id to every rank, e.g.
    ``lambda b: torch.distributed.broadcast_object_list(b, src=0)`` on a one-element list."""

    def __init__(self, ctx: Context, rank: int, world: int, exchange=None, unique_id: Optional[bytes] = None):
        self.ctx, self.rank, self.world, self.handle = ctx, int(rank), int(world), None
        if unique_id is None:
            box = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                check(_lib.lib().ecckd_comm_unique_id(C.cast(buf, C.c_void_p).value))
                box[0] = buf.raw
            if world > 1:
                if exchange is None:
                    raise EcckdError(147, "Comm: world > 1 needs `exchange` or `unique_id`")
                exchange(box)
            unique_id = box[0]
        buf = C.create_string_buffer(bytes(unique_id), 128)
        h = C.c_void_p()
        check(_lib.lib().ecckd_comm_create(ctx.handle, self.world, self.rank, C.cast(buf, C.c_void_p).value, C.byref(h)))
        self.handle = h.value
        ctx._children.add(self)

    def close(self):
        if self.handle:
            _lib.lib().ecckd_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def average_optical_depth_exchange_dev(ctx: Context, comm: Optional[Comm], seg_begin, d_seg_sums, d_seg_minmax,
                                       d_sums, d_minmax, d_seg_count=None, d_count=None):
    """ecckd_average_optical_depth_exchange_dev: d_seg_*[nseg_total][...] slots -> combined d_* on every rank, in
    segment order (bitwise independent of the number of ranks).  seg_begin[nranks + 1]: first segment of each rank."""
    sb = np.ascontiguousarray(seg_begin, dtype=np.int32)
    nseg = int(sb[-1])
    numel = lambda a: 0 if a is None else (int(a.numel()) if hasattr(a, "numel") else int(np.prod(a.shape)))
    n_sums, n_mm, n_cnt = numel(d_sums), numel(d_minmax), numel(d_count)
    check(_lib.lib().ecckd_average_optical_depth_exchange_dev(
        ctx.handle, None if comm is None else comm.handle, nseg, sb.ctypes.data, n_sums, n_mm, n_cnt, _ptr(d_seg_sums),
        _ptr(d_seg_minmax), _ptr(d_seg_count), _ptr(d_sums), _ptr(d_minmax), _ptr(d_count)))


class LutBuilder:
    """ecckd_lut_builder: the state of the create_look_up_table block loop on the device
    (create_look_up_table.cpp:249-591) - g-point map, its plan and the wavenumber grid (longwave) or the
    solar spectrum (shortwave) are uploaded once; block() then moves only that block's optical depths.
    float32 optical depths (what the spectra files store) are widened on the device."""

    def __init__(self, ctx: Context, ng, g_point, wavenumber_cm_1=None, d_wavenumber_cm_1=None, ssi=None):
        self.ctx, self.ng, self.handle = ctx, int(ng), None
        gp = np.ascontiguousarray(g_point, dtype=np.int32)
        self.nwav = gp.size
        wn, wnp = _h(wavenumber_cm_1)
        dwn, dwnp = _h(d_wavenumber_cm_1)
        s, sp = _h(ssi)
        h = C.c_void_p()
        check(_lib.lib().ecckd_lut_builder_create(ctx.handle, self.ng, self.nwav, gp.ctypes.data, wnp, dwnp, sp,
                                                  C.byref(h)))
        self.handle = h.value
        ctx._children.add(self)

    def block(self, reference_surface_vmr, pressure_hl, temperature_hl, optical_depth, averaging_method, out=None):
        """-> (molar_abs, min, max, warn, temperature_fl); optical_depth [nlay][nwav] float64 or float32."""
        od = np.asarray(optical_depth)
        f32 = od.dtype == np.float32
        od = np.ascontiguousarray(od, dtype=np.float32 if f32 else np.float64)
        nlay, nwav = od.shape
        if nwav != self.nwav:
            raise EcckdError(147, "optical_depth does not match the builder's wavenumber grid")
        ph, php = _h(pressure_hl)
        th, thp = _h(temperature_hl)
        ng = self.ng
        if out is None:
            out = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32),
                   np.zeros(nlay))
        check(_lib.lib().ecckd_lut_builder_block(
            self.handle, float(reference_surface_vmr), nlay, php, thp, None if f32 else od.ctypes.data,
            od.ctypes.data if f32 else None, nwav, averaging_method.encode(), out[4].ctypes.data, out[0].ctypes.data,
            out[1].ctypes.data, out[2].ctypes.data, ng, out[3].ctypes.data))
        return out

    def close(self):
        if self.handle:
            _lib.lib().ecckd_lut_builder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def average_optical_depth_to_g_point_dev(ctx: Context, plan: GPointPlan, reference_surface_vmr, pressure_fl,
                                         pressure_hl, d_od, od_stride, weight_mode, d_weight, w_stride, temperature,
                                         d_wavenumber, d_d_wavenumber, averaging_method, d_molar_abs, d_min=None,
                                         d_max=None, out_stride=None, d_warn=None):
    pf, pfp = _h(pressure_fl)
    ph, php = _h(pressure_hl)
    t, tp = _h(temperature)
    check(_lib.lib().ecckd_average_optical_depth_to_g_point_dev(
        ctx.handle, plan.handle, float(reference_surface_vmr), pf.size, pfp, php, _ptr(d_od), od_stride, weight_mode,
        _ptr(d_weight), w_stride, tp, _ptr(d_wavenumber), _ptr(d_d_wavenumber), _method_id(averaging_method),
        _ptr(d_molar_abs), _ptr(d_min), _ptr(d_max), out_stride or plan.ng, _ptr(d_warn)))


def _ptr_array(items: Optional[Sequence]):
    if items is None:
        return None, None
    arr = (C.c_void_p * len(items))(*[_ptr(x) for x in items])
    return arr, C.cast(arr, C.c_void_p).value


def average_optical_depth_batch_dev(ctx: Context, plan: GPointPlan, reference_surface_vmr: Sequence[float], pressure_fl,
                                    pressure_hl, d_ods: Sequence, od_stride, weight_mode, d_weight, w_stride,
                                    temperature, d_wavenumber, d_d_wavenumber, averaging_method, d_molar_abs: Sequence,
                                    d_min: Optional[Sequence] = None, d_max: Optional[Sequence] = None,
                                    out_stride=None, d_warn: Optional[Sequence] = None):
    pf, pfp = _h(pressure_fl)
    ph, php = _h(pressure_hl)
    t, tp = _h(temperature)
    v, vp = _h(reference_surface_vmr)
    a_od, p_od = _ptr_array(d_ods)
    a_out, p_out = _ptr_array(d_molar_abs)
    a_min, p_min = _ptr_array(d_min)
    a_max, p_max = _ptr_array(d_max)
    a_w, p_w = _ptr_array(d_warn)
    check(_lib.lib().ecckd_average_optical_depth_batch_dev(
        ctx.handle, plan.handle, len(d_ods), vp, pf.size, pfp, php, p_od, od_stride, weight_mode, _ptr(d_weight),
        w_stride, tp, _ptr(d_wavenumber), _ptr(d_d_wavenumber), _method_id(averaging_method), p_out, p_min, p_max,
        out_stride or plan.ng, p_w))


def average_optical_depth_partial_dev(ctx: Context, plan: GPointPlan, nlay, pressure_fl, d_ods: Sequence, od_stride,
                                      weight_mode, d_weight, w_stride, temperature, d_wavenumber, d_d_wavenumber,
                                      averaging_method, d_sums, d_minmax):
    """Per-shard streaming stage; d_sums [nlay][nspec][ng][3], d_minmax [nlay][nspec][ng][2]."""
    pf, pfp = _h(pressure_fl)
    t, tp = _h(temperature)
    a_od, p_od = _ptr_array(d_ods)
    check(_lib.lib().ecckd_average_optical_depth_partial_dev(
        ctx.handle, plan.handle, len(d_ods), nlay, pfp, p_od, od_stride, weight_mode, _ptr(d_weight), w_stride, tp,
        _ptr(d_wavenumber), _ptr(d_d_wavenumber), _method_id(averaging_method), _ptr(d_sums), _ptr(d_minmax)))


def average_optical_depth_finalize_dev(ctx: Context, ng, d_count, reference_surface_vmr: Sequence[float], pressure_fl,
                                       pressure_hl, averaging_method, d_sums, d_minmax, d_molar_abs: Sequence,
                                       d_min: Optional[Sequence] = None, d_max: Optional[Sequence] = None,
                                       out_stride=None, d_warn: Optional[Sequence] = None):
    pf, pfp = _h(pressure_fl)
    ph, php = _h(pressure_hl)
    v, vp = _h(reference_surface_vmr)
    a_out, p_out = _ptr_array(d_molar_abs)
    a_min, p_min = _ptr_array(d_min)
    a_max, p_max = _ptr_array(d_max)
    a_w, p_w = _ptr_array(d_warn)
    check(_lib.lib().ecckd_average_optical_depth_finalize_dev(
        ctx.handle, ng, _ptr(d_count), len(d_molar_abs), vp, pf.size, pfp, php, _method_id(averaging_method),
        _ptr(d_sums), _ptr(d_minmax), p_out, p_min, p_max, out_stride or ng, p_w))


def merge_axpy_dev(ctx: Context, nlay, nwav, d_optical_depth, stride, d_od, od_stride, scaling=1.0,
                   scaling_profile=None):
    sp, spp = _h(scaling_profile)
    check(_lib.lib().ecckd_merge_axpy_dev(ctx.handle, nlay, nwav, _ptr(d_optical_depth), stride, _ptr(d_od), od_stride,
                                          float(scaling), spp))


def merge_spectra_dev(ctx: Context, nlay, nwav, d_ods: Sequence, od_strides: Sequence[int], scalings: Sequence[float],
                      d_optical_depth, stride):
    """read_merged_spectrum.cpp:149-168 for scalar scalings, all sources in one pass."""
    a_od, p_od = _ptr_array(d_ods)
    st, stp = _h(od_strides, np.int64)
    sc, scp = _h(scalings)
    check(_lib.lib().ecckd_merge_spectra_dev(ctx.handle, nlay, nwav, len(d_ods), p_od, stp, scp, _ptr(d_optical_depth),
                                             stride))


def gpoint_fraction_dev(ctx: Context, ng, nwav, d_g_point, d_wavenumber, d_d_wavenumber, wavenumber1, wavenumber2,
                        d_out):
    w1, w1p = _h(wavenumber1)
    w2, w2p = _h(wavenumber2)
    check(_lib.lib().ecckd_gpoint_fraction_dev(ctx.handle, ng, nwav, _ptr(d_g_point), _ptr(d_wavenumber),
                                               _ptr(d_d_wavenumber), w1.size, w1p, w2p, _ptr(d_out)))


def ssi_intervals_dev(ctx: Context, nwav, d_wavenumber, d_ssi, wavenumber1, wavenumber2, d_out):
    w1, w1p = _h(wavenumber1)
    w2, w2p = _h(wavenumber2)
    check(_lib.lib().ecckd_ssi_intervals_dev(ctx.handle, nwav, _ptr(d_wavenumber), _ptr(d_ssi), w1.size, w1p, w2p,
                                             _ptr(d_out)))


def planck_lut_dev(ctx: Context, plan: GPointPlan, d_wavenumber, d_d_wavenumber, temperature_lut, d_out):
    t, tp = _h(temperature_lut)
    check(_lib.lib().ecckd_planck_lut_dev(ctx.handle, plan.handle, _ptr(d_wavenumber), _ptr(d_d_wavenumber), t.size, tp,
                                          _ptr(d_out)))


def gpoint_sum_dev(ctx: Context, plan: GPointPlan, nrow, d_values, stride, d_out):
    """out[row][g] = sum of values[row][w] over the wavenumbers w of g point g (lw_spectra.cpp:250-255,
    scale_lut.cpp:119-123)."""
    check(_lib.lib().ecckd_gpoint_sum_dev(ctx.handle, plan.handle, nrow, _ptr(d_values), stride, _ptr(d_out)))


def row_sum_dev(ctx: Context, nrow, n, d_values, stride, d_out):
    """out[row] = sum of values[row][:n] (sum(x, 1) of lw_spectra.cpp:230-231)."""
    check(_lib.lib().ecckd_row_sum_dev(ctx.handle, nrow, n, _ptr(d_values), stride, _ptr(d_out)))


def lw_spectra_profile(ctx: Context, plan: GPointPlan, d_wavenumber, d_d_wavenumber, pressure_hl, temperature_hl,
                       d_optical_depth, nwav):
    """The per-profile body of the lw_spectra tool (lw_spectra.cpp:208-255).  Returns a dict of host arrays:
    flux_dn_lw, flux_up_lw [nlay+1]; optical_depth [nlay][ng]; spectral_flux_dn_lw, spectral_flux_up_lw [nlay+1][ng]."""
    ph, php = _h(pressure_hl)
    th, thp = _h(temperature_hl)
    nlay = ph.size - 1
    ng = plan.ng
    bdn, bup = ctx.empty(nlay + 1), ctx.empty(nlay + 1)
    sod, sdn, sup = ctx.empty((nlay, ng)), ctx.empty((nlay + 1, ng)), ctx.empty((nlay + 1, ng))
    check(_lib.lib().ecckd_lw_spectra_profile_dev(ctx.handle, plan.handle, nlay, _ptr(d_wavenumber), _ptr(d_d_wavenumber),
                                                  php, thp, _ptr(d_optical_depth), nwav, _ptr(bdn), _ptr(bup), _ptr(sod),
                                                  _ptr(sdn), _ptr(sup)))
    return dict(flux_dn_lw=bdn.numpy(), flux_up_lw=bup.numpy(), optical_depth=sod.numpy(),
                spectral_flux_dn_lw=sdn.numpy(), spectral_flux_up_lw=sup.numpy())


def scale_lut_scaling(mu0, flux_dn_gpoint, od_total=None):
    """scale_lut.cpp:117-133, 180-181: (od_best, scaling) from the per-g-point direct fluxes [nz+1][ng]."""
    f, fp = _h(flux_dn_gpoint)
    nz, ng = f.shape[0] - 1, f.shape[1]
    best = np.zeros((nz, ng))
    scaling = np.zeros((nz, ng)) if od_total is not None else None
    t, tp = _h(od_total)
    check(_lib.lib().ecckd_scale_lut_scaling(nz, ng, float(mu0), fp, tp, best.ctypes.data,
                                             scaling.ctypes.data if scaling is not None else None))
    return best, scaling


def compact_g_points_dev(ctx: Context, ng, nwav, d_g_point):
    """Returns (new_ng, g_point_map[new_ng]); d_g_point renumbered in place."""
    gmap = np.zeros(ng, dtype=np.int32)
    n = C.c_int(0)
    check(_lib.lib().ecckd_compact_g_points_dev(ctx.handle, ng, nwav, _ptr(d_g_point), gmap.ctypes.data, C.byref(n)))
    return n.value, gmap[: n.value].copy()


# ----------------------------------------------------------------- ranking
def sort_key_lw_dev(ctx, nlay, nwav, pressure_hl, d_wavenumber, d_d_wavenumber, d_od, od_stride, threshold, d_key,
                    d_column_od=None):
    ph, php = _h(pressure_hl)
    check(_lib.lib().ecckd_sort_key_lw_dev(ctx.handle, nlay, nwav, php, _ptr(d_wavenumber), _ptr(d_d_wavenumber),
                                           _ptr(d_od), od_stride, float(threshold), _ptr(d_key), _ptr(d_column_od)))


def sort_key_sw_dev(ctx, nlay, nwav, pressure_hl, d_od, od_stride, threshold, d_key, d_column_od=None):
    ph, php = _h(pressure_hl)
    check(_lib.lib().ecckd_sort_key_sw_dev(ctx.handle, nlay, nwav, php, _ptr(d_od), od_stride, float(threshold),
                                           _ptr(d_key), _ptr(d_column_od)))


def rank_bands_dev(ctx, nwav, d_wavenumber, d_key, band_bound1, band_bound2, d_iband=None, d_rank=None,
                   d_ordered_index=None):
    b1, b1p = _h(band_bound1)
    b2, b2p = _h(band_bound2)
    check(_lib.lib().ecckd_rank_bands_dev(ctx.handle, nwav, _ptr(d_wavenumber), _ptr(d_key), b1.size, b1p, b2p,
                                          _ptr(d_iband), _ptr(d_rank), _ptr(d_ordered_index)))


def store_g_points_dev(ctx, nwav, d_rank, rank1, rank2, d_g_point):
    r1, r1p = _h(rank1, np.int32)
    r2, r2p = _h(rank2, np.int32)
    check(_lib.lib().ecckd_store_g_points_dev(ctx.handle, nwav, _ptr(d_rank), r1.size, r1p, r2p, _ptr(d_g_point)))


def overlap_g_points(n_g_points, sorting_variables: Sequence[np.ndarray]):
    """overlap_g_points (single_gas_data.cpp:23-124) -> (ng, g_min[ngas][ng], g_max[ngas][ng], band_number[ng])."""
    n_g = np.ascontiguousarray(n_g_points, dtype=np.int32)
    ngas, nband = n_g.shape
    sv = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in sorting_variables]))
    offs = np.zeros(ngas, dtype=np.int32)
    offs[1:] = np.cumsum([len(s) for s in sorting_variables])[:-1]
    cap = int(n_g.sum()) + 1
    g_min = np.zeros((ngas, cap), dtype=np.int32)
    g_max = np.zeros((ngas, cap), dtype=np.int32)
    band = np.zeros(cap, dtype=np.int32)
    ng = C.c_int(0)
    check(_lib.lib().ecckd_overlap_g_points(ngas, nband, n_g.ctypes.data, sv.ctypes.data, offs.ctypes.data, cap,
                                            g_min.ctypes.data, g_max.ctypes.data, band.ctypes.data, C.byref(ng)))
    n = ng.value
    return n, g_min[:, :n].copy(), g_max[:, :n].copy(), band[:n].copy()


def assign_g_points_dev(ctx, nwav, d_gas_g_point, g_min, g_max, d_g_point) -> int:
    gmin = np.ascontiguousarray(g_min, dtype=np.int32)
    gmax = np.ascontiguousarray(g_max, dtype=np.int32)
    ngas, ng = gmin.shape
    n_un = C.c_int64(0)
    check(_lib.lib().ecckd_assign_g_points_dev(ctx.handle, ngas, nwav, _ptr(d_gas_g_point), ng, ng, gmin.ctypes.data,
                                               gmax.ctypes.data, _ptr(d_g_point), C.byref(n_un)))
    return int(n_un.value)


# ----------------------------------------------------------------- stage 2
class CkdModel:
    """ecckd_ckd_model: the look-up tables of a CKD definition on the device.

    gases: sequence of dicts with keys conc_dependence, molar_abs and, as
    needed, vmr_lut and reference_vmr (the per-gas variables of
    ckd_model.cpp:288-641).
    """

    def __init__(self, ctx: Context, gases, log_pressure, temperature, temperature_planck=None, planck_lut=None,
                 solar_irradiance=None, rayleigh_molar_scat=None):
        self.ctx = ctx
        self.handle = None
        keep = []
        tabs = (GasTable * len(gases))()
        for i, g in enumerate(gases):
            k, kp = _h(g["molar_abs"])
            v, vp = _h(g.get("vmr_lut"))
            keep += [k, v]
            tabs[i].conc_dependence = int(g["conc_dependence"])
            tabs[i].nconc = 0 if v is None else v.size
            tabs[i].vmr_lut = vp
            tabs[i].reference_vmr = float(g.get("reference_vmr", 0.0))
            tabs[i].molar_abs = kp
        lp, lpp = _h(log_pressure)
        t, tp = _h(temperature)
        tpl, tplp = _h(temperature_planck)
        pl, plp = _h(planck_lut)
        si, sip = _h(solar_irradiance)
        ra, rap = _h(rayleigh_molar_scat)
        self.nt, self.np = t.shape
        self.ng = int(np.asarray(gases[0]["molar_abs"]).shape[-1])
        self.ngas = len(gases)
        self.is_sw = si is not None
        h = C.c_void_p()
        check(_lib.lib().ecckd_ckd_model_create(ctx.handle, len(gases), C.cast(tabs, C.c_void_p).value, self.nt,
                                                self.np, self.ng, lpp, tp, 0 if tpl is None else tpl.size, tplp, plp,
                                                sip, rap, C.byref(h)))
        self.handle = h.value
        ctx._children.add(self)

    def calc_optical_depth_dev(self, igas, ncol, nlay, d_pressure_hl, d_temperature_fl, d_vmr_fl, d_od,
                               accumulate=False):
        check(_lib.lib().ecckd_ckd_calc_optical_depth_dev(self.ctx.handle, self.handle, igas, ncol, nlay,
                                                          _ptr(d_pressure_hl), _ptr(d_temperature_fl), _ptr(d_vmr_fl),
                                                          _ptr(d_od), 1 if accumulate else 0))

    def calc_planck_function_dev(self, n, d_temperature, d_planck):
        check(_lib.lib().ecckd_ckd_calc_planck_function_dev(self.ctx.handle, self.handle, n, _ptr(d_temperature),
                                                            _ptr(d_planck)))

    def calc_rayleigh_optical_depth_dev(self, ncol, nlay, d_pressure_hl, d_od):
        check(_lib.lib().ecckd_ckd_calc_rayleigh_optical_depth_dev(self.ctx.handle, self.handle, ncol, nlay,
                                                                   _ptr(d_pressure_hl), _ptr(d_od)))

    def run_ckd_dev(self, d_vmr_fl: Sequence, tsi, ncol, nlay, d_pressure_hl, d_temperature_hl, d_od=None,
                    d_planck_hl=None, d_planck_surf=None, d_flux_dn=None, d_flux_up=None, d_bb_flux_dn=None,
                    d_bb_flux_up=None):
        arr, p = _ptr_array(d_vmr_fl)
        check(_lib.lib().ecckd_run_ckd_dev(self.ctx.handle, self.handle, p, float(tsi), ncol, nlay, _ptr(d_pressure_hl),
                                           _ptr(d_temperature_hl), _ptr(d_od), _ptr(d_planck_hl), _ptr(d_planck_surf),
                                           _ptr(d_flux_dn), _ptr(d_flux_up), _ptr(d_bb_flux_dn), _ptr(d_bb_flux_up)))

    def calc_total_optical_depth_dev(self, d_vmr_fl: Sequence, ncol, nlay, d_pressure_hl, d_temperature_hl, d_od):
        """calc_total_optical_depth (solve_adept.cpp:23-69); d_vmr_fl[igas] = device array or None."""
        arr, p = _ptr_array(d_vmr_fl)
        check(_lib.lib().ecckd_calc_total_optical_depth_dev(self.ctx.handle, self.handle, p, ncol, nlay,
                                                            _ptr(d_pressure_hl), _ptr(d_temperature_hl), _ptr(d_od)))

    def calc_total_optical_depth_adjoint_dev(self, d_vmr_fl: Sequence, ncol, nlay, d_pressure_hl, d_temperature_hl, d_grad_od,
                                             d_grad_molar_abs: Sequence, accumulate=False):
        """Adjoint of calc_total_optical_depth: d_grad_molar_abs[igas] (device array shaped like the gas's table, or
        None) += / = sum over (column, layer) of d_grad_od times the interpolation coefficients (ckd_model.cpp:962-1088
        transposed)."""
        arr, p = _ptr_array(d_vmr_fl)
        garr, gp = _ptr_array(d_grad_molar_abs)
        check(_lib.lib().ecckd_calc_total_optical_depth_adjoint_dev(
            self.ctx.handle, self.handle, p, ncol, nlay, _ptr(d_pressure_hl), _ptr(d_temperature_hl), _ptr(d_grad_od), gp,
            1 if accumulate else 0))

    def run_ckd(self, vmr_fl: Sequence[Optional[np.ndarray]], pressure_hl, temperature_hl, tsi=1361.0, out=None):
        """Numerical core of run_ckd (run_ckd.cpp:119-367) on host arrays -> dict of outputs (written into
        the C-contiguous float64 arrays of `out` when given, e.g. pinned buffers)."""
        ph = np.ascontiguousarray(pressure_hl, dtype=np.float64)
        th = np.ascontiguousarray(temperature_hl, dtype=np.float64)
        ncol, nhl = ph.shape
        nlay, ng = nhl - 1, self.ng
        keep = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in vmr_fl]
        arr = (C.c_void_p * len(keep))(*[None if v is None else v.ctypes.data for v in keep])
        if out is None:
            out = {"optical_depth": np.zeros((ncol, nlay, ng)), "flux_dn": np.zeros((ncol, nhl, ng)),
                   "bb_flux_dn": np.zeros((ncol, nhl))}
            if not self.is_sw:
                out.update(planck_hl=np.zeros((ncol, nhl, ng)), planck_surf=np.zeros((ncol, ng)),
                           flux_up=np.zeros((ncol, nhl, ng)), bb_flux_up=np.zeros((ncol, nhl)))
        else:
            shapes = {"optical_depth": (ncol, nlay, ng), "flux_dn": (ncol, nhl, ng), "flux_up": (ncol, nhl, ng),
                      "planck_hl": (ncol, nhl, ng), "planck_surf": (ncol, ng), "bb_flux_dn": (ncol, nhl),
                      "bb_flux_up": (ncol, nhl)}
            for k, a in out.items():
                if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != shapes[k]:
                    raise EcckdError(147, f"run_ckd: out[{k!r}] must be a C-contiguous float64 array of shape {shapes[k]}")
        g = lambda k: out[k].ctypes.data if k in out else None
        check(_lib.lib().ecckd_run_ckd(self.ctx.handle, self.handle, C.cast(arr, C.c_void_p).value, float(tsi), ncol,
                                       nlay, ph.ctypes.data, th.ctypes.data, g("optical_depth"), g("planck_hl"),
                                       g("planck_surf"), g("flux_dn"), g("flux_up"), g("bb_flux_dn"), g("bb_flux_up")))
        return out

    def close(self):
        if self.handle:
            _lib.lib().ecckd_ckd_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def radiative_transfer_lw_dev(ctx, ncol, nlay, n, d_planck, d_od, d_surf_emissivity, d_surf_planck, d_flux_dn,
                              d_flux_up):
    check(_lib.lib().ecckd_radiative_transfer_lw_dev(ctx.handle, ncol, nlay, n, _ptr(d_planck), _ptr(d_od),
                                                     _ptr(d_surf_emissivity), _ptr(d_surf_planck), _ptr(d_flux_dn),
                                                     _ptr(d_flux_up)))


def radiative_transfer_sw_dev(ctx, ncol, nlay, n, cos_sza, d_ssi, d_od, d_albedo, d_flux_dn, d_flux_up=None):
    mu, mup = _h(np.broadcast_to(np.asarray(cos_sza, dtype=np.float64), (ncol,)))
    check(_lib.lib().ecckd_radiative_transfer_sw_dev(ctx.handle, ncol, nlay, n, mup, _ptr(d_ssi), _ptr(d_od),
                                                     _ptr(d_albedo), _ptr(d_flux_dn), _ptr(d_flux_up)))


def heating_rate_dev(ctx, ncol, nlay, n, d_pressure_hl, d_flux_dn, d_flux_up, d_hr):
    check(_lib.lib().ecckd_heating_rate_dev(ctx.handle, ncol, nlay, n, _ptr(d_pressure_hl), _ptr(d_flux_dn),
                                            _ptr(d_flux_up), _ptr(d_hr)))


# Convenience host-array wrappers that mirror the reference call signatures.
def radiative_transfer_lw(ctx, planck, optical_depth, surf_emissivity, surf_planck):
    """radiative_transfer_lw<false> (radiative_transfer_lw.cpp:25-60) -> (flux_dn, flux_up)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, n = od.shape
    d = [ctx.to_device(np.asarray(a, dtype=np.float64)) for a in (planck, od, surf_emissivity, surf_planck)]
    dn, up = ctx.empty((nlay + 1, n)), ctx.empty((nlay + 1, n))
    radiative_transfer_lw_dev(ctx, 1, nlay, n, d[0], d[1], d[2], d[3], dn, up)
    return dn.numpy(), up.numpy()


def radiative_transfer_direct_sw(ctx, cos_sza, ssi, optical_depth):
    """radiative_transfer_direct_sw<false> (radiative_transfer_sw.cpp:24-43)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, n = od.shape
    d_ssi, d_od, dn = ctx.to_device(np.asarray(ssi, dtype=np.float64)), ctx.to_device(od), ctx.empty((nlay + 1, n))
    radiative_transfer_sw_dev(ctx, 1, nlay, n, cos_sza, d_ssi, d_od, None, dn, None)
    return dn.numpy()


def radiative_transfer_norayleigh_sw(ctx, cos_sza, ssi, optical_depth, albedo):
    """radiative_transfer_norayleigh_sw<false> (radiative_transfer_sw.cpp:47-77)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nlay, n = od.shape
    d_ssi, d_od, d_al = (ctx.to_device(np.asarray(a, dtype=np.float64)) for a in (ssi, od, albedo))
    dn, up = ctx.empty((nlay + 1, n)), ctx.empty((nlay + 1, n))
    radiative_transfer_sw_dev(ctx, 1, nlay, n, cos_sza, d_ssi, d_od, d_al, dn, up)
    return dn.numpy(), up.numpy()


def heating_rate(ctx, pressure_hl, flux_dn, flux_up=None):
    """heating_rate<false> (heating_rate.h:25-49)."""
    fd = np.ascontiguousarray(flux_dn, dtype=np.float64)
    nhl, n = fd.shape
    d_p, d_dn = ctx.to_device(np.asarray(pressure_hl, dtype=np.float64)), ctx.to_device(fd)
    d_up = None if flux_up is None else ctx.to_device(np.asarray(flux_up, dtype=np.float64))
    hr = ctx.empty((nhl - 1, n))
    heating_rate_dev(ctx, 1, nhl - 1, n, d_p, d_dn, d_up, hr)
    return hr.numpy()


# ---------------------------------------------------------------------------------------------
# Stage-1 extras, broadband solvers, interval error batches, forward cost (rows a4b, a7b, a13,
# a15, a21-a23 of SURVEY.md section 8)
# ---------------------------------------------------------------------------------------------
def _dev(ctx, a, dtype=np.float64):
    return None if a is None else ctx.to_device(np.ascontiguousarray(a, dtype=dtype), dtype)


def split_base_g_points_dev(ctx, band_wn1, band_wn2, base_wavenumber_boundary, nwav, d_wavenumber, d_ssi, ng,
                            band_number, d_g_point, solar_irradiance=None, cap=None):
    """create_look_up_table.cpp:162-224 -> (ng, band_number, solar_irradiance); d_g_point is updated in place."""
    b1, b1p = _h(band_wn1)
    b2, b2p = _h(band_wn2)
    bb, bbp = _h(np.atleast_1d(np.asarray(base_wavenumber_boundary, dtype=np.float64)))
    cap = int(cap if cap is not None else ng + bb.size)
    bn = np.zeros(cap, dtype=np.int32)
    bn[:ng] = np.asarray(band_number, dtype=np.int32)
    sol = None
    if solar_irradiance is not None:
        sol = np.zeros(cap)
        sol[:ng] = np.asarray(solar_irradiance, dtype=np.float64)
    ng_io = C.c_int(int(ng))
    check(_lib.lib().ecckd_split_base_g_points_dev(
        ctx.handle, b1.size, b1p, b2p, bb.size, bbp, nwav, _ptr(d_wavenumber), _ptr(d_ssi), cap, C.byref(ng_io),
        bn.ctypes.data, _ptr(d_g_point), None if sol is None else sol.ctypes.data))
    n = ng_io.value
    return n, bn[:n].copy(), (None if sol is None else sol[:n].copy())


def rayleigh_molar_scattering_coeff(ctx, wavenumber_cm_1):
    """rayleigh_scattering.h:23-43."""
    d_wn = _dev(ctx, wavenumber_cm_1)
    out = ctx.empty(d_wn.shape)
    check(_lib.lib().ecckd_rayleigh_molar_scattering_coeff_dev(ctx.handle, int(np.prod(d_wn.shape)), d_wn.ptr, out.ptr))
    return out.numpy()


def rayleigh_molar_scat_gpoints(ctx, wavenumber1, wavenumber2, gpoint_fraction, ssi_intervals):
    """CkdModel::calc_rayleigh_molar_scat, ckd_model.h:368-385."""
    gf = np.ascontiguousarray(gpoint_fraction, dtype=np.float64)
    ng, nbin = gf.shape
    d = [_dev(ctx, a) for a in (wavenumber1, wavenumber2, gf, ssi_intervals)]
    out = ctx.empty(ng)
    check(_lib.lib().ecckd_rayleigh_molar_scat_gpoints_dev(ctx.handle, ng, nbin, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr,
                                                          out.ptr))
    return out.numpy()


def radiative_transfer_lw_bb(ctx, planck, spectral_od, grey_od, surf_emissivity, surf_planck):
    """radiative_transfer_lw_bb<false> (radiative_transfer_lw.cpp:85-142) -> (flux_dn, flux_up)[nlay+1]."""
    od = np.ascontiguousarray(spectral_od, dtype=np.float64)
    nlay, n = od.shape
    d = [_dev(ctx, a) for a in (planck, od, grey_od, surf_emissivity, surf_planck)]
    dn, up = ctx.empty(nlay + 1), ctx.empty(nlay + 1)
    check(_lib.lib().ecckd_radiative_transfer_lw_bb_dev(ctx.handle, nlay, n, n, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr,
                                                       d[4].ptr, dn.ptr, up.ptr))
    return dn.numpy(), up.numpy()


def radiative_transfer_sw_bb(ctx, cos_sza, ssi, spectral_od, grey_od, albedo, direct_only=False):
    """radiative_transfer_{direct,norayleigh}_sw_bb<false> (radiative_transfer_sw.cpp:116-184)."""
    od = np.ascontiguousarray(spectral_od, dtype=np.float64)
    nlay, n = od.shape
    d = [_dev(ctx, a) for a in (ssi, od, grey_od)]
    dn = ctx.empty(nlay + 1)
    up = None if direct_only else ctx.empty(nlay + 1)
    check(_lib.lib().ecckd_radiative_transfer_sw_bb_dev(ctx.handle, nlay, n, n, float(cos_sza), d[0].ptr, d[1].ptr,
                                                       d[2].ptr, float(albedo), dn.ptr, _ptr(up)))
    return dn.numpy(), (None if up is None else up.numpy())


def interval_bounds_to_indices(npoints, bound1, bound2):
    """CkdEquipartition::lower_index/upper_index + the checks of calc_error (find_g_points.cpp:284-315)."""
    b1, b1p = _h(np.atleast_1d(bound1))
    b2, b2p = _h(np.atleast_1d(bound2))
    i1, i2 = np.zeros(b1.size, dtype=np.int64), np.zeros(b1.size, dtype=np.int64)
    check(_lib.lib().ecckd_interval_bounds_to_indices(int(npoints), b1.size, b1p, b2p, i1.ctypes.data, i2.ctypes.data))
    return i1, i2


def interval_error_batch_lw_dev(ctx, averaging_method, nlay, n, stride, d_pressure_hl, d_planck_hl, d_surf_emissivity,
                                d_surf_planck, d_bg_od, d_metric, d_flux_dn_surf, d_flux_up_toa, d_hr, flux_weight,
                                d_layer_weight, i1, i2, d_error, d_fit=None):
    a1, a1p = _h(i1, np.int64)
    a2, a2p = _h(i2, np.int64)
    check(_lib.lib().ecckd_interval_error_batch_lw_dev(
        ctx.handle, _method_id(averaging_method), nlay, n, stride, _ptr(d_pressure_hl), _ptr(d_planck_hl),
        _ptr(d_surf_emissivity), _ptr(d_surf_planck), _ptr(d_bg_od), _ptr(d_metric), _ptr(d_flux_dn_surf),
        _ptr(d_flux_up_toa), _ptr(d_hr), float(flux_weight), _ptr(d_layer_weight), a1.size, a1p, a2p, _ptr(d_error),
        _ptr(d_fit)))


def interval_error_batch_lw(ctx, averaging_method, pressure_hl, planck_hl, surf_emissivity, surf_planck, bg_od, metric,
                            flux_dn_surf, flux_up_toa, hr, flux_weight, layer_weight, i1, i2):
    """Batched CkdEquipartition::calc_error, longwave -> (error[nint], optical_depth_fit[nint][nlay])."""
    bg = np.ascontiguousarray(bg_od, dtype=np.float64)
    nlay, n = bg.shape
    d = [_dev(ctx, a) for a in (pressure_hl, planck_hl, surf_emissivity, surf_planck, bg, metric, flux_dn_surf,
                                flux_up_toa, hr, layer_weight)]
    nint = len(np.atleast_1d(i1))
    err, fit = ctx.empty(nint), ctx.empty((nint, nlay))
    interval_error_batch_lw_dev(ctx, averaging_method, nlay, n, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8],
                                flux_weight, d[9], i1, i2, err, fit)
    return err.numpy(), fit.numpy()


def interval_error_batch_sw_dev(ctx, averaging_method, nlay, n, stride, cos_sza, d_pressure_hl, d_ssi, albedo, d_bg_od,
                                d_metric, d_flux_dn_surf, d_flux_up_toa, d_hr, flux_weight, d_layer_weight, i1, i2,
                                d_error, d_fit=None, low=None, high=None, min_scaling=1.0, max_scaling=1.0):
    a1, a1p = _h(i1, np.int64)
    a2, a2p = _h(i2, np.int64)
    low = low or (None, None, None)
    high = high or (None, None, None)
    check(_lib.lib().ecckd_interval_error_batch_sw_dev(
        ctx.handle, _method_id(averaging_method), nlay, n, stride, float(cos_sza), _ptr(d_pressure_hl), _ptr(d_ssi),
        float(albedo), _ptr(d_bg_od), _ptr(d_metric), _ptr(d_flux_dn_surf), _ptr(d_flux_up_toa), _ptr(d_hr),
        float(flux_weight), _ptr(d_layer_weight), _ptr(low[0]), _ptr(low[1]), _ptr(low[2]), _ptr(high[0]),
        _ptr(high[1]), _ptr(high[2]), float(min_scaling), float(max_scaling), a1.size, a1p, a2p, _ptr(d_error),
        _ptr(d_fit)))


def interval_error_batch_sw(ctx, averaging_method, cos_sza, pressure_hl, ssi, albedo, bg_od, metric, flux_dn_surf,
                            flux_up_toa, hr, flux_weight, layer_weight, i1, i2, low=None, high=None, min_scaling=1.0,
                            max_scaling=1.0):
    """Batched CkdEquipartition::calc_error, shortwave; low/high = (flux_dn_surf, flux_up_toa, hr) sets."""
    bg = np.ascontiguousarray(bg_od, dtype=np.float64)
    nlay, n = bg.shape
    d = [_dev(ctx, a) for a in (pressure_hl, ssi, bg, metric, flux_dn_surf, flux_up_toa, hr, layer_weight)]
    dl = None if low is None else tuple(_dev(ctx, a) for a in low)
    dh = None if high is None else tuple(_dev(ctx, a) for a in high)
    nint = len(np.atleast_1d(i1))
    err, fit = ctx.empty(nint), ctx.empty((nint, nlay))
    interval_error_batch_sw_dev(ctx, averaging_method, nlay, n, n, cos_sza, d[0], d[1], albedo, d[2], d[3], d[4], d[5],
                                d[6], flux_weight, d[7], i1, i2, err, fit, dl, dh, min_scaling, max_scaling)
    return err.numpy(), fit.numpy()


def cost_function_ckd_lw(ctx, pressure_hl, planck_hl, surf_emissivity, surf_planck, optical_depth, flux_dn, flux_up, hr,
                         flux_weight, flux_profile_weight, broadband_weight, spectral_boundary_weight, layer_weight=None,
                         pressure_weight_power=0.5, negative_od_penalty=1.0e4, spectral_flux_dn_surf=None,
                         spectral_flux_up_toa=None, rel_dn=None, rel_up=None, band_mapping=None):
    """calc_cost_function_ckd_lw for every profile of a scene (calc_cost_function_lw.cpp:115-232 as driven by
    solve_adept.cpp:96-184) -> dict(cost_profile[nprof], cost_total, penalty)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nprof, nlay, ng = od.shape
    nband = np.asarray(flux_dn).shape[-1]
    d = [_dev(ctx, a) for a in (pressure_hl, planck_hl, surf_emissivity, surf_planck, od, flux_dn, flux_up, hr,
                                spectral_flux_dn_surf, spectral_flux_up_toa, layer_weight, rel_dn, rel_up)]
    d_bm = _dev(ctx, band_mapping, np.int32)
    cp, ct, pt = ctx.empty(nprof), ctx.empty(1), ctx.empty(1)
    check(_lib.lib().ecckd_cost_function_ckd_lw_dev(
        ctx.handle, nprof, nlay, ng, nband, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr, d[6].ptr,
        d[7].ptr, _ptr(d[8]), _ptr(d[9]), float(flux_weight), float(flux_profile_weight), float(broadband_weight),
        float(spectral_boundary_weight), _ptr(d[10]), float(pressure_weight_power), float(negative_od_penalty),
        _ptr(d[11]), _ptr(d[12]), _ptr(d_bm), cp.ptr, ct.ptr, pt.ptr))
    return dict(cost_profile=cp.numpy(), cost_total=float(ct.numpy()[0]), penalty=float(pt.numpy()[0]))


def cost_function_ckd_lw_grad(ctx, pressure_hl, planck_hl, surf_emissivity, surf_planck, optical_depth, flux_dn, flux_up, hr,
                              flux_weight, flux_profile_weight, broadband_weight, spectral_boundary_weight,
                              layer_weight=None, pressure_weight_power=0.5, negative_od_penalty=1.0e4,
                              spectral_flux_dn_surf=None, spectral_flux_up_toa=None, rel_dn=None, rel_up=None,
                              band_mapping=None):
    """Gradient of the scene's longwave cost (sum of calc_cost_function_ckd_lw over the profiles + the
    negative-optical-depth penalty, solve_adept.cpp:96-184) with respect to optical_depth[nprof][nlay][ng]: the
    hand-written reverse mode that stands in for the Adept tape (solve_adept.cpp:91, 201-203)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nprof, nlay, ng = od.shape
    nband = np.asarray(flux_dn).shape[-1]
    d = [_dev(ctx, a) for a in (pressure_hl, planck_hl, surf_emissivity, surf_planck, od, flux_dn, flux_up, hr,
                                spectral_flux_dn_surf, spectral_flux_up_toa, layer_weight, rel_dn, rel_up)]
    d_bm = _dev(ctx, band_mapping, np.int32)
    grad = ctx.empty((nprof, nlay, ng))
    check(_lib.lib().ecckd_cost_function_ckd_lw_grad_dev(
        ctx.handle, nprof, nlay, ng, nband, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr, d[6].ptr,
        d[7].ptr, _ptr(d[8]), _ptr(d[9]), float(flux_weight), float(flux_profile_weight), float(broadband_weight),
        float(spectral_boundary_weight), _ptr(d[10]), float(pressure_weight_power), float(negative_od_penalty),
        _ptr(d[11]), _ptr(d[12]), _ptr(d_bm), grad.ptr))
    return grad.numpy()


def cost_function_ckd_sw(ctx, cos_sza, pressure_hl, ssi, albedo, optical_depth, flux_dn, flux_up, hr, flux_weight,
                         flux_profile_weight, broadband_weight, layer_weight=None, pressure_weight_power=0.5,
                         negative_od_penalty=1.0e4, spectral_boundary_weights=None, spectral_flux_dn_surf=None,
                         spectral_flux_up_toa=None, rel_dn=None, rel_up=None, band_mapping=None, cost_fn_per_band=None):
    """calc_cost_function_ckd_sw for every profile of a scene (calc_cost_function_sw.cpp:115-277)."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nprof, nlay, ng = od.shape
    nband = np.asarray(flux_dn).shape[-1]
    mu = np.broadcast_to(np.asarray(cos_sza, dtype=np.float64), (nprof,))
    d = [_dev(ctx, a) for a in (mu, pressure_hl, ssi, od, flux_dn, flux_up, hr, spectral_flux_dn_surf,
                                spectral_flux_up_toa, spectral_boundary_weights, layer_weight, rel_dn, rel_up)]
    d_bm = _dev(ctx, band_mapping, np.int32)
    alb, albp = _h(np.atleast_1d(albedo))
    d_pb = _dev(ctx, cost_fn_per_band)
    cp, ct, pt = ctx.empty(nprof), ctx.empty(1), ctx.empty(1)
    check(_lib.lib().ecckd_cost_function_ckd_sw_dev(
        ctx.handle, nprof, nlay, ng, nband, d[0].ptr, d[1].ptr, d[2].ptr, alb.size, albp, d[3].ptr, d[4].ptr, d[5].ptr,
        d[6].ptr, _ptr(d[7]), _ptr(d[8]), float(flux_weight), float(flux_profile_weight), float(broadband_weight),
        _ptr(d[9]), _ptr(d[10]), float(pressure_weight_power), float(negative_od_penalty), _ptr(d[11]), _ptr(d[12]),
        _ptr(d_bm), _ptr(d_pb), cp.ptr, ct.ptr, pt.ptr))
    out = dict(cost_profile=cp.numpy(), cost_total=float(ct.numpy()[0]), penalty=float(pt.numpy()[0]))
    if d_pb is not None:
        out["cost_fn_per_band"] = d_pb.numpy()
    return out


def cost_function_ckd_sw_grad(ctx, cos_sza, pressure_hl, ssi, albedo, optical_depth, flux_dn, flux_up, hr, flux_weight,
                              flux_profile_weight, broadband_weight, layer_weight=None, pressure_weight_power=0.5,
                              negative_od_penalty=1.0e4, spectral_boundary_weights=None, spectral_flux_dn_surf=None,
                              rel_dn=None, rel_up=None, band_mapping=None):
    """Gradient of the scene's shortwave cost (calc_cost_function_sw.cpp:115-277 over the profiles + the
    negative-optical-depth penalty) with respect to optical_depth[nprof][nlay][ng]."""
    od = np.ascontiguousarray(optical_depth, dtype=np.float64)
    nprof, nlay, ng = od.shape
    nband = np.asarray(flux_dn).shape[-1]
    mu = np.broadcast_to(np.asarray(cos_sza, dtype=np.float64), (nprof,))
    d = [_dev(ctx, a) for a in (mu, pressure_hl, ssi, od, flux_dn, flux_up, hr, spectral_flux_dn_surf,
                                spectral_boundary_weights, layer_weight, rel_dn, rel_up)]
    d_bm = _dev(ctx, band_mapping, np.int32)
    alb, albp = _h(np.atleast_1d(albedo))
    grad = ctx.empty((nprof, nlay, ng))
    check(_lib.lib().ecckd_cost_function_ckd_sw_grad_dev(
        ctx.handle, nprof, nlay, ng, nband, d[0].ptr, d[1].ptr, d[2].ptr, alb.size, albp, d[3].ptr, d[4].ptr, d[5].ptr,
        d[6].ptr, _ptr(d[7]), float(flux_weight), float(flux_profile_weight), float(broadband_weight), _ptr(d[8]),
        _ptr(d[9]), float(pressure_weight_power), float(negative_od_penalty), _ptr(d[10]), _ptr(d[11]), _ptr(d_bm),
        grad.ptr))
    return grad.numpy()


def background_cost_gas_dev(ctx: Context, nx: int, ng: int, d_inv_background_shape, d_background_error, d_delta_x,
                            d_gradient=None) -> float:
    """Background (prior) term of one active gas, CkdModel<true>::calc_background_cost_function
    (ckd_model.cpp:838-866): returns the gas's cost and adds its term to d_gradient [nx][ng] (None = cost only)."""
    cost = C.c_double(0.0)
    check(_lib.lib().ecckd_background_cost_gas_dev(ctx.handle, int(nx), int(ng), _ptr(d_inv_background_shape),
                                                   _ptr(d_background_error), _ptr(d_delta_x), _ptr(d_gradient),
                                                   C.byref(cost)))
    return cost.value


def background_cost_rayleigh_dev(ctx: Context, ng: int, d_rayleigh_inv_background, d_delta_x) -> float:
    """The Rayleigh part of the background term (ckd_model.cpp:868-873): cost only, as in the reference."""
    cost = C.c_double(0.0)
    check(_lib.lib().ecckd_background_cost_rayleigh_dev(ctx.handle, int(ng), _ptr(d_rayleigh_inv_background),
                                                        _ptr(d_delta_x), C.byref(cost)))
    return cost.value


def cost_function_ckd_scenes_dev(ctx, is_sw, scenes: Sequence[dict], nlay, ng, nband, flux_weight, flux_profile_weight,
                                 broadband_weight, spectral_boundary_weight=0.0, pressure_weight_power=0.5,
                                 negative_od_penalty=1.0e4, d_band_mapping=None, d_cost_fn_per_band=None,
                                 d_cost_scene=None, d_penalty_scene=None, d_cost_total=None):
    """The forward cost of several scenes in one launch (ecckd_cost_function_ckd_scenes_dev: the scene loop of
    solve_adept.cpp:96-199).  Every scene is a dict of the fields of struct ecckd_cost_scene without the d_/h_ prefix
    (device arrays; `albedo` a host array), plus nprof."""
    arr = (_lib.CostScene * len(scenes))()
    keep = []
    for k, sc in enumerate(scenes):
        arr[k].nprof = int(sc["nprof"])
        for name, _ in _lib.CostScene._fields_[2:]:
            if name == "h_albedo":
                continue
            setattr(arr[k], name, _ptr(sc.get(name[2:])))
        if sc.get("albedo") is not None:
            a, ap = _h(np.atleast_1d(sc["albedo"]))
            keep.append(a)
            arr[k].h_albedo = ap
            arr[k].nalbedo = a.size
    check(_lib.lib().ecckd_cost_function_ckd_scenes_dev(
        ctx.handle, int(bool(is_sw)), len(scenes), C.addressof(arr), nlay, ng, nband, float(flux_weight),
        float(flux_profile_weight), float(broadband_weight), float(spectral_boundary_weight),
        float(pressure_weight_power), float(negative_od_penalty), _ptr(d_band_mapping), _ptr(d_cost_fn_per_band),
        _ptr(d_cost_scene), _ptr(d_penalty_scene), _ptr(d_cost_total)))
