"""Loader for the C-ABI shared library (include/ecckd_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``ecckd_b200/lib/libecckd_b200.so``.  There is no fallback of any kind: if the
library is missing, or no sm_100 device is usable, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libecckd_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ecckd_b200.h")

_lib = None

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_void_pp = C.POINTER(C.c_void_p)


class GasTable(C.Structure):
    """struct ecckd_gas_table (include/ecckd_b200.h)."""

    _fields_ = [
        ("conc_dependence", C.c_int32),
        ("nconc", C.c_int32),
        ("vmr_lut", C.c_void_p),
        ("reference_vmr", C.c_double),
        ("molar_abs", C.c_void_p),
    ]


class CostScene(C.Structure):
    """struct ecckd_cost_scene (include/ecckd_b200.h)."""

    _fields_ = [("nprof", C.c_int32), ("nalbedo", C.c_int32)] + [(n, C.c_void_p) for n in (
        "d_pressure_hl", "d_optical_depth", "d_flux_dn", "d_flux_up", "d_hr", "d_spectral_flux_dn_surf",
        "d_spectral_flux_up_toa", "d_layer_weight", "d_relative_ckd_flux_dn", "d_relative_ckd_flux_up", "d_planck_hl",
        "d_surf_emissivity", "d_surf_planck", "d_cos_sza", "d_ssi", "h_albedo", "d_spectral_boundary_weights",
        "d_cost_profile")]


def header_symbols(path: str = HEADER_PATH):
    """Names of every function the header declares."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecckd_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Return the loaded library, loading it on first use."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  ecckd_b200 has no CPU fallback."
            )
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i32, i64, dbl, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    L.ecckd_version.restype = C.c_int
    L.ecckd_last_error.restype = C.c_char_p
    L.ecckd_device_count.restype = C.c_int
    L.ecckd_kernel_launch_count.restype = C.c_int64
    L.ecckd_kernel_launch_count.argtypes = [vp]
    L.ecckd_averaging_method_from_string.argtypes = [C.c_char_p]
    sigs = {
        "ecckd_ctx_create": [i32, vp, c_void_pp],
        "ecckd_ctx_destroy": [vp],
        "ecckd_ctx_set_stream": [vp, vp],
        "ecckd_synchronize": [vp],
        "ecckd_device_malloc": [vp, sz, c_void_pp],
        "ecckd_device_free": [vp, vp],
        "ecckd_host_malloc_pinned": [sz, c_void_pp],
        "ecckd_host_free_pinned": [vp],
        "ecckd_memcpy_h2d": [vp, vp, vp, sz],
        "ecckd_memcpy_d2h": [vp, vp, vp, sz],
        "ecckd_debug_eval": [vp, i32, dbl, i64, vp, vp],
        "ecckd_planck_function_dev": [vp, i32, vp, i64, vp, vp, vp, i64],
        "ecckd_planck_function": [vp, i32, vp, i64, vp, vp, vp, i64],
        "ecckd_gpoint_plan_create": [vp, vp, i64, i32, c_void_pp],
        "ecckd_gpoint_plan_destroy": [vp],
        "ecckd_gpoint_plan_counts": [vp, vp],
        "ecckd_gpoint_plan_export": [vp, vp, i64, vp],
        "ecckd_gpoint_plan_counts_dev": [vp, c_void_pp],
        "ecckd_average_optical_depth_partial_dev": [vp, vp, i32, i32, vp, vp, i64, i32, vp, i64, vp, vp, vp, i32, vp,
                                                    vp],
        "ecckd_average_optical_depth_finalize_dev": [vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, i64,
                                                     vp],
        "ecckd_average_optical_depth_to_g_point_dev": [vp, vp, dbl, i32, vp, vp, vp, i64, i32, vp, i64, vp, vp, vp,
                                                       i32, vp, vp, vp, i64, vp],
        "ecckd_average_optical_depth_batch_dev": [vp, vp, i32, vp, i32, vp, vp, vp, i64, i32, vp, i64, vp, vp, vp,
                                                  i32, vp, vp, vp, i64, vp],
        "ecckd_average_optical_depth_to_g_point": [vp, i32, dbl, i32, i64, vp, vp, vp, vp, i64, vp, i64, C.c_char_p,
                                                   vp, vp, vp, i64, vp],
        "ecckd_planck_average_optical_depth_to_g_point": [vp, i32, dbl, i32, i64, vp, vp, vp, vp, i64, vp, vp, vp,
                                                          C.c_char_p, vp, vp, vp, i64, vp],
        "ecckd_planck_average_optical_depth_to_g_point_f32": [vp, i32, dbl, i32, i64, vp, vp, vp, vp, i64, vp, vp, vp,
                                                              C.c_char_p, vp, vp, vp, i64, vp],
        "ecckd_create_lut_block": [vp, i32, dbl, i32, i64, vp, vp, vp, vp, i64, vp, vp, vp, C.c_char_p, vp, vp, vp, vp,
                                   i64, vp],
        "ecckd_lut_builder_create": [vp, i32, i64, vp, vp, vp, vp, vp],
        "ecckd_lut_builder_block": [vp, dbl, i32, vp, vp, vp, vp, i64, C.c_char_p, vp, vp, vp, vp, i64, vp],
        "ecckd_lut_builder_destroy": [vp],
        "ecckd_ctx_plan_cache_stats": [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "ecckd_comm_unique_id": [vp],
        "ecckd_comm_create": [vp, i32, i32, vp, vp],
        "ecckd_comm_from_nccl": [vp, vp, i32, i32, vp],
        "ecckd_comm_destroy": [vp],
        "ecckd_average_optical_depth_exchange_dev": [vp, vp, i32, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp],
        "ecckd_partition_options_default": [vp],
        "ecckd_partition_search_create": [vp, vp, vp, vp],
        "ecckd_partition_search_destroy": [vp],
        "ecckd_partition_search_n": [vp, i32, vp, vp, i32, vp],
        "ecckd_partition_search_e": [vp, dbl, dbl, dbl, i32, vp, vp, vp, vp],
        "ecckd_partition_search_stats": [vp, vp, vp, vp],
        "ecckd_partition_search_band": [vp, i64, dbl, i32, i32, i32, vp, vp, vp, vp, vp, vp],
        "ecckd_interval_error_lw_eval": [vp, i32, vp, vp, vp],
        "ecckd_interval_error_sw_eval": [vp, i32, vp, vp, vp],
        "ecckd_ctx_set_timing": [vp, i32],
        "ecckd_ctx_reduce_kernel_time": [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
        "ecckd_merge_axpy_dev": [vp, i32, i64, vp, i64, vp, i64, dbl, vp],
        "ecckd_merge_spectra_dev": [vp, i32, i64, i32, vp, vp, vp, vp, i64],
        "ecckd_gpoint_fraction_dev": [vp, i32, i64, vp, vp, vp, i32, vp, vp, vp],
        "ecckd_ssi_intervals_dev": [vp, i64, vp, vp, i32, vp, vp, vp],
        "ecckd_planck_lut_dev": [vp, vp, vp, vp, i32, vp, vp],
        "ecckd_gpoint_sum_dev": [vp, vp, i32, vp, i64, vp],
        "ecckd_row_sum_dev": [vp, i32, i64, vp, i64, vp],
        "ecckd_lw_spectra_profile_dev": [vp, vp, i32, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp],
        "ecckd_scale_lut_scaling": [i32, i32, dbl, vp, vp, vp, vp],
        "ecckd_compact_g_points_dev": [vp, i32, i64, vp, vp, C.POINTER(C.c_int)],
        "ecckd_sort_key_lw_dev": [vp, i32, i64, vp, vp, vp, vp, i64, dbl, vp, vp],
        "ecckd_sort_key_sw_dev": [vp, i32, i64, vp, vp, i64, dbl, vp, vp],
        "ecckd_rank_bands_dev": [vp, i64, vp, vp, i32, vp, vp, vp, vp, vp],
        "ecckd_store_g_points_dev": [vp, i64, vp, i32, vp, vp, vp],
        "ecckd_overlap_g_points": [i32, i32, vp, vp, vp, i32, vp, vp, vp, C.POINTER(C.c_int)],
        "ecckd_assign_g_points_dev": [vp, i32, i64, vp, i32, i32, vp, vp, vp, vp],
        "ecckd_ckd_model_create": [vp, i32, vp, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, c_void_pp],
        "ecckd_ckd_model_destroy": [vp],
        "ecckd_ckd_calc_optical_depth_dev": [vp, vp, i32, i32, i32, vp, vp, vp, vp, i32],
        "ecckd_ckd_calc_planck_function_dev": [vp, vp, i64, vp, vp],
        "ecckd_ckd_calc_rayleigh_optical_depth_dev": [vp, vp, i32, i32, vp, vp],
        "ecckd_radiative_transfer_lw_dev": [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp],
        "ecckd_radiative_transfer_sw_dev": [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp],
        "ecckd_heating_rate_dev": [vp, i32, i32, i64, vp, vp, vp, vp],
        "ecckd_run_ckd_dev": [vp, vp, vp, dbl, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "ecckd_run_ckd": [vp, vp, vp, dbl, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "ecckd_split_base_g_points_dev": [vp, i32, vp, vp, i32, vp, i64, vp, vp, i32, C.POINTER(C.c_int), vp, vp, vp],
        "ecckd_rayleigh_molar_scattering_coeff_dev": [vp, i64, vp, vp],
        "ecckd_rayleigh_molar_scat_gpoints_dev": [vp, i32, i32, vp, vp, vp, vp, vp],
        "ecckd_radiative_transfer_lw_bb_dev": [vp, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp],
        "ecckd_radiative_transfer_sw_bb_dev": [vp, i32, i64, i64, dbl, vp, vp, vp, dbl, vp, vp],
        "ecckd_interval_bounds_to_indices": [i64, i32, vp, vp, vp, vp],
        "ecckd_interval_error_batch_lw_dev": [vp, i32, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, vp, i32,
                                              vp, vp, vp, vp],
        "ecckd_interval_error_batch_sw_dev": [vp, i32, i32, i64, i64, dbl, vp, vp, dbl, vp, vp, vp, vp, vp, dbl, vp,
                                              vp, vp, vp, vp, vp, vp, dbl, dbl, i32, vp, vp, vp, vp],
        "ecckd_calc_total_optical_depth_dev": [vp, vp, vp, i32, i32, vp, vp, vp],
        "ecckd_cost_function_ckd_lw_dev": [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl,
                                           dbl, dbl, vp, dbl, dbl, vp, vp, vp, vp, vp, vp],
        "ecckd_calc_total_optical_depth_adjoint_dev": [vp, vp, vp, i32, i32, vp, vp, vp, vp, i32],
        "ecckd_cost_function_ckd_lw_grad_dev": [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl,
                                                dbl, dbl, vp, dbl, dbl, vp, vp, vp, vp],
        "ecckd_cost_function_ckd_sw_grad_dev": [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, dbl, dbl,
                                                dbl, vp, vp, dbl, dbl, vp, vp, vp, vp],
        "ecckd_cost_function_ckd_sw_dev": [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, dbl,
                                           dbl, dbl, vp, vp, dbl, dbl, vp, vp, vp, vp, vp, vp, vp],
        "ecckd_cost_function_ckd_scenes_dev": [vp, i32, i32, vp, i32, i32, i32, dbl, dbl, dbl, dbl, dbl, dbl, vp, vp, vp,
                                               vp, vp],
        "ecckd_background_cost_gas_dev": [vp, i32, i32, vp, vp, vp, vp, vp],
        "ecckd_background_cost_rayleigh_dev": [vp, i32, vp, vp, vp],
    }
    for name, argtypes in sigs.items():
        if hasattr(L, name):  # symbol-completeness is asserted by tests/test_abi.py
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int


class PartitionOptions(C.Structure):
    """ecckd_partition_options"""
    _fields_ = [("partition_tolerance", C.c_double), ("next_bound_error_tolerance", C.c_double), ("resolution", C.c_double),
                ("partition_max_iterations", C.c_int), ("line_search_max_iterations", C.c_int),
                ("next_bound_max_iterations", C.c_int), ("minimize_frac_range", C.c_int), ("cubic_interpolation", C.c_int),
                ("speculative_line_search", C.c_int)]


class IntervalErrorLwArgs(C.Structure):
    """ecckd_interval_error_lw_args"""
    _fields_ = [("ctx", C.c_void_p), ("averaging_method", C.c_int), ("nlay", C.c_int), ("n", C.c_int64), ("stride", C.c_int64)] + [
        (k, C.c_void_p) for k in ("d_pressure_hl", "d_planck_hl", "d_surf_emissivity", "d_surf_planck", "d_bg_optical_depth",
                                  "d_metric", "d_flux_dn_surf", "d_flux_up_toa", "d_hr")] + [
        ("flux_weight", C.c_double), ("d_layer_weight", C.c_void_p)]


class IntervalErrorSwArgs(C.Structure):
    """ecckd_interval_error_sw_args"""
    _fields_ = [("ctx", C.c_void_p), ("averaging_method", C.c_int), ("nlay", C.c_int), ("n", C.c_int64), ("stride", C.c_int64),
                ("cos_sza", C.c_double), ("d_pressure_hl", C.c_void_p), ("d_ssi", C.c_void_p), ("albedo", C.c_double)] + [
        (k, C.c_void_p) for k in ("d_bg_optical_depth", "d_metric", "d_flux_dn_surf", "d_flux_up_toa", "d_hr")] + [
        ("flux_weight", C.c_double)] + [
        (k, C.c_void_p) for k in ("d_layer_weight", "d_flux_dn_surf_low", "d_flux_up_toa_low", "d_hr_low",
                                  "d_flux_dn_surf_high", "d_flux_up_toa_high", "d_hr_high")] + [
        ("min_scaling", C.c_double), ("max_scaling", C.c_double)]


INTERVAL_ERROR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_double))


class EcckdError(RuntimeError):
    """A non-zero status from the C ABI; .code carries the ESA exit code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


def check(rc: int):
    if rc != 0:
        msg = lib().ecckd_last_error()
        raise EcckdError(rc, msg.decode() if msg else "")
