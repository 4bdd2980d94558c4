// Reference-side bindings: the bodies a maintainer of ecCKD puts in place of the hot functions so that the
// tools run on libecckd_b200.so (INTEGRATION.md sections 1-3 and 6, here as code that is compiled and tested).
//
// This translation unit is compiled against the REFERENCE's own headers (src/ecckd, src/include) and Adept
// (the real library, or oracle/adept_shim in this repository's test build) and is linked into the tools
//   * INSTEAD of  src/ecckd/average_optical_depth.cpp   (average_optical_depth_to_g_point, :22-197)
//   * INSTEAD of  src/ecckd/planck_function.cpp         (planck_function, :22-54)
//   * providing   radiative_transfer_lw<false>          (radiative_transfer_lw.cpp:25-60; the stock unit is
//                 compiled with -Dradiative_transfer_lw=radiative_transfer_lw_stock so that its other
//                 functions, radiative_transfer_lw_bb and the <true> instantiations, stay available)
// Nothing else in the tools changes: config grammar, file variables, logging, exit codes.
// oracle/build_ref_tools.sh builds both variants; tests/test_cxx_consumer_gpu.py runs the unmodified mains
// of create_look_up_table and run_ckd in both and compares what they write.
#include <cstdint>
#include <string>
#include <vector>

#include <adept_arrays.h>

#include "Error.h"
#include "Logging.h"
#include "average_optical_depth.h"
#include "planck_function.h"
#include "radiative_transfer_lw.h"

#include "ecckd_b200.h"

using adept::Matrix;
using adept::Real;
using adept::Vector;

namespace {

ecckd_ctx* gpu() {
  static ecckd_ctx* ctx = nullptr;
  if (!ctx && ecckd_ctx_create(0, nullptr, &ctx) != ECCKD_OK) {
    ERROR << ecckd_last_error();
    THROW(UNEXPECTED_EXCEPTION);
  }
  return ctx;
}

void gpu_check(int rc) {  // the C ABI returns the integer codes the reference throws (EsaExitCodes.h)
  if (rc != ECCKD_OK) {
    ERROR << ecckd_last_error();
    THROW(rc);
  }
}

// A matrix whose rows are contiguous: the argument itself, or a packed copy of it.
Matrix packed_rows(const Matrix& a) {
  if (a.empty() || a.offset(1) == 1) return a;  // shallow link
  Matrix c(a.size(0), a.size(1));
  c = a;
  return c;
}

Vector packed(const Vector& a) {
  if (a.empty() || a.offset(0) == 1) return a;
  Vector c(a.size());
  c = a;
  return c;
}

}  // namespace

void average_optical_depth_to_g_point(int ng, Real reference_surface_vmr, const Vector& pressure_fl,
                                      const Vector& pressure_hl, const Vector& g_point, const Matrix& optical_depth,
                                      const Matrix& planck_fl, const std::string& averaging_method, Matrix molar_abs,
                                      Matrix min_molar_abs, Matrix max_molar_abs) {
  // g_point arrives as a Vector of doubles holding integers (create_look_up_table.cpp:327)
  const int nwav = g_point.size();
  std::vector<int32_t> g(nwav);
  for (int i = 0; i < nwav; ++i) g[i] = static_cast<int32_t>(g_point(i));
  std::vector<int32_t> warn(ng, 0);
  const Vector p_fl = packed(pressure_fl), p_hl = packed(pressure_hl);
  const Matrix od = packed_rows(optical_depth), w = packed_rows(planck_fl);
  const bool mm = !min_molar_abs.empty() && !max_molar_abs.empty();
  // outputs are slices such as molar_abs(icol,__,__), written in place through their row stride
  const bool direct = molar_abs.offset(1) == 1 && (!mm || (min_molar_abs.offset(1) == 1 && max_molar_abs.offset(1) == 1 &&
                                                           min_molar_abs.offset(0) == molar_abs.offset(0) &&
                                                           max_molar_abs.offset(0) == molar_abs.offset(0)));
  Matrix out = direct ? molar_abs : Matrix(molar_abs.size(0), ng);
  Matrix out_min = !mm ? Matrix() : (direct ? min_molar_abs : Matrix(molar_abs.size(0), ng));
  Matrix out_max = !mm ? Matrix() : (direct ? max_molar_abs : Matrix(molar_abs.size(0), ng));
  gpu_check(ecckd_average_optical_depth_to_g_point(
      gpu(), ng, reference_surface_vmr, od.size(0), od.size(1), p_fl.const_data(), p_hl.const_data(), g.data(),
      od.const_data(), od.offset(0), w.const_data(), w.size(0) == 1 ? 0 : w.offset(0), averaging_method.c_str(),
      out.data(), mm ? out_min.data() : nullptr, mm ? out_max.data() : nullptr, out.offset(0), warn.data()));
  if (!direct) {
    molar_abs = out;
    if (mm) {
      min_molar_abs = out_min;
      max_molar_abs = out_max;
    }
  }
  for (int ig = 0; ig < ng; ++ig) {  // the WARNINGs of average_optical_depth.cpp:136-137, 154, 159
    if (warn[ig] & ECCKD_WARN_EMPTY_G) {
      WARNING << "No wavenumbers with g_point == " << ig << ": skipping";
      ENDWARNING;
    }
    if (warn[ig] & ECCKD_WARN_MIN_GT_FIT) {
      WARNING << "G-point " << ig << ": min optical depth > average optical depth: correcting";
      ENDWARNING;
    }
    if (warn[ig] & ECCKD_WARN_MIN_GE_MAX) {
      WARNING << "G-point " << ig << ": min optical depth >= max optical depth";
      ENDWARNING;
    }
  }
}

void planck_function(const Vector& temperature, const Vector& wavenumber_cm_1, const Vector& d_wavenumber_cm_1,
                     Matrix ans) {
  if (wavenumber_cm_1.size() != d_wavenumber_cm_1.size()) {  // planck_function.cpp:35-38
    ERROR << "wavenumber_cm_1 and d_wavenumber_cm_1 must have the same length";
    THROW(PARAMETER_ERROR);
  }
  const Vector t = packed(temperature), wn = packed(wavenumber_cm_1), dwn = packed(d_wavenumber_cm_1);
  const bool direct = ans.offset(1) == 1;
  Matrix out = direct ? ans : Matrix(ans.size(0), ans.size(1));
  if (ans.size(0) != t.size() || ans.size(1) != wn.size()) {  // :40-44
    ERROR << "Planck function output must be dimensioned (temperature, wavenumber)";
    THROW(PARAMETER_ERROR);
  }
  gpu_check(ecckd_planck_function(gpu(), t.size(), t.const_data(), wn.size(), wn.const_data(), dwn.const_data(),
                                  out.data(), out.offset(0)));
  if (!direct) ans = out;
}

template <>
void radiative_transfer_lw<false>(const Matrix& planck, const Matrix& optical_depth, const Vector& surf_emissivity,
                                  const Vector& surf_planck, Matrix flux_dn, Matrix flux_up) {
  const int nlay = optical_depth.size(0), n = optical_depth.size(1);
  const Matrix pl = packed_rows(planck), od = packed_rows(optical_depth);
  const Vector em = packed(surf_emissivity), sp = packed(surf_planck);
  // device staging through the ABI's own memory helpers: one column of n spectral points
  ecckd_ctx* ctx = gpu();
  const size_t n_hl = (size_t)(nlay + 1) * n, n_fl = (size_t)nlay * n;
  double* d = nullptr;
  gpu_check(ecckd_device_malloc(ctx, (2 * n_hl + n_fl + 2 * (size_t)n + n_hl) * sizeof(double), (void**)&d));
  double *d_pl = d, *d_od = d_pl + n_hl, *d_em = d_od + n_fl, *d_sp = d_em + n, *d_dn = d_sp + n, *d_up = d_dn + n_hl;
  int rc = ECCKD_OK;
  for (int i = 0; i <= nlay && rc == ECCKD_OK; ++i)  // row by row: the arguments may have padded rows
    rc = ecckd_memcpy_h2d(ctx, d_pl + (size_t)i * n, pl.const_data() + (size_t)i * pl.offset(0), n * sizeof(double));
  for (int i = 0; i < nlay && rc == ECCKD_OK; ++i)
    rc = ecckd_memcpy_h2d(ctx, d_od + (size_t)i * n, od.const_data() + (size_t)i * od.offset(0), n * sizeof(double));
  if (rc == ECCKD_OK) rc = ecckd_memcpy_h2d(ctx, d_em, em.const_data(), n * sizeof(double));
  if (rc == ECCKD_OK) rc = ecckd_memcpy_h2d(ctx, d_sp, sp.const_data(), n * sizeof(double));
  if (rc == ECCKD_OK) rc = ecckd_radiative_transfer_lw_dev(ctx, 1, nlay, n, d_pl, d_od, d_em, d_sp, d_dn, d_up);
  Matrix dn(nlay + 1, n), up(nlay + 1, n);
  if (rc == ECCKD_OK) rc = ecckd_memcpy_d2h(ctx, dn.data(), d_dn, n_hl * sizeof(double));
  if (rc == ECCKD_OK) rc = ecckd_memcpy_d2h(ctx, up.data(), d_up, n_hl * sizeof(double));
  ecckd_device_free(ctx, d);
  gpu_check(rc);
  flux_dn = dn;  // element-wise copies into the caller's (possibly sliced) arrays
  flux_up = up;
}

// The differentiated instantiation stays with the stock unit (the reverse-mode tape is Adept's): the stock
// radiative_transfer_lw.cpp is compiled with -Dradiative_transfer_lw=radiative_transfer_lw_stock, so its
// <true> body is reached through this forwarder.
template <bool IsActive>
void radiative_transfer_lw_stock(const Matrix& planck, const adept::Array<2, adept::Real, IsActive>& optical_depth,
                                 const Vector& surf_emissivity, const Vector& surf_planck,
                                 adept::Array<2, adept::Real, IsActive> flux_dn,
                                 adept::Array<2, adept::Real, IsActive> flux_up);
template <>
void radiative_transfer_lw<true>(const Matrix& planck, const adept::Array<2, adept::Real, true>& optical_depth,
                                 const Vector& surf_emissivity, const Vector& surf_planck,
                                 adept::Array<2, adept::Real, true> flux_dn, adept::Array<2, adept::Real, true> flux_up) {
  radiative_transfer_lw_stock<true>(planck, optical_depth, surf_emissivity, surf_planck, flux_dn, flux_up);
}
