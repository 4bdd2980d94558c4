// ref_glue.cpp - C ABI over the reference's OWN functions.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).
//
// This file contains no restatement of any algorithm: every entry point below wraps
// the caller's raw arrays as Adept arrays and calls the function that the reference
// itself defines in the translation unit compiled, unmodified and in place, from
// /root/reference/src/ecckd by oracle/build_ref_shim.sh (against the Adept API shim in
// oracle/adept_shim/ - Adept itself is not installable here).  The entry points carry
// the same names and signatures as the restated oracle (ecckd_oracle.cpp), so
// tests/oracle_lib.py can load either library and tests/test_reference_sources.py can
// hold the restatement against the reference's code on the same inputs.
//
// Result: oracle/_ref/libecckd_ref.so  (git-ignored, travels to the GPU box).

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <adept_arrays.h>

#include "Error.h"
#include "average_optical_depth.h"
#include "calc_cost_function_lw.h"
#include "calc_cost_function_sw.h"
#include "heating_rate.h"
#include "planck_function.h"
#include "radiative_transfer_lw.h"
#include "radiative_transfer_sw.h"
#include "rayleigh_scattering.h"
#include "single_gas_data.h"

using namespace adept;

#define REF_API extern "C" __attribute__((visibility("default")))

// ---- the few symbols of src/tools/{Error,Logging}.cpp that the compiled units reference
// (those files need the autoconf-generated config.h and the C logging API, so they are not built).
bool trace_exceptions_ = false;
static int g_warning_count = 0;
static LogStream g_stream(LOG_LEVEL_INFO);
LogStream::~LogStream() {}
void LogStream::flush() { ss_.str(std::string()); }
LogStream& log_stream(LogLevel, const char*, const char*, int) {
  g_stream.active_line_ = false;  // drop the text; warnings are counted in flush_stream
  return g_stream;
}
void flush_stream(LogLevel log_level) {
  if (log_level == LOG_LEVEL_WARNING) ++g_warning_count;
}

REF_API int ref_warning_count(int reset) {
  const int n = g_warning_count;
  if (reset) g_warning_count = 0;
  return n;
}

REF_API const char* ref_build_info() {
  return "reference sources compiled in place: average_optical_depth.cpp planck_function.cpp "
         "radiative_transfer_lw.cpp radiative_transfer_sw.cpp single_gas_data.cpp calc_cost_function_lw.cpp "
         "calc_cost_function_sw.cpp (+ heating_rate.h rayleigh_scattering.h) against oracle/adept_shim";
}

// ---- wrapping raw memory
static Vector vec(const double* p, int64_t n) {
  if (!p) return Vector();
  int d[1] = {(int)n};
  long s[1] = {1};
  return Vector::wrap(const_cast<double*>(p), d, s);
}
static Matrix mat(const double* p, int64_t n0, int64_t n1, int64_t stride = -1) {
  if (!p) return Matrix();
  int d[2] = {(int)n0, (int)n1};
  long s[2] = {(long)(stride < 0 ? n1 : stride), 1};
  return Matrix::wrap(const_cast<double*>(p), d, s);
}
static aMatrix amat(const double* p, int64_t n0, int64_t n1) {
  int d[2] = {(int)n0, (int)n1};
  long s[2] = {(long)n1, 1};
  return aMatrix::wrap(const_cast<double*>(p), d, s);
}
static intVector ivec(const int32_t* p, int64_t n) {
  if (!p) return intVector();
  int d[1] = {(int)n};
  long s[1] = {1};
  return intVector::wrap(const_cast<int32_t*>(p), d, s);
}

static const char* const kMethods[] = {"linear",          "transmission", "transmission-2", "transmission-3",
                                       "transmission-10", "square-root",  "logarithmic",    "hybrid-logarithmic-transmission-3"};

// Run `body`, mapping what the reference throws to the ESA exit codes it would exit with.
template <typename F>
static int guarded(F body) {
  try {
    body();
  } catch (int code) {  // THROW(PARAMETER_ERROR) etc. (Logging.h:108-117)
    return code;
  } catch (const adept::size_mismatch&) {
    return -10;  // Adept would raise adept::size_mismatch
  } catch (const adept::index_out_of_bounds&) {
    return -11;
  }
  return 0;
}

// planck_function.cpp:22-54
REF_API int orc_planck_function(int nt, const double* temperature, int64_t nwav, const double* wavenumber_cm_1,
                                const double* d_wavenumber_cm_1, double* ans, int64_t ans_stride) {
  return guarded([&] {
    planck_function(vec(temperature, nt), vec(wavenumber_cm_1, nwav), vec(d_wavenumber_cm_1, nwav),
                    mat(ans, nt, nwav, ans_stride));
  });
}

// average_optical_depth.cpp:22-197.  `weight` with w_stride == 0 is one row shared by all layers
// (create_look_up_table.cpp:335 replicates ssi the same way).  The per-g warning flags are not
// observable through the reference's interface; warn[] is left untouched and the number of
// WARNING ... ENDWARNING blocks executed is available through ref_warning_count().
REF_API int orc_average_optical_depth_to_g_point(int ng, double reference_surface_vmr, int nlay, int64_t nwav,
                                                 const double* pressure_fl, const double* pressure_hl,
                                                 const int32_t* g_point, const double* od, int64_t od_stride,
                                                 const double* weight, int64_t w_stride, int method, double* molar_abs,
                                                 double* min_molar_abs, double* max_molar_abs, int64_t out_stride,
                                                 int32_t* /*warn*/) {
  return guarded([&] {
    Vector gp(nwav);  // the reference passes g_point as a Vector of doubles (create_look_up_table.cpp:327)
    for (int64_t i = 0; i < nwav; ++i) gp((int)i) = (double)g_point[i];
    Matrix planck;
    if (w_stride == 0) {
      planck.resize(nlay, (int)nwav);
      for (int iz = 0; iz < nlay; ++iz)
        for (int64_t i = 0; i < nwav; ++i) planck(iz, (int)i) = weight[i];
    } else {
      planck = mat(weight, nlay, nwav, w_stride);
    }
    const std::string name = (method >= 0 && method <= 7) ? kMethods[method] : "no-such-method";
    average_optical_depth_to_g_point(ng, reference_surface_vmr, vec(pressure_fl, nlay), vec(pressure_hl, nlay + 1), gp,
                                     mat(od, nlay, nwav, od_stride), planck, name, mat(molar_abs, nlay, ng, out_stride),
                                     mat(min_molar_abs, nlay, ng, out_stride), mat(max_molar_abs, nlay, ng, out_stride));
  });
}

// radiative_transfer_lw.cpp:25-60
REF_API int orc_radiative_transfer_lw(int nlay, int64_t n, const double* planck, const double* od,
                                      const double* surf_emissivity, const double* surf_planck, double* flux_dn,
                                      double* flux_up) {
  return guarded([&] {
    radiative_transfer_lw<false>(mat(planck, nlay + 1, n), mat(od, nlay, n), vec(surf_emissivity, n), vec(surf_planck, n),
                                 mat(flux_dn, nlay + 1, n), mat(flux_up, nlay + 1, n));
  });
}

// radiative_transfer_lw.cpp:85-142
REF_API int orc_radiative_transfer_lw_bb(int nlay, int64_t n, const double* planck, const double* spectral_od,
                                         const double* grey_od, const double* surf_emissivity, const double* surf_planck,
                                         double* flux_dn, double* flux_up) {
  return guarded([&] {
    radiative_transfer_lw_bb<false>(mat(planck, nlay + 1, n), mat(spectral_od, nlay, n), vec(grey_od, nlay),
                                    vec(surf_emissivity, n), vec(surf_planck, n), vec(flux_dn, nlay + 1),
                                    vec(flux_up, nlay + 1));
  });
}

// radiative_transfer_sw.cpp:24-43
REF_API int orc_radiative_transfer_direct_sw(int nlay, int64_t n, double cos_sza, const double* ssi, const double* od,
                                             double* flux_dn) {
  return guarded(
      [&] { radiative_transfer_direct_sw<false>(cos_sza, vec(ssi, n), mat(od, nlay, n), mat(flux_dn, nlay + 1, n)); });
}

// radiative_transfer_sw.cpp:47-77
REF_API int orc_radiative_transfer_norayleigh_sw(int nlay, int64_t n, double cos_sza, const double* ssi, const double* od,
                                                 const double* albedo, double* flux_dn, double* flux_up) {
  return guarded([&] {
    radiative_transfer_norayleigh_sw<false>(cos_sza, vec(ssi, n), mat(od, nlay, n), vec(albedo, n),
                                            mat(flux_dn, nlay + 1, n), mat(flux_up, nlay + 1, n));
  });
}

// radiative_transfer_sw.cpp:116-141 (flux_up == NULL), :145-184
REF_API int orc_radiative_transfer_sw_bb(int nlay, int64_t n, double cos_sza, const double* ssi, const double* spectral_od,
                                         const double* grey_od, double albedo, double* flux_dn, double* flux_up) {
  return guarded([&] {
    if (flux_up) {
      radiative_transfer_norayleigh_sw_bb<false>(cos_sza, vec(ssi, n), mat(spectral_od, nlay, n), vec(grey_od, nlay), albedo,
                                                 vec(flux_dn, nlay + 1), vec(flux_up, nlay + 1));
    } else {
      radiative_transfer_direct_sw_bb<false>(cos_sza, vec(ssi, n), mat(spectral_od, nlay, n), vec(grey_od, nlay),
                                             vec(flux_dn, nlay + 1));
    }
  });
}

// heating_rate.h:25-49
REF_API int orc_heating_rate(int nlay, int64_t n, const double* pressure_hl, const double* flux_dn, const double* flux_up,
                             double* hr) {
  return guarded([&] {
    heating_rate<false>(vec(pressure_hl, nlay + 1), mat(flux_dn, nlay + 1, n), mat(flux_up, nlay + 1, n), mat(hr, nlay, n));
  });
}

// heating_rate.h:51-72
REF_API int orc_heating_rate_single(int nlay, const double* pressure_hl, const double* flux_dn, const double* flux_up,
                                    double* hr) {
  return guarded([&] {
    heating_rate_single<false>(vec(pressure_hl, nlay + 1), vec(flux_dn, nlay + 1), vec(flux_up, nlay + 1), vec(hr, nlay));
  });
}

// rayleigh_scattering.h:23-43
REF_API int orc_rayleigh_molar_scattering_coeff(int64_t nwav, const double* wavenumber_cm_1, double* scat_coeff) {
  return guarded([&] {
    Vector r = rayleigh_molar_scattering_coeff(vec(wavenumber_cm_1, nwav));
    for (int64_t i = 0; i < nwav; ++i) scat_coeff[i] = r((int)i);
  });
}

// SingleGasData::store_g_points, single_gas_data.h:56-62.  The reference sizes g_point as
// maxval(rank)+1, so rank must be a permutation of 0..nwav-1 for the output to be nwav long.
REF_API int orc_store_g_points(int64_t nwav, const int32_t* rank, int ng, const int32_t* rank1, const int32_t* rank2,
                               int32_t* g_point) {
  return guarded([&] {
    SingleGasData gas;
    gas.rank1 = ivec(rank1, ng);
    gas.rank2 = ivec(rank2, ng);
    gas.store_g_points(ivec(rank, nwav));
    if (gas.g_point.size() != (int)nwav) throw adept::size_mismatch("rank is not a permutation");
    for (int64_t i = 0; i < nwav; ++i) g_point[i] = gas.g_point((int)i);
  });
}

// overlap_g_points, single_gas_data.cpp:23-124 (arguments as in ecckd_oracle.cpp)
REF_API int orc_overlap_g_points(int ngas, int nband, const int32_t* n_g_points, const double* sorting_variable,
                                 const int32_t* gas_offset, int cap, int32_t* g_min, int32_t* g_max,
                                 int32_t* band_number) {
  int ng = 0;
  const int rc = guarded([&] {
    std::vector<SingleGasData> gas(ngas);
    for (int igas = 0; igas < ngas; ++igas) {
      gas[igas].molecule = gas[igas].Molecule = "gas" + std::to_string(igas);
      gas[igas].n_g_points = ivec(n_g_points + igas * nband, nband);
      int ngi = 0;
      for (int ib = 0; ib < nband; ++ib) ngi += n_g_points[igas * nband + ib];
      gas[igas].sorting_variable = vec(sorting_variable + gas_offset[igas], ngi);
    }
    intVector bn;
    ng = overlap_g_points(gas, bn);
    if (ng > cap) { ng = -2; return; }
    for (int ig = 0; ig < ng; ++ig) band_number[ig] = bn(ig);
    for (int igas = 0; igas < ngas; ++igas)
      for (int ig = 0; ig < ng; ++ig) {
        g_min[igas * cap + ig] = gas[igas].g_min(ig);
        g_max[igas * cap + ig] = gas[igas].g_max(ig);
      }
  });
  return rc == 0 ? ng : -1;
}

// calc_cost_function_lw, calc_cost_function_lw.cpp:23-110: RMS error of one candidate g-point interval.
// All spectral arrays hold the n wavenumbers of the interval (the caller of the reference passes
// range(ibound1,ibound2) slices, find_g_points.cpp:318-328).
REF_API int orc_calc_cost_function_lw(int nlay, int64_t n, const double* pressure_hl, const double* planck_hl,
                                      const double* surf_emissivity, const double* surf_planck,
                                      const double* bg_optical_depth, const double* optical_depth_fit,
                                      const double* flux_dn_surf, const double* flux_up_toa, const double* hr,
                                      double flux_weight, const double* layer_weight, double* cost) {
  return guarded([&] {
    *cost = calc_cost_function_lw(vec(pressure_hl, nlay + 1), mat(planck_hl, nlay + 1, n), vec(surf_emissivity, n),
                                  vec(surf_planck, n), mat(bg_optical_depth, nlay, n), vec(optical_depth_fit, nlay),
                                  vec(flux_dn_surf, n), vec(flux_up_toa, n), mat(hr, nlay, n), flux_weight,
                                  vec(layer_weight, nlay), intVector());
  });
}

// calc_cost_function_sw, calc_cost_function_sw.cpp:20-110
REF_API int orc_calc_cost_function_sw(int nlay, int64_t n, double cos_sza, const double* pressure_hl, const double* ssi,
                                      double albedo, const double* bg_optical_depth, const double* optical_depth_fit,
                                      const double* flux_dn_surf, const double* flux_up_toa, const double* hr,
                                      double flux_weight, const double* layer_weight, double* cost) {
  return guarded([&] {
    *cost = calc_cost_function_sw(cos_sza, vec(pressure_hl, nlay + 1), vec(ssi, n), albedo, mat(bg_optical_depth, nlay, n),
                                  vec(optical_depth_fit, nlay), vec(flux_dn_surf, n), vec(flux_up_toa, n), mat(hr, nlay, n),
                                  flux_weight, vec(layer_weight, nlay), intVector(), false);
  });
}

// calc_cost_function_ckd_lw, calc_cost_function_lw.cpp:115-232 (forward value).
// band_mapping == NULL: nband = ng and surf_emissivity has ng entries; else surf_emissivity has
// nband entries.  Truth arrays flux_dn/flux_up[nlay+1][nband], hr[nlay][nband].
REF_API int orc_calc_cost_function_ckd_lw(int nlay, int ng, int nband, const double* pressure_hl, const double* planck_hl,
                                          const double* surf_emissivity, const double* surf_planck,
                                          const double* optical_depth, const double* flux_dn, const double* flux_up,
                                          const double* hr, const double* spectral_flux_dn_surf,
                                          const double* spectral_flux_up_toa, double flux_weight,
                                          double flux_profile_weight, double broadband_weight,
                                          double spectral_boundary_weight, const double* layer_weight,
                                          const double* relative_ckd_flux_dn, const double* relative_ckd_flux_up,
                                          const int32_t* band_mapping, double* cost) {
  return guarded([&] {
    Matrix rel_dn = mat(relative_ckd_flux_dn, nlay + 1, ng), rel_up = mat(relative_ckd_flux_up, nlay + 1, ng);
    *cost = calc_cost_function_ckd_lw(
        vec(pressure_hl, nlay + 1), mat(planck_hl, nlay + 1, ng), vec(surf_emissivity, band_mapping ? nband : ng),
        vec(surf_planck, ng), amat(optical_depth, nlay, ng), mat(flux_dn, nlay + 1, nband), mat(flux_up, nlay + 1, nband),
        mat(hr, nlay, nband), vec(spectral_flux_dn_surf, ng), vec(spectral_flux_up_toa, ng), flux_weight,
        flux_profile_weight, broadband_weight, spectral_boundary_weight, vec(layer_weight, nlay),
        relative_ckd_flux_dn ? &rel_dn : 0, relative_ckd_flux_up ? &rel_up : 0, ivec(band_mapping, ng));
  });
}

// calc_cost_function_ckd_sw, calc_cost_function_sw.cpp:115-277 (forward value).
// albedo[nalbedo] is indexed by band_mapping; cost_fn_per_band[nband] (may be NULL) is added to.
REF_API int orc_calc_cost_function_ckd_sw(int nlay, int ng, int nband, double cos_sza, const double* pressure_hl,
                                          const double* ssi, int nalbedo, const double* albedo,
                                          const double* optical_depth, const double* flux_dn, const double* flux_up,
                                          const double* hr, const double* spectral_flux_dn_surf,
                                          const double* spectral_flux_up_toa, double flux_weight,
                                          double flux_profile_weight, double broadband_weight,
                                          const double* spectral_boundary_weights, const double* layer_weight,
                                          const double* relative_ckd_flux_dn, const double* relative_ckd_flux_up,
                                          const int32_t* band_mapping, double* cost_fn_per_band, double* cost) {
  return guarded([&] {
    Matrix rel_dn = mat(relative_ckd_flux_dn, nlay + 1, ng), rel_up = mat(relative_ckd_flux_up, nlay + 1, ng);
    *cost = calc_cost_function_ckd_sw(
        cos_sza, vec(pressure_hl, nlay + 1), vec(ssi, ng), vec(albedo, nalbedo), amat(optical_depth, nlay, ng),
        mat(flux_dn, nlay + 1, nband), mat(flux_up, nlay + 1, nband), mat(hr, nlay, nband), vec(spectral_flux_dn_surf, ng),
        vec(spectral_flux_up_toa, ng), flux_weight, flux_profile_weight, broadband_weight,
        vec(spectral_boundary_weights, ng), vec(layer_weight, nlay), relative_ckd_flux_dn ? &rel_dn : 0,
        relative_ckd_flux_up ? &rel_up : 0, ivec(band_mapping, ng), vec(cost_fn_per_band, nband));
  });
}
