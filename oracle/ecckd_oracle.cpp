// ecckd_oracle.cpp - CPU restatement of ecCKD's data-parallel hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (ecckd_b200/, include/)
// may call, link or import this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs do.
//
// Every function restates, in plain loops over raw row-major arrays, one
// function of the reference (ecmwf-ifs/ecckd v1.6, paths relative to
// /root/reference) and cites the file:line it follows.  The reference uses the
// Adept array library (not vendored, not installable here), so its own
// translation units cannot be compiled directly; see oracle/Makefile target
// `_ref` for the build of the reference's *own sources* against an
// Adept-API shim, which is used to cross-check this restatement.
//
// PARITY STATUS: the reference ships no tests, golden vectors or fixtures for
// this path, so it pins nothing itself.  This oracle is
// pinned by (i) analytic known-answer tests, (ii) an independent NumPy
// restatement (oracle/oracle_np.py) and (iii) the reference sources built
// against the shim (oracle/_ref, tests/test_reference_sources.py, tests/test_golden_reference.py).
//
// Summation order: the reference's order is Adept-internal and unspecified;
// this file sums sequentially in ascending index order.
//
// It deliberately keeps the reference's inefficiencies where they determine
// the CPU cost (per-g-point rescans of the whole spectrum, O(ng*nbin*nwav)
// gpoint_fraction, serial columns) because it doubles as the timed CPU
// baseline.  OpenMP sites mirror the reference's four `omp parallel for`.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

// src/ecckd/constants.h:22-26
static const double ACCEL_GRAVITY = 9.80665;
static const double SPECIFIC_HEAT_AIR = 1004.0;
static const double LW_DIFFUSIVITY = 1.66;
static const double MOLAR_MASS_DRY_AIR = 28.970;

// Error codes: src/include/EsaExitCodes.h:16-51
enum { ORC_OK = 0, ORC_PARAMETER_ERROR = 147, ORC_PROCESSING_ERROR = 148 };

// averaging_method ids (strings at average_optical_depth.cpp:43-129,
// find_g_points.cpp:171)
enum {
  AVG_LINEAR = 0,
  AVG_TRANSMISSION = 1,
  AVG_TRANSMISSION_2 = 2,
  AVG_TRANSMISSION_3 = 3,
  AVG_TRANSMISSION_10 = 4,
  AVG_SQUARE_ROOT = 5,
  AVG_LOGARITHMIC = 6,
  AVG_HYBRID_LOG_TRANS3 = 7,
  AVG_TOTAL_TRANSMISSION = 8
};

// warning bit flags returned per g point (average_optical_depth.cpp:136,154,159)
enum { WARN_EMPTY_G = 1, WARN_MIN_GT_FIT = 2, WARN_MIN_GE_MAX = 4 };

// ---------------------------------------------------------------------------
// planck_function: src/ecckd/planck_function.cpp:22-54
// ans[it][iw] = prefactor[iw] / (exp((h/k)*(freq[iw]/T[it])) - 1)
ORC_API int orc_planck_function(int nt, const double* temperature, int64_t nwav,
                                const double* wavenumber_cm_1,
                                const double* d_wavenumber_cm_1, double* ans,
                                int64_t ans_stride) {
  const double h = 6.62606896e-34;
  const double c = 2.99792458e8;
  const double k = 1.3806504e-23;
  const double inv_cm_2_Hz = 100.0 * c;
  const double pi = 3.14159265358979323846;
  std::vector<double> freq(nwav), prefactor(nwav);
  for (int64_t i = 0; i < nwav; ++i) {
    freq[i] = wavenumber_cm_1[i] * inv_cm_2_Hz;
    // evaluation order of planck_function.cpp:47-48 (left to right)
    prefactor[i] = (d_wavenumber_cm_1[i] * 2.0 * h * inv_cm_2_Hz * pi / (c * c)) *
                   (freq[i] * freq[i] * freq[i]);
  }
#pragma omp parallel for
  for (int it = 0; it < nt; ++it) {
    double* row = ans + (int64_t)it * ans_stride;
    const double t = temperature[it];
    for (int64_t i = 0; i < nwav; ++i) {
      row[i] = prefactor[i] / (std::exp((h / k) * (freq[i] / t)) - 1.0);
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// average_optical_depth_to_g_point: src/ecckd/average_optical_depth.cpp:22-197
//
// od[nlay][nwav] (row stride od_stride), weight[nlay][nwav] (row stride
// w_stride; pass w_stride = 0 for a single row broadcast over layers, which is
// what spread<0>(ssi,nlay) at create_look_up_table.cpp:335 produces),
// g_point[nwav] int32.  Outputs molar_abs/min/max [nlay][ng] with row stride
// out_stride (min/max may be NULL), warn[ng] bit flags (may be NULL).
static inline double trans_fit(double sum_fw, double sum_w, double kd) {
  double x = std::min(0.9999999999999999, sum_fw / sum_w);
  return std::fabs(-std::log(1.0 - x) / kd);
}

ORC_API int orc_average_optical_depth_to_g_point(
    int ng, double reference_surface_vmr, int nlay, int64_t nwav,
    const double* pressure_fl, const double* pressure_hl, const int32_t* g_point,
    const double* od, int64_t od_stride, const double* weight, int64_t w_stride,
    int method, double* molar_abs, double* min_molar_abs, double* max_molar_abs,
    int64_t out_stride, int32_t* warn) {
  const double OD_SCALING = 1.0;  // :35
  if (method < AVG_LINEAR || method > AVG_HYBRID_LOG_TRANS3) {
    return ORC_PARAMETER_ERROR;  // :130-133
  }
#pragma omp parallel for
  for (int ig = 0; ig < ng; ++ig) {
    // index = find(g_point == ig): a full rescan of the spectrum per g (:40)
    std::vector<int64_t> index;
    for (int64_t i = 0; i < nwav; ++i) {
      if (g_point[i] == ig) index.push_back(i);
    }
    const int64_t n = (int64_t)index.size();
    std::vector<double> fit(nlay, 0.0), mn(nlay, 0.0), mx(nlay, 0.0);
    int wflag = 0;

    for (int iz = 0; iz < nlay; ++iz) {
      const double* odr = od + (int64_t)iz * od_stride;
      const double* wr = weight + (int64_t)iz * w_stride;
      int m = method;
      if (method == AVG_HYBRID_LOG_TRANS3) {
        m = (pressure_fl[iz] > 100.0e2) ? AVG_LOGARITHMIC : AVG_TRANSMISSION_3;  // :101
      }
      if (m == AVG_LINEAR) {  // :43-46
        double s = 0.0, sw = 0.0;
        for (int64_t j = 0; j < n; ++j) {
          s += odr[index[j]] * wr[index[j]];
          sw += wr[index[j]];
        }
        fit[iz] = s / sw;
      } else if (m >= AVG_TRANSMISSION && m <= AVG_TRANSMISSION_10) {  // :47-70
        double kk = 1.0;
        if (m == AVG_TRANSMISSION_2) kk = 2.0;
        if (m == AVG_TRANSMISSION_3) kk = 3.0;
        if (m == AVG_TRANSMISSION_10) kk = 10.0;
        // (LW_DIFFUSIVITY*OD_SCALING) for k=1, (LW_DIFFUSIVITY*OD_SCALING*k) otherwise
        const double kd = (m == AVG_TRANSMISSION) ? (LW_DIFFUSIVITY * OD_SCALING)
                                                  : (LW_DIFFUSIVITY * OD_SCALING * kk);
        double s = 0.0, sw = 0.0;
        for (int64_t j = 0; j < n; ++j) {
          s += (1.0 - std::exp(-odr[index[j]] * kd)) * wr[index[j]];
          sw += wr[index[j]];
        }
        fit[iz] = trans_fit(s, sw, kd);
      } else if (m == AVG_SQUARE_ROOT) {  // :71-75
        double s = 0.0, sw = 0.0;
        for (int64_t j = 0; j < n; ++j) {
          s += std::sqrt(odr[index[j]]) * wr[index[j]];
          sw += wr[index[j]];
        }
        double r = s / sw;
        fit[iz] = r * r;
      } else {  // AVG_LOGARITHMIC :76-97 (and :103-117)
        // iindex = find(g_point == ig && od(iz,:) > 0): another rescan (:79)
        int64_t nnz = 0;
        double slog = 0.0, sw_nz = 0.0, sw_all = 0.0;
        for (int64_t i = 0; i < nwav; ++i) {
          if (g_point[i] == ig && odr[i] > 0.0) {
            ++nnz;
            slog += std::log(odr[i]) * wr[i];
            sw_nz += wr[i];
          }
        }
        for (int64_t j = 0; j < n; ++j) sw_all += wr[index[j]];
        if (nnz == n) {
          fit[iz] = std::exp(slog / sw_all);  // :83-84
        } else if (nnz == 0) {
          fit[iz] = 0.0;  // :88
        } else {
          fit[iz] = std::exp(slog / sw_nz) * ((double)nnz / (double)n);  // :92-94
        }
      }
    }

    if (n == 0) {  // :135-141
      wflag |= WARN_EMPTY_G;
      for (int iz = 0; iz < nlay; ++iz) fit[iz] = mn[iz] = mx[iz] = 0.0;
    } else {
      bool any_min_gt = false, any_eq = false;
      for (int iz = 0; iz < nlay; ++iz) {  // :146-151
        const double* odr = od + (int64_t)iz * od_stride;
        double lo = odr[index[0]], hi = odr[index[0]];
        for (int64_t j = 1; j < n; ++j) {
          lo = std::min(lo, odr[index[j]]);
          hi = std::max(hi, odr[index[j]]);
        }
        mn[iz] = lo;
        mx[iz] = hi;
        fit[iz] = std::max(lo, std::min(fit[iz], hi));
      }
      for (int iz = 0; iz < nlay; ++iz) {  // :153-157
        if (mn[iz] > fit[iz]) {
          any_min_gt = true;
          mn[iz] = fit[iz];
        }
      }
      for (int iz = 0; iz < nlay; ++iz) {  // :158-165
        if (mn[iz] > 0.0 && mn[iz] >= mx[iz]) {
          any_eq = true;
          mn[iz] *= 0.99;
          mx[iz] *= 1.01;
        }
      }
      if (any_min_gt) wflag |= WARN_MIN_GT_FIT;
      if (any_eq) wflag |= WARN_MIN_GE_MAX;
    }

    for (int iz = 0; iz < nlay; ++iz) {  // :170-193
      double* o = molar_abs + (int64_t)iz * out_stride + ig;
      if (reference_surface_vmr > 0.0) {
        const double f = (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR) / reference_surface_vmr;
        const double dp = pressure_hl[iz + 1] - pressure_hl[iz];
        *o = f * fit[iz] / dp;
        if (min_molar_abs) {
          min_molar_abs[(int64_t)iz * out_stride + ig] = f * mn[iz] / dp;
          max_molar_abs[(int64_t)iz * out_stride + ig] = f * mx[iz] / dp;
        }
      } else {
        *o = fit[iz];
        if (min_molar_abs) {
          min_molar_abs[(int64_t)iz * out_stride + ig] = mn[iz];
          max_molar_abs[(int64_t)iz * out_stride + ig] = mx[iz];
        }
      }
    }
    if (warn) warn[ig] = wflag;
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// Merge AXPY: src/ecckd/read_merged_spectrum.cpp:149-168
// optical_depth += od * scaling   (scaling_profile == NULL)
// optical_depth += od * spread<1>(scaling_profile, nwav)
ORC_API int orc_merge_axpy(int nlay, int64_t nwav, double* optical_depth, int64_t stride,
                           const double* od, int64_t od_stride, double scaling,
                           const double* scaling_profile) {
  for (int iz = 0; iz < nlay; ++iz) {
    double* dst = optical_depth + (int64_t)iz * stride;
    const double* src = od + (int64_t)iz * od_stride;
    if (scaling_profile) {
      const double s = scaling_profile[iz];
      for (int64_t i = 0; i < nwav; ++i) dst[i] += src[i] * s;
    } else if (scaling != 1.0) {
      for (int64_t i = 0; i < nwav; ++i) dst[i] += src[i] * scaling;
    } else {
      for (int64_t i = 0; i < nwav; ++i) dst[i] += src[i];
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// gpoint_fraction: src/ecckd/create_look_up_table.cpp:512-548
// wavenumber1/2[nbin] are the bin edges; out[ng][nbin].
ORC_API int orc_gpoint_fraction(int ng, int64_t nwav, const int32_t* g_point,
                                const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
                                int nbin, const double* wavenumber1, const double* wavenumber2,
                                double* out) {
#pragma omp parallel for schedule(dynamic)
  for (int ig = 0; ig < ng; ++ig) {
    double wav_per_gpoint = 0.0;
    for (int64_t i = 0; i < nwav; ++i) {
      if (g_point[i] == ig) wav_per_gpoint += d_wavenumber_cm_1[i];
    }
    for (int ib = 0; ib < nbin; ++ib) {
      // one masked scan of the whole spectrum per (g, bin) as in :541-546
      double s = 0.0;
      for (int64_t i = 0; i < nwav; ++i) {
        if (g_point[i] == ig && wavenumber_cm_1[i] > wavenumber1[ib] &&
            wavenumber_cm_1[i] <= wavenumber2[ib]) {
          s += d_wavenumber_cm_1[i];
        }
      }
      out[(int64_t)ig * nbin + ib] = s / wav_per_gpoint;
    }
  }
  return ORC_OK;
}

// Bin edges of the gpoint_fraction grid: create_look_up_table.cpp:518-535.
// Returns nbin; fills wavenumber1/2 if non-NULL (capacity >= nbin).
ORC_API int orc_gpoint_fraction_grid(int do_sw, int nband, const double* band_wn1,
                                     const double* band_wn2, double* wavenumber1,
                                     double* wavenumber2) {
  const int dwav = do_sw ? 50 : 10;
  double lo = band_wn1[0], hi = band_wn2[0];
  for (int i = 1; i < nband; ++i) {
    lo = std::min(lo, band_wn1[i]);
    hi = std::max(hi, band_wn2[i]);
  }
  const int startwav = (int)(std::floor(lo / dwav) * dwav);
  const int endwav = (int)(std::ceil(hi / dwav) * dwav);
  const int i0 = startwav / dwav, i1 = endwav / dwav;
  const int nbin = i1 - i0;
  if (wavenumber1 && wavenumber2) {
    for (int i = 0; i < nbin; ++i) {
      wavenumber1[i] = (double)dwav * (i0 + i);
      wavenumber2[i] = (double)dwav * (i0 + i + 1);
    }
  }
  return nbin;
}

// Solar irradiance per bin: create_look_up_table.cpp:555-561
ORC_API int orc_ssi_intervals(int64_t nwav, const double* wavenumber_cm_1, const double* ssi,
                              int nbin, const double* wavenumber1, const double* wavenumber2,
                              double* out) {
  for (int ib = 0; ib < nbin; ++ib) {
    double s = 0.0;
    for (int64_t i = 0; i < nwav; ++i) {
      if (wavenumber_cm_1[i] > wavenumber1[ib] && wavenumber_cm_1[i] <= wavenumber2[ib]) s += ssi[i];
    }
    out[ib] = s;
  }
  return ORC_OK;
}

// Planck look-up table per g point: create_look_up_table.cpp:579-591
// temperature_lut = 120..350 K (231 values); out[nlut][ng]
ORC_API int orc_planck_lut(int ng, int64_t nwav, const int32_t* g_point,
                           const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
                           int nlut, const double* temperature_lut, double* out) {
  for (int ig = 0; ig < ng; ++ig) {
    std::vector<double> wn, dwn;
    for (int64_t i = 0; i < nwav; ++i) {
      if (g_point[i] == ig) {
        wn.push_back(wavenumber_cm_1[i]);
        dwn.push_back(d_wavenumber_cm_1[i]);
      }
    }
    const int64_t n = (int64_t)wn.size();
    std::vector<double> tmp((size_t)nlut * (size_t)std::max<int64_t>(n, 1));
    if (n > 0) orc_planck_function(nlut, temperature_lut, n, wn.data(), dwn.data(), tmp.data(), n);
    for (int it = 0; it < nlut; ++it) {
      double s = 0.0;
      for (int64_t j = 0; j < n; ++j) s += tmp[(size_t)it * n + j];
      out[(int64_t)it * ng + ig] = s;
    }
  }
  return ORC_OK;
}

// Remove g points that occupy none of the spectrum:
// create_look_up_table.cpp:111-160.  g_point is rewritten in place;
// g_point_map[new_ng] receives the old index of each kept g point.
// Returns the new ng, or -1 if a spectral point is left unassigned (:146-149).
ORC_API int orc_compact_g_points(int ng, int64_t nwav, int32_t* g_point, int32_t* g_point_map) {
  std::vector<char> used(ng, 0);
  for (int64_t i = 0; i < nwav; ++i) {
    if (g_point[i] >= 0 && g_point[i] < ng) used[g_point[i]] = 1;
  }
  std::vector<int32_t> new_of_old(ng, -1);
  int new_ng = 0;
  for (int ig = 0; ig < ng; ++ig) {
    if (used[ig]) {
      g_point_map[new_ng] = ig;
      new_of_old[ig] = new_ng++;
    }
  }
  int bad = 0;
  for (int64_t i = 0; i < nwav; ++i) {
    int32_t g = g_point[i];
    int32_t v = (g >= 0 && g < ng) ? new_of_old[g] : -1;
    if (v < 0) bad = 1;
    g_point[i] = v;
  }
  return bad ? -1 : new_ng;
}

// ---------------------------------------------------------------------------
// radiative_transfer_lw<false>: src/ecckd/radiative_transfer_lw.cpp:25-60
// planck[nlay+1][n], od[nlay][n], flux_dn/up[nlay+1][n]; all dense rows.
ORC_API int orc_radiative_transfer_lw(int nlay, int64_t n, const double* planck, const double* od,
                                      const double* surf_emissivity, const double* surf_planck,
                                      double* flux_dn, double* flux_up) {
  std::vector<double> emis((size_t)nlay * n), factor((size_t)nlay * n);
  for (int iz = 0; iz < nlay; ++iz) {
    for (int64_t i = 0; i < n; ++i) {
      const double tau = od[(int64_t)iz * n + i];
      const double e = 1.0 - std::exp(-LW_DIFFUSIVITY * tau);
      emis[(size_t)iz * n + i] = e;
      factor[(size_t)iz * n + i] = (e > 1.0e-5) ? (1.0 - e * (1.0 / LW_DIFFUSIVITY) / tau) : (0.5 * e);
    }
  }
  for (int64_t i = 0; i < n; ++i) flux_dn[i] = 0.0;
  for (int iz = 0; iz < nlay; ++iz) {
    for (int64_t i = 0; i < n; ++i) {
      const double e = emis[(size_t)iz * n + i], f = factor[(size_t)iz * n + i];
      flux_dn[(int64_t)(iz + 1) * n + i] = flux_dn[(int64_t)iz * n + i] * (1.0 - e) +
                                           planck[(int64_t)iz * n + i] * (e - f) +
                                           planck[(int64_t)(iz + 1) * n + i] * f;
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    flux_up[(int64_t)nlay * n + i] =
        surf_planck[i] * surf_emissivity[i] + (1.0 - surf_emissivity[i]) * flux_dn[(int64_t)nlay * n + i];
  }
  for (int iz = nlay - 1; iz >= 0; --iz) {
    for (int64_t i = 0; i < n; ++i) {
      const double e = emis[(size_t)iz * n + i], f = factor[(size_t)iz * n + i];
      flux_up[(int64_t)iz * n + i] = flux_up[(int64_t)(iz + 1) * n + i] * (1.0 - e) +
                                     planck[(int64_t)(iz + 1) * n + i] * (e - f) +
                                     planck[(int64_t)iz * n + i] * f;
    }
  }
  return ORC_OK;
}

// radiative_transfer_lw_bb<false>: src/ecckd/radiative_transfer_lw.cpp:85-142
// broadband (sum over n) fluxes flux_dn/up[nlay+1]
ORC_API int orc_radiative_transfer_lw_bb(int nlay, int64_t n, const double* planck,
                                         const double* spectral_od, const double* grey_od,
                                         const double* surf_emissivity, const double* surf_planck,
                                         double* flux_dn, double* flux_up) {
  const double THRESHOLD_EMISSIVITY = 1.0e-5;
  std::vector<double> flux(n, 0.0);
  flux_dn[0] = 0.0;
  for (int iz = 0; iz < nlay; ++iz) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const double tau = spectral_od[(int64_t)iz * n + i] + grey_od[iz];
      const double e = 1.0 - std::exp(-LW_DIFFUSIVITY * tau);
      const double f = std::max(1.0 - (1.0 / LW_DIFFUSIVITY) * std::max(e, THRESHOLD_EMISSIVITY) /
                                          std::max(tau, THRESHOLD_EMISSIVITY / LW_DIFFUSIVITY),
                                0.5 * THRESHOLD_EMISSIVITY);
      flux[i] = flux[i] * (1.0 - e) + planck[(int64_t)iz * n + i] * (e - f) +
                planck[(int64_t)(iz + 1) * n + i] * f;
      s += flux[i];
    }
    flux_dn[iz + 1] = s;
  }
  {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      flux[i] = surf_planck[i] * surf_emissivity[i] + (1.0 - surf_emissivity[i]) * flux[i];
      s += flux[i];
    }
    flux_up[nlay] = s;
  }
  for (int iz = nlay - 1; iz >= 0; --iz) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const double tau = spectral_od[(int64_t)iz * n + i] + grey_od[iz];
      const double e = 1.0 - std::exp(-LW_DIFFUSIVITY * tau);
      const double f = std::max(1.0 - (1.0 / LW_DIFFUSIVITY) * std::max(e, THRESHOLD_EMISSIVITY) /
                                          std::max(tau, THRESHOLD_EMISSIVITY / LW_DIFFUSIVITY),
                                0.5 * THRESHOLD_EMISSIVITY);
      flux[i] = flux[i] * (1.0 - e) + planck[(int64_t)(iz + 1) * n + i] * (e - f) +
                planck[(int64_t)iz * n + i] * f;
      s += flux[i];
    }
    flux_up[iz] = s;
  }
  return ORC_OK;
}

// radiative_transfer_direct_sw<false>: src/ecckd/radiative_transfer_sw.cpp:24-43
ORC_API int orc_radiative_transfer_direct_sw(int nlay, int64_t n, double cos_sza, const double* ssi,
                                             const double* od, double* flux_dn) {
  const double minus_sec_sza = -1.0 / cos_sza;
  for (int64_t i = 0; i < n; ++i) flux_dn[i] = cos_sza * ssi[i];
  for (int iz = 0; iz < nlay; ++iz) {
    for (int64_t i = 0; i < n; ++i) {
      flux_dn[(int64_t)(iz + 1) * n + i] =
          flux_dn[(int64_t)iz * n + i] * std::exp(minus_sec_sza * od[(int64_t)iz * n + i]);
    }
  }
  return ORC_OK;
}

// radiative_transfer_norayleigh_sw<false>: src/ecckd/radiative_transfer_sw.cpp:47-77
ORC_API int orc_radiative_transfer_norayleigh_sw(int nlay, int64_t n, double cos_sza,
                                                 const double* ssi, const double* od,
                                                 const double* albedo, double* flux_dn,
                                                 double* flux_up) {
  const double minus_sec_tsza = -2.0;
  orc_radiative_transfer_direct_sw(nlay, n, cos_sza, ssi, od, flux_dn);
  for (int64_t i = 0; i < n; ++i) flux_up[(int64_t)nlay * n + i] = flux_dn[(int64_t)nlay * n + i] * albedo[i];
  for (int iz = nlay - 1; iz >= 0; --iz) {
    for (int64_t i = 0; i < n; ++i) {
      flux_up[(int64_t)iz * n + i] =
          flux_up[(int64_t)(iz + 1) * n + i] * std::exp(minus_sec_tsza * od[(int64_t)iz * n + i]);
    }
  }
  return ORC_OK;
}

// radiative_transfer_direct_sw_bb / norayleigh_sw_bb:
// src/ecckd/radiative_transfer_sw.cpp:116-141, 145-184.  flux_up may be NULL
// (direct only); albedo is a scalar in the bb variant.
ORC_API int orc_radiative_transfer_sw_bb(int nlay, int64_t n, double cos_sza, const double* ssi,
                                         const double* spectral_od, const double* grey_od,
                                         double albedo, double* flux_dn, double* flux_up) {
  const double minus_sec_sza = -1.0 / cos_sza;
  const double minus_sec_tsza = -2.0;
  std::vector<double> flux(n);
  double s0 = 0.0;
  for (int64_t i = 0; i < n; ++i) s0 += ssi[i];
  flux_dn[0] = cos_sza * s0;
  for (int64_t i = 0; i < n; ++i) flux[i] = cos_sza * ssi[i];
  for (int iz = 0; iz < nlay; ++iz) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      flux[i] = flux[i] * std::exp(minus_sec_sza * (spectral_od[(int64_t)iz * n + i] + grey_od[iz]));
      s += flux[i];
    }
    flux_dn[iz + 1] = s;
  }
  if (flux_up) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      flux[i] *= albedo;
      s += flux[i];
    }
    flux_up[nlay] = s;
    for (int iz = nlay - 1; iz >= 0; --iz) {
      double t = 0.0;
      for (int64_t i = 0; i < n; ++i) {
        flux[i] = flux[i] * std::exp(minus_sec_tsza * (spectral_od[(int64_t)iz * n + i] + grey_od[iz]));
        t += flux[i];
      }
      flux_up[iz] = t;
    }
  }
  return ORC_OK;
}

// heating_rate<false>: src/ecckd/heating_rate.h:25-49.  flux_up may be NULL.
ORC_API int orc_heating_rate(int nlay, int64_t n, const double* pressure_hl, const double* flux_dn,
                             const double* flux_up, double* hr) {
  for (int iz = 0; iz < nlay; ++iz) {
    const double conv = -(ACCEL_GRAVITY / SPECIFIC_HEAT_AIR) / (pressure_hl[iz + 1] - pressure_hl[iz]);
    for (int64_t i = 0; i < n; ++i) {
      if (flux_up) {
        hr[(int64_t)iz * n + i] = conv * (flux_dn[(int64_t)(iz + 1) * n + i] - flux_dn[(int64_t)iz * n + i] -
                                          flux_up[(int64_t)(iz + 1) * n + i] + flux_up[(int64_t)iz * n + i]);
      } else {
        hr[(int64_t)iz * n + i] = conv * (flux_dn[(int64_t)(iz + 1) * n + i] - flux_dn[(int64_t)iz * n + i]);
      }
    }
  }
  return ORC_OK;
}

// heating_rate_single<false>: src/ecckd/heating_rate.h:52-72
ORC_API int orc_heating_rate_single(int nlay, const double* pressure_hl, const double* flux_dn,
                                    const double* flux_up, double* hr) {
  for (int iz = 0; iz < nlay; ++iz) {
    const double c = -((ACCEL_GRAVITY / SPECIFIC_HEAT_AIR) / (pressure_hl[iz + 1] - pressure_hl[iz]));
    if (flux_up) {
      hr[iz] = c * (flux_dn[iz + 1] - flux_dn[iz] - flux_up[iz + 1] + flux_up[iz]);
    } else {
      hr[iz] = c * (flux_dn[iz + 1] - flux_dn[iz]);
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// CkdModel::calc_optical_depth: src/ecckd/ckd_model.cpp:923-1102 (linear
// interpolation branches only; logarithmic_interpolation is hard-wired false,
// ckd_model.h:359).
//
// conc_dependence codes: ckd_model.cpp:409-413 (0 none, 1 linear, 2 lut,
// 3 relative-linear).  molar_abs is [nt][np][ng] or, for LUT, [nconc][nt][np][ng].
// temperature_lut[nt][np]; log_pressure_lut[np]; vmr_fl may be NULL.
// od[ncol][nlay][ng] is OVERWRITTEN (the reference returns a fresh array).
ORC_API int orc_ckd_calc_optical_depth(int conc_dependence, int nconc, const double* vmr_lut,
                                       double reference_vmr, int nt, int np_lut, int ng,
                                       const double* log_pressure_lut, const double* temperature_lut,
                                       const double* molar_abs, int ncol, int nlay,
                                       const double* pressure_hl, const double* temperature_fl,
                                       const double* vmr_fl, double* od) {
  const double log_p_0 = log_pressure_lut[0];
  const double d_log_p = log_pressure_lut[1] - log_pressure_lut[0];
  const double d_t = temperature_lut[(size_t)1 * np_lut + 0] - temperature_lut[0];
  const double global_weight = 1.0 / (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR);
  if ((conc_dependence == 1 || conc_dependence == 2 || conc_dependence == 3) && !vmr_fl) return 1;
  for (int icol = 0; icol < ncol; ++icol) {
    for (int ip = 0; ip < nlay; ++ip) {
      const double p1 = pressure_hl[(int64_t)icol * (nlay + 1) + ip + 1];
      const double p0 = pressure_hl[(int64_t)icol * (nlay + 1) + ip];
      const double log_pressure_fl = std::log(0.5 * (p1 + p0));
      double pindex0 = (log_pressure_fl - log_p_0) / d_log_p;
      pindex0 = std::fmax(0.0, std::fmin(pindex0, np_lut - 1.0001));
      const int ip0 = (int)pindex0;
      const double pweight1 = pindex0 - ip0;
      const double pweight0 = 1.0 - pweight1;
      const double t_0 = pweight0 * temperature_lut[ip0] + pweight1 * temperature_lut[ip0 + 1];
      double tindex0 = (temperature_fl[(int64_t)icol * nlay + ip] - t_0) / d_t;
      tindex0 = std::fmax(0.0, std::fmin(tindex0, nt - 1.0001));
      const int it0 = (int)tindex0;
      const double tweight1 = tindex0 - it0;
      const double tweight0 = 1.0 - tweight1;
      const double simple_weight = global_weight * (p1 - p0);
      double weight = 0.0;
      if (vmr_fl) {
        const double v = vmr_fl[(int64_t)icol * nlay + ip];
        weight = (conc_dependence == 3) ? simple_weight * (v - reference_vmr) : simple_weight * v;
      }
      double* o = od + ((int64_t)icol * nlay + ip) * ng;
      const size_t sT = (size_t)np_lut * ng, sP = (size_t)ng;
      if (conc_dependence == 2) {
        const double log_conc = std::log(vmr_fl[(int64_t)icol * nlay + ip]);
        const double d_log_c = std::log(vmr_lut[1] / vmr_lut[0]);
        double cindex0 = (log_conc - std::log(vmr_lut[0])) / d_log_c;
        cindex0 = std::fmax(0.0, std::fmin(cindex0, nconc - 1.0001));
        const int ic0 = (int)cindex0;
        const double cweight1 = cindex0 - ic0;
        const double cweight0 = 1.0 - cweight1;
        const size_t sC = (size_t)nt * sT;
        const double* k00 = molar_abs + ic0 * sC + it0 * sT + ip0 * sP;
        const double* k01 = k00 + sP;
        const double* k10 = k00 + sT;
        const double* k11 = k10 + sP;
        const double* l00 = k00 + sC;
        const double* l01 = l00 + sP;
        const double* l10 = l00 + sT;
        const double* l11 = l10 + sP;
        for (int g = 0; g < ng; ++g) {
          o[g] = weight * (cweight0 * (tweight0 * (pweight0 * k00[g] + pweight1 * k01[g]) +
                                       tweight1 * (pweight0 * k10[g] + pweight1 * k11[g])) +
                           cweight1 * (tweight0 * (pweight0 * l00[g] + pweight1 * l01[g]) +
                                       tweight1 * (pweight0 * l10[g] + pweight1 * l11[g])));
        }
      } else {
        const double w = (conc_dependence == 0) ? simple_weight : weight;
        const double* k00 = molar_abs + it0 * sT + ip0 * sP;
        const double* k01 = k00 + sP;
        const double* k10 = k00 + sT;
        const double* k11 = k10 + sP;
        for (int g = 0; g < ng; ++g) {
          o[g] = w * (tweight0 * (pweight0 * k00[g] + pweight1 * k01[g]) +
                      tweight1 * (pweight0 * k10[g] + pweight1 * k11[g]));
        }
      }
    }
  }
  return ORC_OK;
}

// CkdModel::calc_planck_function: src/ecckd/ckd_model.cpp:1105-1145
// planck_lut[nlut][ng], temperature_planck[nlut] (uniform spacing);
// temperature[n] -> planck[n][ng]
ORC_API int orc_ckd_calc_planck_function(int nlut, int ng, const double* temperature_planck,
                                         const double* planck_lut, int64_t n, const double* temperature,
                                         double* planck) {
  const double d_t = temperature_planck[1] - temperature_planck[0];
  const double t0 = temperature_planck[0];
  for (int64_t it = 0; it < n; ++it) {
    const double tindex0 = (temperature[it] - t0) / d_t;
    double* o = planck + it * ng;
    if (tindex0 >= 0) {
      const int it0 = std::min((int)tindex0, nlut - 2);
      const double tweight1 = tindex0 - it0;
      const double tweight0 = 1.0 - tweight1;
      for (int g = 0; g < ng; ++g) {
        o[g] = tweight0 * planck_lut[(size_t)it0 * ng + g] + tweight1 * planck_lut[(size_t)(it0 + 1) * ng + g];
      }
    } else {
      for (int g = 0; g < ng; ++g) o[g] = (temperature[it] / t0) * planck_lut[g];
    }
  }
  return ORC_OK;
}

// CkdModel::calc_rayleigh_optical_depth: src/ecckd/ckd_model.h:242-252
ORC_API int orc_ckd_calc_rayleigh_optical_depth(int ncol, int nlay, int ng, const double* pressure_hl,
                                                const double* rayleigh_molar_scat, double* od) {
  for (int icol = 0; icol < ncol; ++icol) {
    for (int ip = 0; ip < nlay; ++ip) {
      const double dp = pressure_hl[(int64_t)icol * (nlay + 1) + ip + 1] - pressure_hl[(int64_t)icol * (nlay + 1) + ip];
      const double w = dp * (1.0 / (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR));
      for (int g = 0; g < ng; ++g) od[((int64_t)icol * nlay + ip) * ng + g] = w * rayleigh_molar_scat[g];
    }
  }
  return ORC_OK;
}

// Full-level temperature from half levels: src/ecckd/run_ckd.cpp:119-122
ORC_API int orc_temperature_fl(int ncol, int nlay, const double* temperature_hl, const double* pressure_hl,
                               double* temperature_fl) {
  for (int icol = 0; icol < ncol; ++icol) {
    const double* t = temperature_hl + (int64_t)icol * (nlay + 1);
    const double* p = pressure_hl + (int64_t)icol * (nlay + 1);
    for (int i = 0; i < nlay; ++i) {
      temperature_fl[(int64_t)icol * nlay + i] = (t[i] * p[i] + t[i + 1] * p[i + 1]) / (p[i] + p[i + 1]);
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// run_ckd numerical core: src/ecckd/run_ckd.cpp:119-132, 261-367.
// Gas tables are passed as parallel arrays of length ngas.  vmr_fl[igas] may be
// NULL ("variable not in the input file", :277-281).  Outputs (any may be NULL
// except od): od[ncol][nlay][ng] total clamped optical depth, planck_hl
// [ncol][nlay+1][ng], planck_surf[ncol][ng], flux_dn/up[ncol][nlay+1][ng]
// (LW) or flux_dn = direct SW flux; bb_flux_dn/up[ncol][nlay+1] = sum over g.
struct orc_gas_table {
  int32_t conc_dependence;
  int32_t nconc;
  const double* vmr_lut;
  double reference_vmr;
  const double* molar_abs;
};

ORC_API int orc_run_ckd(int is_sw, int ngas, const orc_gas_table* gases, const double* const* vmr_fl,
                        int nt, int np_lut, int ng, const double* log_pressure_lut,
                        const double* temperature_lut, int nlut, const double* temperature_planck,
                        const double* planck_lut, const double* solar_irradiance,
                        const double* rayleigh_molar_scat, double tsi, int ncol, int nlay,
                        const double* pressure_hl, const double* temperature_hl, double* od,
                        double* planck_hl, double* planck_surf, double* flux_dn, double* flux_up,
                        double* bb_flux_dn, double* bb_flux_up) {
  const int64_t nod = (int64_t)ncol * nlay * ng;
  std::vector<double> t_fl((size_t)ncol * nlay), od_gas((size_t)nod);
  orc_temperature_fl(ncol, nlay, temperature_hl, pressure_hl, t_fl.data());
  for (int64_t i = 0; i < nod; ++i) od[i] = 0.0;
  for (int igas = 0; igas < ngas; ++igas) {
    int rc = orc_ckd_calc_optical_depth(gases[igas].conc_dependence, gases[igas].nconc, gases[igas].vmr_lut,
                                        gases[igas].reference_vmr, nt, np_lut, ng, log_pressure_lut,
                                        temperature_lut, gases[igas].molar_abs, ncol, nlay, pressure_hl,
                                        t_fl.data(), vmr_fl ? vmr_fl[igas] : nullptr, od_gas.data());
    if (rc) return rc;
    for (int64_t i = 0; i < nod; ++i) od[i] += od_gas[i];  // :312
  }
  for (int64_t i = 0; i < nod; ++i) od[i] = std::max(od[i], 0.0);  // :317
  const int64_t nhl = (int64_t)(nlay + 1) * ng;
  if (!is_sw) {
    std::vector<double> plh_local, pls_local;
    if (!planck_hl) {
      plh_local.resize((size_t)ncol * nhl);
      planck_hl = plh_local.data();
    }
    if (!planck_surf) {
      pls_local.resize((size_t)ncol * ng);
      planck_surf = pls_local.data();
    }
    orc_ckd_calc_planck_function(nlut, ng, temperature_planck, planck_lut, (int64_t)ncol * (nlay + 1),
                                 temperature_hl, planck_hl);  // :131
    std::vector<double> t_surf(ncol);
    for (int icol = 0; icol < ncol; ++icol) t_surf[icol] = temperature_hl[(int64_t)icol * (nlay + 1) + nlay];
    orc_ckd_calc_planck_function(nlut, ng, temperature_planck, planck_lut, ncol, t_surf.data(), planck_surf);  // :132
    if (flux_dn && flux_up) {
      std::vector<double> emis(ng, 1.0);  // :344-345
      for (int icol = 0; icol < ncol; ++icol) {  // serial over columns, :347
        orc_radiative_transfer_lw(nlay, ng, planck_hl + icol * nhl, od + (int64_t)icol * nlay * ng, emis.data(),
                                  planck_surf + (int64_t)icol * ng, flux_dn + icol * nhl, flux_up + icol * nhl);
      }
    }
  } else {
    double ssum = 0.0;
    for (int g = 0; g < ng; ++g) ssum += solar_irradiance[g];
    const double tsi_scaling = tsi / ssum;  // :98
    std::vector<double> ray((size_t)nod), tot((size_t)nlay * ng), ssi(ng);
    orc_ckd_calc_rayleigh_optical_depth(ncol, nlay, ng, pressure_hl, rayleigh_molar_scat, ray.data());
    for (int g = 0; g < ng; ++g) ssi[g] = tsi_scaling * solar_irradiance[g];
    if (flux_dn) {
      for (int icol = 0; icol < ncol; ++icol) {  // :360-365
        for (int64_t i = 0; i < (int64_t)nlay * ng; ++i) tot[i] = od[(int64_t)icol * nlay * ng + i] + ray[(int64_t)icol * nlay * ng + i];
        orc_radiative_transfer_direct_sw(nlay, ng, 0.5, ssi.data(), tot.data(), flux_dn + icol * nhl);
      }
    }
  }
  // sum(flux,2): :355-356, :367
  if (flux_dn && bb_flux_dn) {
    for (int64_t r = 0; r < (int64_t)ncol * (nlay + 1); ++r) {
      double s = 0.0;
      for (int g = 0; g < ng; ++g) s += flux_dn[r * ng + g];
      bb_flux_dn[r] = s;
    }
  }
  if (!is_sw && flux_up && bb_flux_up) {
    for (int64_t r = 0; r < (int64_t)ncol * (nlay + 1); ++r) {
      double s = 0.0;
      for (int g = 0; g < ng; ++g) s += flux_up[r * ng + g];
      bb_flux_up[r] = s;
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// Band membership + per-band stable sort -> rank:
// src/ecckd/reorder_spectrum.cpp:30-34, 262-300.
// key[nwav] is the sorting variable; outputs iband[nwav], rank[nwav],
// ordered_index[nwav] (any may be NULL).  Returns PARAMETER_ERROR when a band
// selects no wavenumber (the reference would index an empty vector at :291).
ORC_API int orc_rank_bands(int64_t nwav, const double* wavenumber_cm_1, const double* key, int nband,
                           const double* band_bound1, const double* band_bound2, int32_t* iband,
                           int32_t* rank, int32_t* ordered_index) {
  std::vector<int32_t> g_index(nwav);
  for (int64_t i = 0; i < nwav; ++i) g_index[i] = (int32_t)i;
  if (iband) {
    for (int64_t i = 0; i < nwav; ++i) iband[i] = -1;
  }
  for (int jband = 0; jband < nband; ++jband) {
    int64_t index1 = -1, index2 = -1;
    for (int64_t i = 0; i < nwav; ++i) {
      const double nu = wavenumber_cm_1[i];
      const bool in = (jband < nband - 1) ? (nu >= band_bound1[jband] && nu < band_bound2[jband])
                                          : (nu >= band_bound1[jband] && nu <= band_bound2[jband]);
      if (in) {
        if (index1 < 0) index1 = i;
        index2 = i;
        if (iband) iband[i] = jband;
      }
    }
    if (index1 < 0) return ORC_PARAMETER_ERROR;
    // MyCompare: data[i1] < data[i2]  (:30-34)
    std::stable_sort(g_index.begin() + index1, g_index.begin() + index2 + 1,
                     [key](int32_t a, int32_t b) { return key[a] < key[b]; });
  }
  if (ordered_index) std::memcpy(ordered_index, g_index.data(), sizeof(int32_t) * nwav);
  if (rank) {
    for (int64_t i = 0; i < nwav; ++i) rank[g_index[i]] = (int32_t)i;  // :299-300
  }
  return ORC_OK;
}

// Idealised temperature profile of reorder_spectrum.cpp:119-124:
// linear in log(p) between (log 1 Pa, 173.15 K) and (log 1e5 Pa, 288.15 K).
// Inputs are kept inside [1, 1e5] Pa; outside, values are
// clamped to the end temperatures.
static double ideal_temperature(double p) {
  const double x0 = std::log(1.0), x1 = std::log(100000.0);
  const double y0 = 273.15 - 100.0, y1 = 273.15 + 15.0;
  double x = std::log(p);
  if (x <= x0) return y0;
  if (x >= x1) return y1;
  return y0 + (y1 - y0) * (x - x0) / (x1 - x0);
}

// LW sorting key (pseudo-height of peak cooling): reorder_spectrum.cpp:111-190.
// od[nlay][nwav] -> key[nwav], column_od[nwav].
ORC_API int orc_sort_key_lw(int nlay, int64_t nwav, const double* pressure_hl, const double* wavenumber_cm_1,
                            const double* d_wavenumber_cm_1, const double* od, double threshold_optical_depth,
                            double* key, double* column_od) {
  std::vector<double> t_hl(nlay + 1);
  for (int i = 0; i <= nlay; ++i) t_hl[i] = ideal_temperature(pressure_hl[i]);
  std::vector<double> planck((size_t)(nlay + 1) * nwav), surf_planck(nwav), emis(nwav, 1.0);
  orc_planck_function(nlay + 1, t_hl.data(), nwav, wavenumber_cm_1, d_wavenumber_cm_1, planck.data(), nwav);
  orc_planck_function(1, &t_hl[nlay], nwav, wavenumber_cm_1, d_wavenumber_cm_1, surf_planck.data(), nwav);
  std::vector<double> fdn((size_t)(nlay + 1) * nwav), fup((size_t)(nlay + 1) * nwav), hr((size_t)nlay * nwav);
  orc_radiative_transfer_lw(nlay, nwav, planck.data(), od, emis.data(), surf_planck.data(), fdn.data(), fup.data());
  orc_heating_rate(nlay, nwav, pressure_hl, fdn.data(), fup.data(), hr.data());
  std::vector<double> dh(nlay), dhh(nlay);
  for (int i = 0; i < nlay; ++i) {
    const double ph = std::log(pressure_hl[nlay]) - 0.5 * (std::log(pressure_hl[i]) + std::log(pressure_hl[i + 1]));
    dh[i] = std::log(pressure_hl[i + 1]) - std::log(pressure_hl[i]);
    dhh[i] = dh[i] * ph;
  }
  for (int64_t w = 0; w < nwav; ++w) {
    double col = 0.0, num = 0.0, den = 0.0;
    for (int i = 0; i < nlay; ++i) {
      col += od[(int64_t)i * nwav + w];
      double h = hr[(size_t)i * nwav + w];
      if (h > 0.0) h = 0.0;  // :175
      num += h * dhh[i];
      den += h * dh[i];
    }
    double k = num / den;
    if (threshold_optical_depth > 0.0 && col < threshold_optical_depth) {
      k = -threshold_optical_depth + col;  // :187-190
    }
    key[w] = k;
    if (column_od) column_od[w] = col;
  }
  return ORC_OK;
}

// SW sorting key (height where optical depth from TOA reaches threshold):
// reorder_spectrum.cpp:197-228.  Returns PROCESSING_ERROR where the reference
// throws (key > 30, :214-216).
ORC_API int orc_sort_key_sw(int nlay, int64_t nwav, const double* pressure_hl, const double* od,
                            double threshold_optical_depth, double* key, double* column_od) {
  std::vector<double> h_hl(nlay + 1);
  for (int i = 0; i <= nlay; ++i) h_hl[i] = std::log(pressure_hl[nlay]) - std::log(pressure_hl[i]);
  int rc = ORC_OK;
  for (int64_t w = 0; w < nwav; ++w) {
    double col = 0.0;
    for (int i = 0; i < nlay; ++i) col += od[(int64_t)i * nwav + w];
    if (column_od) column_od[w] = col;
    if (col <= threshold_optical_depth) {
      key[w] = col - threshold_optical_depth;
    } else {
      double cum = 0.0;
      key[w] = 0.0;
      for (int i = 0; i < nlay; ++i) {
        const double tau = od[(int64_t)i * nwav + w];
        const double next = cum + tau;
        if (next >= threshold_optical_depth) {
          key[w] = ((threshold_optical_depth - cum) * h_hl[i + 1] + (next - threshold_optical_depth) * h_hl[i]) /
                   std::max(1.0e-12, tau);
          if (key[w] > 30.0) rc = ORC_PROCESSING_ERROR;
          break;
        }
        cum = next;
      }
    }
  }
  return rc;
}

// ---------------------------------------------------------------------------
// SingleGasData::store_g_points: src/ecckd/single_gas_data.h:56-62
// g[w] = ig for the LAST ig with rank1[ig] <= rank[w] <= rank2[ig], else -1.
ORC_API int orc_store_g_points(int64_t nwav, const int32_t* rank, int ng, const int32_t* rank1,
                               const int32_t* rank2, int32_t* g_point) {
  for (int64_t i = 0; i < nwav; ++i) g_point[i] = -1;
  for (int ig = 0; ig < ng; ++ig) {
    for (int64_t i = 0; i < nwav; ++i) {
      if (rank[i] >= rank1[ig] && rank[i] <= rank2[ig]) g_point[i] = ig;
    }
  }
  return ORC_OK;
}

// overlap_g_points: src/ecckd/single_gas_data.cpp:23-124 (hypercube partition
// of Hogan 2010).  n_g_points[ngas][nband]; sorting_variable is the
// concatenation over gases of each gas's per-g array, with gas_offset[ngas]
// giving the start of each gas's slice.  g_min/g_max[ngas][ng_total_cap],
// band_number[ng_total_cap].  Returns the merged ng, or -1 on failure (:92-95),
// or -2 if cap is too small.
ORC_API int orc_overlap_g_points(int ngas, int nband, const int32_t* n_g_points, const double* sorting_variable,
                                 const int32_t* gas_offset, int cap, int32_t* g_min, int32_t* g_max,
                                 int32_t* band_number) {
  std::vector<int> ng_band(nband);
  int ng = 0;
  for (int ib = 0; ib < nband; ++ib) {
    ng_band[ib] = 1 - ngas;
    for (int igas = 0; igas < ngas; ++igas) ng_band[ib] += n_g_points[igas * nband + ib];
    ng += ng_band[ib];
  }
  if (ng > cap) return -2;
  {
    int ig = 0;
    for (int ib = 0; ib < nband; ++ib) {
      for (int j = 0; j < ng_band[ib]; ++j) band_number[ig++] = ib;
    }
  }
  int ig = 0;
  std::vector<int> ig_gas(ngas, 0), ig_gas_start(ngas);
  for (int ib = 0; ib < nband; ++ib) {
    ig_gas_start = ig_gas;
    for (int igas = 0; igas < ngas; ++igas) {
      g_min[igas * cap + ig] = ig_gas_start[igas];
      g_max[igas * cap + ig] = ig_gas_start[igas];
    }
    for (int igb = 1; igb < ng_band[ib]; ++igb) {
      double min_sv = 1.0e30;
      int found = -1;
      for (int igas = 0; igas < ngas; ++igas) {
        double my = 1.0e30;
        if (ig_gas[igas] < ig_gas_start[igas] + n_g_points[igas * nband + ib] - 1) {
          my = sorting_variable[gas_offset[igas] + ig_gas[igas] + 1];
        }
        if (my < min_sv) {
          min_sv = my;
          found = igas;
        }
      }
      if (found < 0) return -1;
      ++ig_gas[found];
      ++ig;
      for (int igas = 0; igas < ngas; ++igas) {
        if (igas == found) {
          g_min[igas * cap + ig] = ig_gas[igas];
          g_max[igas * cap + ig] = ig_gas[igas];
        } else {
          g_min[igas * cap + ig] = ig_gas_start[igas];
          g_max[igas * cap + ig] = ig_gas[igas];
        }
      }
    }
    ++ig;
    for (int igas = 0; igas < ngas; ++igas) ++ig_gas[igas];
  }
  return ng;
}

// Final multi-gas g-point assignment: src/ecckd/find_g_points.cpp:1459-1475.
// gas_g_point[ngas][nwav]; g_min/g_max[ngas][cap].  Ascending ig, later ig
// overwrites.  n_unassigned (may be NULL) counts g_point == -1 (:1478).
ORC_API int orc_assign_g_points(int ngas, int64_t nwav, const int32_t* gas_g_point, int ng, int cap,
                                const int32_t* g_min, const int32_t* g_max, int32_t* g_point,
                                int64_t* n_unassigned) {
  for (int64_t i = 0; i < nwav; ++i) g_point[i] = -1;
  std::vector<char> is_found(nwav);
  for (int ig = 0; ig < ng; ++ig) {
    std::fill(is_found.begin(), is_found.end(), 1);
    for (int igas = 0; igas < ngas; ++igas) {
      const int32_t lo = g_min[igas * cap + ig], hi = g_max[igas * cap + ig];
      const int32_t* gp = gas_g_point + (int64_t)igas * nwav;
      for (int64_t i = 0; i < nwav; ++i) {
        if (gp[i] < lo || gp[i] > hi) is_found[i] = 0;
      }
    }
    for (int64_t i = 0; i < nwav; ++i) {
      if (is_found[i]) g_point[i] = ig;
    }
  }
  if (n_unassigned) {
    int64_t c = 0;
    for (int64_t i = 0; i < nwav; ++i) c += (g_point[i] == -1);
    *n_unassigned = c;
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------
// Rayleigh molar scattering: src/ecckd/rayleigh_scattering.h:23-43 and
// src/ecckd/ckd_model.h:368-385 are restated in oracle_extra.cpp together with
// the find_g_points interval error (a13) and the forward cost functions (a22).

// ===========================================================================
// Rows a4b, a7b, a13, a15, a21, a22 of SURVEY.md section 8.
// ===========================================================================

// rayleigh_molar_scattering_coeff: src/ecckd/rayleigh_scattering.h:23-43
// (Bucholtz 1995 two-branch power law, m2 mol-1).
ORC_API int orc_rayleigh_molar_scattering_coeff(int64_t nwav, const double* wavenumber_cm_1,
                                                double* scat_coeff) {
  const double AVOGADRO_CONSTANT = 6.02214076e23;  // constants.h:26
  for (int64_t i = 0; i < nwav; ++i) {
    const double wl = 10000.0 / wavenumber_cm_1[i];
    double s;
    if (wl < 0.5) {
      s = 3.01577e-32 * std::pow(wl, -(3.55212 + 1.35579 * wl + 0.11563 / wl));
    } else {
      s = 4.01061e-32 * std::pow(wl, -(3.99668 + 0.00110298 * wl + 0.0271393 / wl));
    }
    scat_coeff[i] = s * AVOGADRO_CONSTANT;
  }
  return ORC_OK;
}

// CkdModel::calc_rayleigh_molar_scat: src/ecckd/ckd_model.h:368-385.
// gpoint_fraction[ng][nbin], ssi_intervals[nbin], wavenumber1/2[nbin] -> rayleigh_molar_scat[ng].
ORC_API int orc_rayleigh_molar_scat_gpoints(int ng, int nbin, const double* wavenumber1,
                                            const double* wavenumber2, const double* gpoint_fraction,
                                            const double* ssi_intervals, double* rayleigh_molar_scat) {
  const double REFERENCE_COS_SZA = 0.5;  // calc_cost_function_sw.h:20
  std::vector<double> mid(nbin), scat(nbin), trans_hr(nbin);
  for (int i = 0; i < nbin; ++i) mid[i] = 0.5 * (wavenumber1[i] + wavenumber2[i]);
  orc_rayleigh_molar_scattering_coeff(nbin, mid.data(), scat.data());
  const double molar_column = 1.0e5 / (ACCEL_GRAVITY * 0.001 * MOLAR_MASS_DRY_AIR);
  for (int i = 0; i < nbin; ++i) trans_hr[i] = std::exp(-(molar_column * scat[i]) / REFERENCE_COS_SZA);
  for (int ig = 0; ig < ng; ++ig) {
    double num = 0.0, den = 0.0;
    for (int i = 0; i < nbin; ++i) {
      num += gpoint_fraction[(size_t)ig * nbin + i] * (ssi_intervals[i] * trans_hr[i]);
      den += gpoint_fraction[(size_t)ig * nbin + i] * ssi_intervals[i];
    }
    const double transmission = num / den;
    const double od = -std::log(std::max(1.0e-14, transmission)) * REFERENCE_COS_SZA;
    rayleigh_molar_scat[ig] = od / molar_column;
  }
  return ORC_OK;
}

// Base g-point split: src/ecckd/create_look_up_table.cpp:162-224.
// In/out: *ng, band_number[cap], g_point[nwav], solar_irradiance[cap] (NULL in the longwave,
// together with ssi).  Returns PROCESSING_ERROR when cap is too small.
// Note: for the LAST band the reference leaves new_band_number / new_solar_irradiance of the g
// points above the split one uninitialised (:190-193 only copies them when iband < nband-1);
// here they are copied for every band, which is what any initialised value would have to be.
ORC_API int orc_split_base_g_points(int nband, const double* band_wn1, const double* band_wn2, int nbound,
                                    const double* base_wavenumber_boundary, int64_t nwav,
                                    const double* wavenumber_cm_1, const double* ssi, int cap, int* ng_io,
                                    int32_t* band_number, int32_t* g_point, double* solar_irradiance) {
  int ng = *ng_io;
  for (int iband = 0; iband < nband; ++iband) {
    std::vector<double> inside;
    for (int k = 0; k < nbound; ++k) {
      if (base_wavenumber_boundary[k] > band_wn1[iband] && base_wavenumber_boundary[k] < band_wn2[iband]) {
        inside.push_back(base_wavenumber_boundary[k]);
      }
    }
    const int nidx = (int)inside.size();
    if (nidx == 0) continue;
    int ig = -1;
    for (int k = 0; k < ng; ++k) {
      if (band_number[k] == iband) { ig = k; break; }
    }
    if (ig < 0) return ORC_PROCESSING_ERROR;  // minval of an empty find() in the reference
    const int new_ng = ng + nidx;
    if (new_ng > cap) return ORC_PROCESSING_ERROR;
    std::vector<int32_t> new_band(new_ng);
    std::vector<double> new_sol(new_ng, 0.0);
    for (int k = 0; k <= ig; ++k) new_band[k] = band_number[k];
    for (int k = ig + 1; k <= ig + nidx; ++k) new_band[k] = iband;
    for (int k = 0; k < ig; ++k) if (solar_irradiance) new_sol[k] = solar_irradiance[k];
    for (int k = ig + 1; k < ng; ++k) {
      new_band[k + nidx] = band_number[k];
      if (solar_irradiance) new_sol[k + nidx] = solar_irradiance[k];
    }
    std::vector<double> bounds(nidx + 2);
    bounds[0] = band_wn1[iband];
    for (int k = 0; k < nidx; ++k) bounds[k + 1] = inside[k];
    bounds[nidx + 1] = band_wn2[iband];
    std::vector<int32_t> new_g(nwav);
    for (int64_t i = 0; i < nwav; ++i) new_g[i] = g_point[i] > ig ? g_point[i] + nidx : g_point[i];
    for (int new_ig = ig; new_ig <= ig + nidx; ++new_ig) {
      for (int64_t i = 0; i < nwav; ++i) {
        if (g_point[i] == ig && wavenumber_cm_1[i] >= bounds[new_ig - ig] &&
            wavenumber_cm_1[i] < bounds[new_ig - ig + 1]) {
          new_g[i] = new_ig;
        }
      }
    }
    if (solar_irradiance && ssi) {
      for (int new_ig = ig; new_ig <= ig + nidx; ++new_ig) {
        double s = 0.0;
        for (int64_t i = 0; i < nwav; ++i) if (new_g[i] == new_ig) s += ssi[i];
        new_sol[new_ig] = s;
      }
    }
    for (int k = 0; k < new_ng; ++k) {
      band_number[k] = new_band[k];
      if (solar_irradiance) solar_irradiance[k] = new_sol[k];
    }
    std::memcpy(g_point, new_g.data(), sizeof(int32_t) * nwav);
    ng = new_ng;
  }
  *ng_io = ng;
  return ORC_OK;
}

// fit_optical_depth_lw: src/ecckd/find_g_points.cpp:54-116.  metric[nlay][n], planck_hl[nlay+1][n]
// (row stride n), interval [i1,i2] inclusive.  Methods: linear, transmission, transmission-2,
// square-root, logarithmic (ids as above; others -> PARAMETER_ERROR).  Kept quirks: the weights
// are planck_hl rows 1..nlay (:62), and the logarithmic branch divides by the sum of row iz while
// weighting by row iz+1 (:88-89).
ORC_API int orc_fit_optical_depth_lw(int method, int nlay, int64_t n, int64_t i1, int64_t i2, const double* planck_hl,
                                     const double* metric, double* optical_depth_fit) {
  if (!(method == AVG_LINEAR || method == AVG_TRANSMISSION || method == AVG_TRANSMISSION_2 ||
        method == AVG_SQUARE_ROOT || method == AVG_LOGARITHMIC)) {
    return ORC_PARAMETER_ERROR;
  }
  for (int iz = 0; iz < nlay; ++iz) {
    const double* m = metric + (int64_t)iz * n;
    const double* w = planck_hl + (int64_t)(iz + 1) * n;
    if (method != AVG_LOGARITHMIC) {
      double num = 0.0, den = 0.0;
      for (int64_t i = i1; i <= i2; ++i) { num += m[i] * w[i]; den += w[i]; }
      double x = num / den;
      if (method == AVG_TRANSMISSION) {
        x = std::min(0.9999999999999999, x);
        x = std::fabs(-std::log(1.0 - x) / LW_DIFFUSIVITY);
      } else if (method == AVG_TRANSMISSION_2) {
        x = std::min(0.9999999999999999, x);
        x = std::fabs(-std::log(1.0 - x) / (LW_DIFFUSIVITY * 2.0));
      } else if (method == AVG_SQUARE_ROOT) {
        x *= x;
      }
      optical_depth_fit[iz] = x;
    } else {
      const double* wden = planck_hl + (int64_t)iz * n;
      double num = 0.0, den = 0.0;
      int64_t nnz = 0;
      for (int64_t i = i1; i <= i2; ++i) {
        if (m[i] > 0.0) { num += std::log(m[i]) * w[i]; den += wden[i]; ++nnz; }
      }
      if (nnz == i2 - i1 + 1) optical_depth_fit[iz] = std::exp(num / den);
      else if (nnz == 0) optical_depth_fit[iz] = 0.0;
      else optical_depth_fit[iz] = std::exp(num / den) * ((double)nnz / (double)(i2 - i1 + 1));
    }
  }
  return ORC_OK;
}

// fit_optical_depth_sw: src/ecckd/find_g_points.cpp:119-173.  Kept quirk: in the transmission
// branches min(0.9999999999999999, .) is applied to the un-normalised sum (:131,:136).
// AVG_TOTAL_TRANSMISSION takes the logarithmic branch (:146-147).
ORC_API int orc_fit_optical_depth_sw(int method, int nlay, int64_t n, int64_t i1, int64_t i2, const double* ssi,
                                     const double* metric, double* optical_depth_fit) {
  if (!(method == AVG_LINEAR || method == AVG_TRANSMISSION || method == AVG_TRANSMISSION_2 ||
        method == AVG_SQUARE_ROOT || method == AVG_LOGARITHMIC || method == AVG_TOTAL_TRANSMISSION)) {
    return ORC_PARAMETER_ERROR;
  }
  double ssum = 0.0;
  for (int64_t i = i1; i <= i2; ++i) ssum += ssi[i];
  const double norm_factor = 1.0 / ssum;
  for (int iz = 0; iz < nlay; ++iz) {
    const double* m = metric + (int64_t)iz * n;
    if (method != AVG_LOGARITHMIC && method != AVG_TOTAL_TRANSMISSION) {
      double num = 0.0;
      for (int64_t i = i1; i <= i2; ++i) num += m[i] * ssi[i];
      double x;
      if (method == AVG_TRANSMISSION) {
        x = std::min(0.9999999999999999, num) * norm_factor;
        x = std::fabs(-std::log(1.0 - x) / LW_DIFFUSIVITY);
      } else if (method == AVG_TRANSMISSION_2) {
        x = std::min(0.9999999999999999, num) * norm_factor;
        x = std::fabs(-std::log(1.0 - x) / (LW_DIFFUSIVITY * 2.0));
      } else if (method == AVG_SQUARE_ROOT) {
        x = num * norm_factor;
        x *= x;
      } else {
        x = num * norm_factor;
      }
      optical_depth_fit[iz] = x;
    } else {
      double num = 0.0, den = 0.0;
      int64_t nnz = 0;
      for (int64_t i = i1; i <= i2; ++i) {
        if (m[i] > 0.0) { num += std::log(m[i]) * ssi[i]; den += ssi[i]; ++nnz; }
      }
      if (nnz == i2 - i1 + 1) optical_depth_fit[iz] = std::exp(num / den);
      else if (nnz == 0) optical_depth_fit[iz] = 0.0;
      else optical_depth_fit[iz] = std::exp(num / den) * ((double)nnz / (double)(i2 - i1 + 1));
    }
  }
  return ORC_OK;
}

// fit_optical_depth_sw_total_trans: src/ecckd/find_g_points.cpp:176-204.  Kept quirk: when a
// broadband flux underflows to zero the WHOLE profile is overwritten by the linear average (:197),
// and later layers may overwrite their own entry again.
ORC_API int orc_fit_optical_depth_sw_total_trans(int nlay, int64_t n, int64_t i1, int64_t i2, const double* ssi,
                                                 const double* bg_od, const double* od, double* optical_depth_fit) {
  const int64_t m = i2 - i1 + 1;
  std::vector<double> flux_dn(m), bg_flux_dn(m);
  double top = 0.0;
  for (int64_t i = 0; i < m; ++i) { flux_dn[i] = bg_flux_dn[i] = ssi[i1 + i]; top += ssi[i1 + i]; }
  double bb_flux_dn_top = top, bb_bg_flux_dn_top = top;
  const double norm_factor = 1.0 / top;
  for (int iz = 0; iz < nlay; ++iz) optical_depth_fit[iz] = 0.0;  // Vector(nz): unset in the reference
  for (int iz = 0; iz < nlay; ++iz) {
    double bb_bg = 0.0, bb = 0.0;
    for (int64_t i = 0; i < m; ++i) {
      const double b = bg_od[(int64_t)iz * n + i1 + i], o = od[(int64_t)iz * n + i1 + i];
      bg_flux_dn[i] *= std::exp(-2.0 * b);
      flux_dn[i] *= std::exp(-2.0 * (b + o));
    }
    for (int64_t i = 0; i < m; ++i) bb_bg += bg_flux_dn[i];
    for (int64_t i = 0; i < m; ++i) bb += flux_dn[i];
    if (bb_bg > 0.0 && bb > 0.0) {
      const double bg_od_fit = -0.5 * std::log(bb_bg / bb_bg_flux_dn_top);
      optical_depth_fit[iz] = -0.5 * std::log(bb / bb_flux_dn_top) - bg_od_fit;
    } else {
      for (int jz = 0; jz < nlay; ++jz) {
        double s = 0.0;
        for (int64_t i = 0; i < m; ++i) s += od[(int64_t)jz * n + i1 + i] * ssi[i1 + i];
        optical_depth_fit[jz] = s * norm_factor;
      }
    }
    bb_flux_dn_top = bb;
    bb_bg_flux_dn_top = bb_bg;
  }
  return ORC_OK;
}

// calc_cost_function_lw: src/ecckd/calc_cost_function_lw.cpp:23-110 (index empty: the caller
// passes interval slices).  Arrays hold the n wavenumbers of the interval.
ORC_API int orc_calc_cost_function_lw(int nlay, int64_t n, const double* pressure_hl, const double* planck_hl,
                                      const double* surf_emissivity, const double* surf_planck,
                                      const double* bg_optical_depth, const double* optical_depth_fit,
                                      const double* flux_dn_surf, const double* flux_up_toa, const double* hr,
                                      double flux_weight, const double* layer_weight, double* cost) {
  const double hr_weight = 3600.0 * 24.0;
  std::vector<double> hr_true(nlay), dn(nlay + 1), up(nlay + 1), hr_fit(nlay);
  for (int iz = 0; iz < nlay; ++iz) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += hr[(int64_t)iz * n + i];
    hr_true[iz] = s;
  }
  double dn_true = 0.0, up_true = 0.0;
  for (int64_t i = 0; i < n; ++i) dn_true += flux_dn_surf[i];
  for (int64_t i = 0; i < n; ++i) up_true += flux_up_toa[i];
  orc_radiative_transfer_lw_bb(nlay, n, planck_hl, bg_optical_depth, optical_depth_fit, surf_emissivity, surf_planck,
                               dn.data(), up.data());
  orc_heating_rate_single(nlay, pressure_hl, dn.data(), up.data(), hr_fit.data());
  double s = 0.0;
  for (int iz = 0; iz < nlay; ++iz) s += layer_weight[iz] * ((hr_fit[iz] - hr_true[iz]) * (hr_fit[iz] - hr_true[iz]));
  *cost = std::sqrt(hr_weight * hr_weight * s +
                    flux_weight * ((dn[nlay] - dn_true) * (dn[nlay] - dn_true) + (up[0] - up_true) * (up[0] - up_true)));
  return ORC_OK;
}

// calc_cost_function_sw: src/ecckd/calc_cost_function_sw.cpp:20-110.  albedo <= 0: direct beam
// only and flux_up_fit stays 0 (:56); the fitted heating rate never includes the upwelling (:87).
ORC_API int orc_calc_cost_function_sw(int nlay, int64_t n, double cos_sza, const double* pressure_hl, const double* ssi,
                                      double albedo, const double* bg_optical_depth, const double* optical_depth_fit,
                                      const double* flux_dn_surf, const double* flux_up_toa, const double* hr,
                                      double flux_weight, const double* layer_weight, double* cost) {
  const double hr_weight = 3600.0 * 24.0;
  std::vector<double> hr_true(nlay), dn(nlay + 1), up(nlay + 1, 0.0), hr_fit(nlay);
  for (int iz = 0; iz < nlay; ++iz) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += hr[(int64_t)iz * n + i];
    hr_true[iz] = s;
  }
  double dn_true = 0.0, up_true = 0.0;
  for (int64_t i = 0; i < n; ++i) dn_true += flux_dn_surf[i];
  for (int64_t i = 0; i < n; ++i) up_true += flux_up_toa[i];
  orc_radiative_transfer_sw_bb(nlay, n, cos_sza, ssi, bg_optical_depth, optical_depth_fit, albedo, dn.data(),
                               albedo <= 0.0 ? (double*)0 : up.data());
  orc_heating_rate_single(nlay, pressure_hl, dn.data(), (const double*)0, hr_fit.data());
  double s = 0.0;
  for (int iz = 0; iz < nlay; ++iz) s += layer_weight[iz] * ((hr_fit[iz] - hr_true[iz]) * (hr_fit[iz] - hr_true[iz]));
  *cost = std::sqrt(hr_weight * hr_weight * s +
                    flux_weight * ((dn[nlay] - dn_true) * (dn[nlay] - dn_true) + (up[0] - up_true) * (up[0] - up_true)));
  return ORC_OK;
}

// Shared tail of calc_cost_function_ckd_{lw,sw}: band sums, heating rates and the cost terms.
// fdo/fuo[nlay+1][ng] are the forward g-point fluxes (already minus the relative-to fluxes).
static double ckd_cost_terms(bool sw, bool up_in_bb, int nlay, int ng, int nband, const int32_t* band_mapping,
                             const double* pressure_hl, const std::vector<double>& fdo, const std::vector<double>& fuo,
                             const double* flux_dn, const double* flux_up, const double* hr, double flux_weight,
                             double flux_profile_weight, double broadband_weight, const double* layer_weight,
                             double* cost_fn_per_band) {
  const double hr_weight = 3600.0 * 24.0;
  const int nlev = nlay + 1;
  std::vector<double> fd((size_t)nlev * nband), fu((size_t)nlev * nband), hrf((size_t)nlay * nband);
  if (!band_mapping) {
    fd = fdo;
    fu = fuo;
  } else {
    for (int ib = 0; ib < nband; ++ib) {
      for (int il = 0; il < nlev; ++il) {
        double sd = 0.0, su = 0.0;
        for (int ig = 0; ig < ng; ++ig) {
          if (band_mapping[ig] == ib) { sd += fdo[(size_t)il * ng + ig]; su += fuo[(size_t)il * ng + ig]; }
        }
        fd[(size_t)il * nband + ib] = sd;
        fu[(size_t)il * nband + ib] = su;
      }
    }
  }
  // heating_rate(): LW with upwelling, SW with aMatrix() i.e. downwelling only (calc_cost_function_sw.cpp:195)
  orc_heating_rate(nlay, nband, pressure_hl, fd.data(), sw ? (const double*)0 : fu.data(), hrf.data());
  std::vector<double> iw(nlay > 1 ? nlay - 1 : 0);
  for (int iz = 0; iz + 1 < nlay; ++iz) iw[iz] = flux_profile_weight * 0.5 * (layer_weight[iz] + layer_weight[iz + 1]);
  double cost = 0.0;
  for (int ib = 0; ib < nband; ++ib) {
    double s = 0.0;
    for (int iz = 0; iz < nlay; ++iz) {
      const double d = hrf[(size_t)iz * nband + ib] - hr[(size_t)iz * nband + ib];
      s += layer_weight[iz] * d * d;
    }
    const double dd = fd[(size_t)nlay * nband + ib] - flux_dn[(size_t)nlay * nband + ib];
    const double du = fu[ib] - flux_up[ib];
    double local = hr_weight * hr_weight * s + flux_weight * (dd * dd + (sw ? 20.0 : 1.0) * du * du);
    if (flux_profile_weight > 0.0) {
      double t = 0.0;
      for (int il = 1; il < nlay; ++il) {
        const double a = fd[(size_t)il * nband + ib] - flux_dn[(size_t)il * nband + ib];
        const double b = fu[(size_t)il * nband + ib] - flux_up[(size_t)il * nband + ib];
        t += iw[il - 1] * (a * a + b * b);
      }
      local += t;
    }
    cost += local;
    if (cost_fn_per_band) cost_fn_per_band[ib] += local;
  }
  if (!sw || broadband_weight > 0.0) {  // the LW applies the mix unconditionally (:205), the SW if > 0 (:243)
    double s = 0.0;
    for (int iz = 0; iz < nlay; ++iz) {
      double e = 0.0;
      for (int ib = 0; ib < nband; ++ib) e += hrf[(size_t)iz * nband + ib] - hr[(size_t)iz * nband + ib];
      s += layer_weight[iz] * (e * e);
    }
    double ed = 0.0, eu = 0.0;
    for (int ib = 0; ib < nband; ++ib) ed += fd[(size_t)nlay * nband + ib] - flux_dn[(size_t)nlay * nband + ib];
    for (int ib = 0; ib < nband; ++ib) eu += fu[ib] - flux_up[ib];
    if (!sw) {
      cost = (cost * (1.0 - broadband_weight)) / nband + broadband_weight * hr_weight * hr_weight * s +
             broadband_weight * flux_weight * (ed * ed + eu * eu);
    } else {
      cost = (cost * (1.0 - broadband_weight)) / nband + broadband_weight * hr_weight * hr_weight * s;
      cost += broadband_weight * flux_weight * (ed * ed);
      if (up_in_bb) cost += broadband_weight * flux_weight * (eu * eu);
    }
    if (flux_profile_weight > 0.0) {
      double td = 0.0, tu = 0.0;
      for (int il = 1; il < nlay; ++il) {
        double a = 0.0, b = 0.0;
        for (int ib = 0; ib < nband; ++ib) {
          a += fd[(size_t)il * nband + ib] - flux_dn[(size_t)il * nband + ib];
          b += fu[(size_t)il * nband + ib] - flux_up[(size_t)il * nband + ib];
        }
        td += iw[il - 1] * (a * a);
        tu += iw[il - 1] * (b * b);
      }
      if (!sw) {
        double t = 0.0;  // one sum of (dn^2 + up^2) in the LW (:216-217)
        for (int il = 1; il < nlay; ++il) {
          double a = 0.0, b = 0.0;
          for (int ib = 0; ib < nband; ++ib) {
            a += fd[(size_t)il * nband + ib] - flux_dn[(size_t)il * nband + ib];
            b += fu[(size_t)il * nband + ib] - flux_up[(size_t)il * nband + ib];
          }
          t += iw[il - 1] * (a * a + b * b);
        }
        cost += broadband_weight * t;
      } else {
        cost += broadband_weight * td;
        if (up_in_bb) cost += broadband_weight * tu;
      }
    }
  }
  return cost;
}

// calc_cost_function_ckd_lw (forward value): src/ecckd/calc_cost_function_lw.cpp:115-232.
ORC_API int orc_calc_cost_function_ckd_lw(int nlay, int ng, int nband, const double* pressure_hl, const double* planck_hl,
                                          const double* surf_emissivity, const double* surf_planck,
                                          const double* optical_depth, const double* flux_dn, const double* flux_up,
                                          const double* hr, const double* spectral_flux_dn_surf,
                                          const double* spectral_flux_up_toa, double flux_weight,
                                          double flux_profile_weight, double broadband_weight,
                                          double spectral_boundary_weight, const double* layer_weight,
                                          const double* relative_ckd_flux_dn, const double* relative_ckd_flux_up,
                                          const int32_t* band_mapping, double* cost) {
  const int nlev = nlay + 1;
  if (!band_mapping && nband != ng) return ORC_PARAMETER_ERROR;
  std::vector<double> emis(ng), fdo((size_t)nlev * ng), fuo((size_t)nlev * ng);
  for (int ig = 0; ig < ng; ++ig) emis[ig] = band_mapping ? surf_emissivity[band_mapping[ig]] : surf_emissivity[ig];
  orc_radiative_transfer_lw(nlay, ng, planck_hl, optical_depth, emis.data(), surf_planck, fdo.data(), fuo.data());
  if (relative_ckd_flux_dn) {
    for (size_t k = 0; k < fdo.size(); ++k) { fdo[k] -= relative_ckd_flux_dn[k]; fuo[k] -= relative_ckd_flux_up[k]; }
  }
  double c = ckd_cost_terms(false, true, nlay, ng, nband, band_mapping, pressure_hl, fdo, fuo, flux_dn, flux_up, hr,
                            flux_weight, flux_profile_weight, broadband_weight, layer_weight, (double*)0);
  if (spectral_boundary_weight > 0.0 && spectral_flux_dn_surf && spectral_flux_up_toa) {
    double s = 0.0;
    for (int ig = 0; ig < ng; ++ig) {
      const double a = fdo[(size_t)nlay * ng + ig] - spectral_flux_dn_surf[ig];
      const double b = fuo[ig] - spectral_flux_up_toa[ig];
      s += a * a + b * b;
    }
    c += spectral_boundary_weight * s;
  }
  *cost = c;
  return ORC_OK;
}

// calc_cost_function_ckd_sw (forward value): src/ecckd/calc_cost_function_sw.cpp:115-277.
// The ">0.1 W m-2 incoming flux" warning (:232-237) is not part of the value and is not restated.
ORC_API int orc_calc_cost_function_ckd_sw(int nlay, int ng, int nband, double cos_sza, const double* pressure_hl,
                                          const double* ssi, int nalbedo, const double* albedo,
                                          const double* optical_depth, const double* flux_dn, const double* flux_up,
                                          const double* hr, const double* spectral_flux_dn_surf,
                                          const double* spectral_flux_up_toa, double flux_weight,
                                          double flux_profile_weight, double broadband_weight,
                                          const double* spectral_boundary_weights, const double* layer_weight,
                                          const double* relative_ckd_flux_dn, const double* relative_ckd_flux_up,
                                          const int32_t* band_mapping, double* cost_fn_per_band, double* cost) {
  (void)spectral_flux_up_toa;
  const int nlev = nlay + 1;
  if (!band_mapping && nband != ng) return ORC_PARAMETER_ERROR;
  bool all_le0 = true, all_gt0 = true;
  for (int k = 0; k < nalbedo; ++k) {
    if (!(albedo[k] <= 0.0)) all_le0 = false;
    if (!(albedo[k] > 0.0)) all_gt0 = false;
  }
  std::vector<double> fdo((size_t)nlev * ng), fuo((size_t)nlev * ng, 0.0);
  if (all_le0) {
    orc_radiative_transfer_direct_sw(nlay, ng, cos_sza, ssi, optical_depth, fdo.data());
  } else {
    if (!band_mapping) return ORC_PARAMETER_ERROR;  // albedo(band_mapping) would be empty: Adept size mismatch
    std::vector<double> alb(ng);
    for (int ig = 0; ig < ng; ++ig) alb[ig] = albedo[band_mapping[ig]];
    orc_radiative_transfer_norayleigh_sw(nlay, ng, cos_sza, ssi, optical_depth, alb.data(), fdo.data(), fuo.data());
  }
  if (relative_ckd_flux_dn) {
    for (size_t k = 0; k < fdo.size(); ++k) { fdo[k] -= relative_ckd_flux_dn[k]; fuo[k] -= relative_ckd_flux_up[k]; }
  }
  double c = ckd_cost_terms(true, all_gt0, nlay, ng, nband, band_mapping, pressure_hl, fdo, fuo, flux_dn, flux_up, hr,
                            flux_weight, flux_profile_weight, broadband_weight, layer_weight, cost_fn_per_band);
  if (spectral_boundary_weights && spectral_flux_dn_surf) {
    double s = 0.0;
    for (int ig = 0; ig < ng; ++ig) {
      const double a = fdo[(size_t)nlay * ng + ig] - spectral_flux_dn_surf[ig];
      s += spectral_boundary_weights[ig] * a * a;
    }
    c += s;
  }
  *cost = c;
  return ORC_OK;
}

// Background (prior) term of the optimize_lut cost function for one active gas:
// CkdModel<true>::calc_background_cost_function, ckd_model.cpp:838-866.  The gas's slice of the state vector is
// [nx][ng] with the g point the fastest index (:851-855); per g point
//   gradient_local = (1 / background_error(g)^2) * (inv_background_shape ** delta_x_local)      (:857-858)
//   cost          += 0.5 * dot_product(delta_x_local, gradient_local)                           (:859,864)
//   gradient      += gradient_local                                                             (:861)
// gradient is accumulated into (the reference zeroes the whole vector once, :842, before the gas loop).
ORC_API int orc_background_cost_gas(int nx, int ng, const double* inv_background_shape, const double* background_error,
                                    const double* delta_x, double* gradient, double* cost) {
  double c = 0.0;
  std::vector<double> gl((size_t)nx);
  for (int ig = 0; ig < ng; ++ig) {
    const double inv_err2 = 1.0 / (background_error[ig] * background_error[ig]);
    for (int i = 0; i < nx; ++i) {
      double s = 0.0;
      for (int j = 0; j < nx; ++j) s += inv_background_shape[(size_t)i * nx + j] * delta_x[(size_t)j * ng + ig];
      gl[i] = inv_err2 * s;
    }
    double d = 0.0;
    for (int i = 0; i < nx; ++i) d += delta_x[(size_t)i * ng + ig] * gl[i];
    c += 0.5 * d;
    for (int i = 0; i < nx; ++i) gradient[(size_t)i * ng + ig] += gl[i];
  }
  *cost = c;
  return ORC_OK;
}

// The Rayleigh part, ckd_model.cpp:868-873: the cost only (the reference forms gradient_local and drops it).
ORC_API int orc_background_cost_rayleigh(int ng, const double* rayleigh_inv_background, const double* delta_x,
                                         double* cost) {
  double d = 0.0;
  for (int ig = 0; ig < ng; ++ig) d += delta_x[ig] * (rayleigh_inv_background[ig] * delta_x[ig]);
  *cost = 0.5 * d;
  return ORC_OK;
}
