// cost_function.cu - the error / cost evaluations built on the flux solvers:
//   * broadband solvers radiative_transfer_lw_bb, radiative_transfer_{direct,norayleigh}_sw_bb
//     (radiative_transfer_lw.cpp:85-142, radiative_transfer_sw.cpp:116-184),
//   * the candidate-interval error of find_g_points, batched over intervals:
//     fit_optical_depth_{lw,sw,sw_total_trans} (find_g_points.cpp:54-204),
//     calc_cost_function_{lw,sw} (calc_cost_function_lw.cpp:23-110, calc_cost_function_sw.cpp:20-110),
//     CkdEquipartition::calc_error (find_g_points.cpp:291-405),
//   * the forward cost of optimize_lut, batched over profiles:
//     calc_cost_function_ckd_{lw,sw} (calc_cost_function_lw.cpp:115-232, calc_cost_function_sw.cpp:115-277)
//     with the layer weights and negative-optical-depth penalty of solve_adept.cpp:107-116,133-145.
//
// Interval batches.  The spectral arrays are [layer][n] in rank order, so a candidate interval is a
// contiguous slab [i1,i2].  Every interval is cut into tiles of 256 wavenumbers; one CTA handles one
// (interval, tile) item and writes partial sums, which a second tiny kernel adds in tile order:
// the result does not depend on the launch geometry and is bitwise reproducible.  All kernels are
// HBM streams (16-24 B per wavenumber x layer).
#include <algorithm>

#include "common.cuh"
#include "trans_core.cuh"

namespace ecckd {

constexpr int kTile = 256;  // wavenumbers per item == threads per CTA
constexpr int kWarps = kTile / 32;

struct IntervalParams {
  int sw, method, nlay, nint, nitems;
  int64_t n, stride;
  const double* pressure_hl;      // [nlay+1]
  const double* planck_hl;        // LW [nlay+1][n]
  const double* ssi;              // SW [n]
  const double* surf_emissivity;  // LW [n]
  const double* surf_planck;      // LW [n]
  const double* bg_od;            // [nlay][n]
  const double* metric;           // [nlay][n]
  const double* flux_dn_surf;     // [n]
  const double* flux_up_toa;      // [n]
  const double* hr;               // [nlay][n]
  const double* layer_weight;     // [nlay]
  double cos_sza, albedo, flux_weight, od_scale;
  const int64_t* i1;          // [nint] inclusive bounds
  const int64_t* i2;          // [nint]
  const int64_t* item_start;  // [nint+1] first item of each interval
  double* part_fit;           // [nitems][nlay][4]: fit numerator, denominator, count of tau>0, sum of hr
  double* part_scal;          // [nitems][4]: sum flux_dn_surf, sum flux_up_toa, sum ssi
  double* part_flux;          // [nitems][2][nlay+1]
  double* fit;                // [nint][nlay]
  double* truth;              // [nint][nlay+4]: hr_true[nlay], flux_dn_surf_true, flux_up_toa_true, sum ssi
  double* bb_flux;            // [nint][2][nlay+1] (optional output)
  double* error;              // [nint]
  int error_mode;             // 0: store, 1: store 0.5*, 2: add 0.5*  (total-transmission low/high mean)
  TransConst256 tc;           // 1 - exp(-LW_DIFFUSIVITY tau) by the table scheme of trans_core.cuh
};

// emissivity and Planck-gradient factor of radiative_transfer_lw_bb (radiative_transfer_lw.cpp:113-119)
// with the table exponential and a Newton reciprocal instead of libm exp and a division (the broadband
// fluxes change by parts in 1e13)
__device__ __forceinline__ void lw_bb_layer(double tau, const TransConst256& K, const unsigned char* tab, double& e,
                                            double& f) {
  const double THRESHOLD_EMISSIVITY = 1.0e-5;  // radiative_transfer_lw.cpp:98
  e = trans256_one_minus_exp(tau, K, tab);
  const double den = fmax(tau, THRESHOLD_EMISSIVITY / ECCKD_LW_DIFFUSIVITY);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(r, fma(-den, r, 1.0), r);
  r = fma(r, fma(-den, r, 1.0), r);
  f = fmax(1.0 - ((1.0 / ECCKD_LW_DIFFUSIVITY) * fmax(e, THRESHOLD_EMISSIVITY)) * r, 0.5 * THRESHOLD_EMISSIVITY);
}

__device__ __forceinline__ int find_interval(const int64_t* __restrict__ item_start, int nint, int64_t item) {
  int lo = 0, hi = nint - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (item_start[mid] <= item) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- phase A: per-tile partial sums of the fit and of the "true" quantities.
// Warp w handles layers w, w+8, ...; lanes stride the tile (coalesced 256 B rows).
__global__ void __launch_bounds__(kTile) interval_fit_partial_kernel(const IntervalParams p) {
  const int64_t item = blockIdx.x;
  const int k = find_interval(p.item_start, p.nint, item);
  const int64_t a = p.i1[k] + (item - p.item_start[k]) * kTile;
  const int64_t b = min(a + kTile, p.i2[k] + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool logm = p.method == ECCKD_AVG_LOGARITHMIC || (p.sw && p.method == ECCKD_AVG_TOTAL_TRANSMISSION);
  const bool tt_linear = p.sw && p.method == ECCKD_AVG_TOTAL_TRANSMISSION;  // also wants the linear sums (fallback)
  for (int l = warp; l < p.nlay; l += kWarps) {
    const double* m = p.metric + (int64_t)l * p.stride;
    const double* h = p.hr + (int64_t)l * p.stride;
    // weights: LW planck_hl row l+1 (find_g_points.cpp:62); SW ssi
    const double* w = p.sw ? p.ssi : p.planck_hl + (int64_t)(l + 1) * p.stride;
    const double* wden = p.sw ? p.ssi : p.planck_hl + (int64_t)l * p.stride;  // log branch divides by row l (:89)
    double num = 0.0, den = 0.0, cnt = 0.0, hs = 0.0;
    for (int64_t i = a + lane; i < b; i += 32) {
      const double mv = m[i];
      hs += h[i];
      if (logm && !tt_linear) {
        if (mv > 0.0) { num += log(mv) * w[i]; den += wden[i]; cnt += 1.0; }
      } else {
        num += mv * w[i];
        den += w[i];
      }
    }
    num = warp_sum(num); den = warp_sum(den); cnt = warp_sum(cnt); hs = warp_sum(hs);
    if (lane == 0) {
      double* o = p.part_fit + ((size_t)item * p.nlay + l) * 4;
      o[0] = num; o[1] = den; o[2] = cnt; o[3] = hs;
    }
  }
  if (warp == 0) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = a + lane; i < b; i += 32) {
      s0 += p.flux_dn_surf[i];
      s1 += p.flux_up_toa[i];
      if (p.sw) s2 += p.ssi[i];
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
      double* o = p.part_scal + (size_t)item * 4;
      o[0] = s0; o[1] = s1; o[2] = s2; o[3] = 0.0;
    }
  }
}

// Fixed-order sum over the tiles of an interval: the tiles are cut into kSeg contiguous segments (a
// function of the tile count only), every segment is added in tile order by its own thread with eight
// loads in flight, and the segment sums are added in segment order.  Bitwise reproducible and
// independent of the launch geometry.  (One thread walking all 1024 tiles of a 2^18-point interval
// made the two CTA-per-interval kernels 0.76 ms of a 3.3 ms batch.)
constexpr int kSeg = 8;
__device__ __forceinline__ void seg_bounds(int64_t it0, int64_t it1, int seg, int64_t& a, int64_t& b) {
  const int64_t t = it1 - it0;
  a = it0 + (t * seg) / kSeg;
  b = it0 + (t * (seg + 1)) / kSeg;
}
__device__ __forceinline__ double seg_sum(const double* __restrict__ base, size_t stride, int64_t a, int64_t b) {
  double s = 0.0;
  int64_t it = a;
  for (; it + 8 <= b; it += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = base[(size_t)(it + q) * stride];
#pragma unroll
    for (int q = 0; q < 8; ++q) s += v[q];
  }
  for (; it < b; ++it) s += base[(size_t)it * stride];
  return s;
}

// ---- phase A2: add the tile partials in tile order and apply the fit formula.  CTA = interval.
// Dynamic shared memory: kSeg*(nlay*4 + 4) doubles (+ nlay for the total-transmission fit).
__global__ void interval_fit_finish_kernel(const IntervalParams p, int fit_from_sweep) {
  extern __shared__ double s_dyn[];
  double* s_seg = s_dyn;                           // [kSeg][nlay][4]
  double* s_sseg = s_dyn + (size_t)kSeg * p.nlay * 4;  // [kSeg][4]
  double* s_lin = s_sseg + kSeg * 4;               // [nlay]
  const int k = blockIdx.x;
  const int64_t it0 = p.item_start[k], it1 = p.item_start[k + 1];
  const int64_t npts = p.i2[k] - p.i1[k] + 1;
  __shared__ double s_scal[3];
  for (int idx = threadIdx.x; idx < kSeg * p.nlay * 4 + kSeg * 3; idx += blockDim.x) {
    int64_t a, b;
    if (idx < kSeg * p.nlay * 4) {
      const int seg = idx / (p.nlay * 4), rem = idx - seg * (p.nlay * 4);  // rem = l*4 + component
      seg_bounds(it0, it1, seg, a, b);
      s_seg[idx] = seg_sum(p.part_fit + rem, (size_t)p.nlay * 4, a, b);
    } else {
      const int j = idx - kSeg * p.nlay * 4, seg = j / 3, c = j - seg * 3;
      seg_bounds(it0, it1, seg, a, b);
      s_sseg[seg * 4 + c] = seg_sum(p.part_scal + c, 4, a, b);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int seg = 0; seg < kSeg; ++seg) s += s_sseg[seg * 4 + threadIdx.x];
    s_scal[threadIdx.x] = s;
    p.truth[(size_t)k * (p.nlay + 4) + p.nlay + threadIdx.x] = s;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < p.nlay; l += blockDim.x) {
    double num = 0.0, den = 0.0, cnt = 0.0, hs = 0.0;
    for (int seg = 0; seg < kSeg; ++seg) {
      const double* o = s_seg + ((size_t)seg * p.nlay + l) * 4;
      num += o[0]; den += o[1]; cnt += o[2]; hs += o[3];
    }
    p.truth[(size_t)k * (p.nlay + 4) + l] = hs;
    if (fit_from_sweep == 2) continue;  // only the truth sums were wanted
    double x;
    if (!p.sw) {
      switch (p.method) {
        case ECCKD_AVG_LINEAR: x = num / den; break;
        case ECCKD_AVG_TRANSMISSION:
          x = fmin(0.9999999999999999, num / den);
          x = fabs(-log(1.0 - x) / ECCKD_LW_DIFFUSIVITY);
          break;
        case ECCKD_AVG_TRANSMISSION_2:
          x = fmin(0.9999999999999999, num / den);
          x = fabs(-log(1.0 - x) / (ECCKD_LW_DIFFUSIVITY * 2.0));
          break;
        case ECCKD_AVG_SQUARE_ROOT: x = num / den; x *= x; break;
        default:  // logarithmic (find_g_points.cpp:80-103)
          if (cnt == (double)npts) x = exp(num / den);
          else if (cnt == 0.0) x = 0.0;
          else x = exp(num / den) * (cnt / (double)npts);
      }
    } else {
      const double norm_factor = 1.0 / s_scal[2];
      switch (p.method) {
        case ECCKD_AVG_LINEAR: x = num * norm_factor; break;
        case ECCKD_AVG_TRANSMISSION:  // min() before the normalisation, as the reference (:131)
          x = fmin(0.9999999999999999, num) * norm_factor;
          x = fabs(-log(1.0 - x) / ECCKD_LW_DIFFUSIVITY);
          break;
        case ECCKD_AVG_TRANSMISSION_2:
          x = fmin(0.9999999999999999, num) * norm_factor;
          x = fabs(-log(1.0 - x) / (ECCKD_LW_DIFFUSIVITY * 2.0));
          break;
        case ECCKD_AVG_SQUARE_ROOT: x = num * norm_factor; x *= x; break;
        case ECCKD_AVG_TOTAL_TRANSMISSION: x = num * norm_factor; break;  // linear fallback profile, see below
        default:
          if (cnt == (double)npts) x = exp(num / den);
          else if (cnt == 0.0) x = 0.0;
          else x = exp(num / den) * (cnt / (double)npts);
      }
    }
    p.fit[(size_t)k * p.nlay + l] = x;
  }
  if (!fit_from_sweep) return;
  // fit_optical_depth_sw_total_trans (find_g_points.cpp:176-204): part_flux holds, per tile and level,
  // the broadband direct flux with the background only [0] and with the gas added [1].
  __syncthreads();
  if (threadIdx.x == 0 && fit_from_sweep == 1) {
    const int nlev = p.nlay + 1;
    double top_bg = s_scal[2], top = s_scal[2];
    bool fallback_used = false;
    // the linear profile is already in fit[]; keep a copy semantics: layers not overwritten later keep it
    // only if a fallback happened at or before... the reference overwrites the WHOLE profile at the failing
    // layer and then continues, so layers after the failure are set individually again if they succeed.
    // Walk the layers in order, tracking what the reference's vector would contain.
    for (int l = 0; l < p.nlay; ++l) s_lin[l] = p.fit[(size_t)k * p.nlay + l];
    for (int l = 0; l < p.nlay; ++l) p.fit[(size_t)k * p.nlay + l] = 0.0;
    for (int l = 0; l < p.nlay; ++l) {
      double base_bg = 0.0, base = 0.0;
      for (int64_t it = it0; it < it1; ++it) {
        base_bg += p.part_flux[((size_t)it * 2 + 0) * nlev + l + 1];
        base += p.part_flux[((size_t)it * 2 + 1) * nlev + l + 1];
      }
      if (base_bg > 0.0 && base > 0.0) {
        const double bg_fit = -0.5 * log(base_bg / top_bg);
        p.fit[(size_t)k * p.nlay + l] = -0.5 * log(base / top) - bg_fit;
      } else {
        for (int j = 0; j < p.nlay; ++j) p.fit[(size_t)k * p.nlay + j] = s_lin[j];
        fallback_used = true;
      }
      top = base;
      top_bg = base_bg;
    }
    (void)fallback_used;
  }
}

// ---- sweep for the total-transmission fit: direct beam at secant 2 through background / background+gas.
__global__ void __launch_bounds__(kTile) interval_tt_sweep_kernel(const IntervalParams p) {
  extern __shared__ double s_part[];  // [2][nlev][kWarps]
  const int64_t item = blockIdx.x;
  const int k = find_interval(p.item_start, p.nint, item);
  const int64_t i = p.i1[k] + (item - p.item_start[k]) * kTile + threadIdx.x;
  const bool on = i <= p.i2[k];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nlev = p.nlay + 1;
  double f_bg = on ? p.ssi[i] : 0.0, f = f_bg;
  for (int l = 0; l < p.nlay; ++l) {
    if (on) {
      const double b = p.bg_od[(int64_t)l * p.stride + i], o = p.metric[(int64_t)l * p.stride + i];
      f_bg *= exp(-2.0 * b);
      f *= exp(-2.0 * (b + o));
    }
    const double s0 = warp_sum(f_bg), s1 = warp_sum(f);
    if (lane == 0) {
      s_part[(0 * nlev + l + 1) * kWarps + warp] = s0;
      s_part[(1 * nlev + l + 1) * kWarps + warp] = s1;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * nlev; j += blockDim.x) {
    double s = 0.0;
    if (j % nlev != 0)
      for (int w = 0; w < kWarps; ++w) s += s_part[j * kWarps + w];
    p.part_flux[(size_t)item * 2 * nlev + j] = s;
  }
}

// ---- phase B: broadband flux sweep of one tile (radiative_transfer_lw_bb / _sw_bb).
template <bool SW>
__global__ void __launch_bounds__(kTile) interval_flux_kernel(const IntervalParams p) {
  extern __shared__ double s_part[];  // [2][nlev][kWarps], then the 2^(j/256) table (longwave)
  const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_part + 2 * (p.nlay + 1) * kWarps);
  if (!SW) {
    trans_tab_init(s_part + 2 * (p.nlay + 1) * kWarps);
    __syncthreads();
  }
  const int64_t item = blockIdx.x;
  const int k = find_interval(p.item_start, p.nint, item);
  const int64_t i = p.i1[k] + (item - p.item_start[k]) * kTile + threadIdx.x;
  const bool on = i <= p.i2[k];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nlev = p.nlay + 1;
  const double* fit = p.fit + (size_t)k * p.nlay;
  const double sc = p.od_scale;
  double flux = 0.0;
  if (!SW) {
    double b_top = on ? p.planck_hl[i] : 0.0;
    if (lane == 0) s_part[(0 * nlev + 0) * kWarps + warp] = 0.0;
    for (int l = 0; l < p.nlay; ++l) {
      double b_bot = 0.0;
      if (on) {
        const double tau = p.bg_od[(int64_t)l * p.stride + i] + fit[l] * sc;
        double e, f;
        lw_bb_layer(tau, p.tc, tab, e, f);
        b_bot = p.planck_hl[(int64_t)(l + 1) * p.stride + i];
        flux = flux * (1.0 - e) + b_top * (e - f) + b_bot * f;
      }
      const double s = warp_sum(flux);
      if (lane == 0) s_part[(0 * nlev + l + 1) * kWarps + warp] = s;
      b_top = b_bot;
    }
    double b_bot = b_top;
    if (on) {
      const double es = p.surf_emissivity[i];
      flux = p.surf_planck[i] * es + (1.0 - es) * flux;
    }
    {
      const double s = warp_sum(flux);
      if (lane == 0) s_part[(1 * nlev + p.nlay) * kWarps + warp] = s;
    }
    for (int l = p.nlay - 1; l >= 0; --l) {
      double b_t = 0.0;
      if (on) {
        const double tau = p.bg_od[(int64_t)l * p.stride + i] + fit[l] * sc;
        double e, f;
        lw_bb_layer(tau, p.tc, tab, e, f);
        b_t = p.planck_hl[(int64_t)l * p.stride + i];
        flux = flux * (1.0 - e) + b_bot * (e - f) + b_t * f;
      }
      const double s = warp_sum(flux);
      if (lane == 0) s_part[(1 * nlev + l) * kWarps + warp] = s;
      b_bot = b_t;
    }
  } else {
    const double minus_sec_sza = -1.0 / p.cos_sza;
    flux = on ? p.cos_sza * p.ssi[i] : 0.0;
    if (lane == 0) s_part[(0 * nlev + 0) * kWarps + warp] = 0.0;  // level 0 is cos_sza*sum(ssi): set in phase C
    for (int l = 0; l < p.nlay; ++l) {
      if (on) flux = flux * exp(minus_sec_sza * (p.bg_od[(int64_t)l * p.stride + i] + fit[l] * sc));
      const double s = warp_sum(flux);
      if (lane == 0) s_part[(0 * nlev + l + 1) * kWarps + warp] = s;
    }
    if (p.albedo > 0.0) {
      flux *= p.albedo;
      {
        const double s = warp_sum(flux);
        if (lane == 0) s_part[(1 * nlev + p.nlay) * kWarps + warp] = s;
      }
      for (int l = p.nlay - 1; l >= 0; --l) {
        if (on) flux = flux * exp(-2.0 * (p.bg_od[(int64_t)l * p.stride + i] + fit[l] * sc));
        const double s = warp_sum(flux);
        if (lane == 0) s_part[(1 * nlev + l) * kWarps + warp] = s;
      }
    } else {
      for (int l = lane; l < nlev; l += 32) s_part[(1 * nlev + l) * kWarps + warp] = 0.0;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * nlev; j += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kWarps; ++w) s += s_part[j * kWarps + w];
    p.part_flux[(size_t)item * 2 * nlev + j] = s;
  }
}

// ---- phase C: broadband fluxes of each interval, heating rate, RMS error.  CTA = interval.
__global__ void interval_error_kernel(const IntervalParams p) {
  extern __shared__ double s_c[];  // dn[nlev], up[nlev], term[nlay]
  const int k = blockIdx.x, nlev = p.nlay + 1;
  double* dn = s_c;
  double* up = s_c + nlev;
  double* term = up + nlev;
  const int64_t it0 = p.item_start[k], it1 = p.item_start[k + 1];
  const double* truth = p.truth ? p.truth + (size_t)k * (p.nlay + 4) : nullptr;
  double* s_seg = term + p.nlay;  // [kSeg][2*nlev]
  for (int idx = threadIdx.x; idx < kSeg * 2 * nlev; idx += blockDim.x) {
    const int seg = idx / (2 * nlev), j = idx - seg * 2 * nlev;
    int64_t a, b;
    seg_bounds(it0, it1, seg, a, b);
    s_seg[idx] = seg_sum(p.part_flux + j, (size_t)2 * nlev, a, b);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * nlev; j += blockDim.x) {
    double s = 0.0;
    for (int seg = 0; seg < kSeg; ++seg) s += s_seg[seg * 2 * nlev + j];
    if (p.sw && j == 0) s = p.cos_sza * truth[p.nlay + 2];  // flux_dn(0) = cos_sza*sum(ssi), radiative_transfer_sw.cpp:134
    s_c[j] = s;
    if (p.bb_flux) p.bb_flux[(size_t)k * 2 * nlev + j] = s;
  }
  if (!p.error) return;
  __syncthreads();
  const double hr_weight = 3600.0 * 24.0;
  for (int l = threadIdx.x; l < p.nlay; l += blockDim.x) {
    const double c = -((ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p.pressure_hl[l + 1] - p.pressure_hl[l]));
    // LW: heating_rate_single with both fluxes; SW: Vector() for the upwelling (calc_cost_function_sw.cpp:87)
    const double hr_fit = p.sw ? c * (dn[l + 1] - dn[l]) : c * (dn[l + 1] - dn[l] - up[l + 1] + up[l]);
    const double d = hr_fit - truth[l];
    term[l] = p.layer_weight[l] * (d * d);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int l = 0; l < p.nlay; ++l) s += term[l];
    const double dd = dn[p.nlay] - truth[p.nlay], du = up[0] - truth[p.nlay + 1];
    const double e = sqrt(hr_weight * hr_weight * s + p.flux_weight * (dd * dd + du * du));
    if (p.error_mode == 0) p.error[k] = e;
    else if (p.error_mode == 1) p.error[k] = 0.5 * e;
    else p.error[k] += 0.5 * e;
  }
}

static inline size_t finish_smem(int nlay) { return ((size_t)kSeg * (nlay * 4 + 4) + nlay) * sizeof(double); }

// Host side of a batch: item table, scratch carving, launches.
struct IntervalBatch {
  IntervalParams p;
  std::vector<int64_t> item_start;
};

static int setup_batch(ecckd_ctx* ctx, IntervalParams& p, int nint, const int64_t* h_i1, const int64_t* h_i2, int64_t n,
                       const char* who) {
  ECCKD_REQUIRE(nint > 0 && h_i1 && h_i2, ECCKD_PARAMETER_ERROR, "%s: no intervals", who);
  ECCKD_REQUIRE(finish_smem(p.nlay) <= 48 * 1024, ECCKD_PARAMETER_ERROR, "%s: more than 180 layers", who);
  std::vector<int64_t> host((size_t)3 * nint + 1);
  int64_t* i1 = host.data();
  int64_t* i2 = i1 + nint;
  int64_t* start = i2 + nint;
  int64_t items = 0;
  for (int k = 0; k < nint; ++k) {
    // find_g_points.cpp:297-312: indices outside the valid range or out of order are a PROCESSING_ERROR
    ECCKD_REQUIRE(h_i1[k] >= 0 && h_i2[k] < n && h_i2[k] >= h_i1[k], ECCKD_PROCESSING_ERROR,
                  "%s: interval %d = [%lld,%lld] outside 0-%lld or out of order", who, k, (long long)h_i1[k],
                  (long long)h_i2[k], (long long)(n - 1));
    i1[k] = h_i1[k];
    i2[k] = h_i2[k];
    start[k] = items;
    items += (h_i2[k] - h_i1[k] + 1 + kTile - 1) / kTile;
  }
  start[nint] = items;
  ECCKD_REQUIRE(items < 2147483647LL, ECCKD_PARAMETER_ERROR, "%s: too many tiles", who);
  p.nint = nint;
  p.nitems = (int)items;
  const int nlev = p.nlay + 1;
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_idx = up256(host.size() * sizeof(int64_t));
  const size_t b_fit = up256((size_t)items * p.nlay * 4 * sizeof(double));
  const size_t b_scal = up256((size_t)items * 4 * sizeof(double));
  const size_t b_flux = up256((size_t)items * 2 * nlev * sizeof(double));
  const size_t b_truth = up256((size_t)nint * (p.nlay + 4) * sizeof(double));
  const size_t b_fitout = up256((size_t)nint * p.nlay * sizeof(double));
  void* base = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 2, b_idx + b_fit + b_scal + b_flux + b_truth + b_fitout, &base));
  char* c = static_cast<char*>(base);
  ECCKD_TRY(upload_small(ctx, c, host.data(), host.size() * sizeof(int64_t)));
  p.i1 = reinterpret_cast<int64_t*>(c);
  p.i2 = p.i1 + nint;
  p.item_start = p.i2 + nint;
  c += b_idx;
  p.part_fit = reinterpret_cast<double*>(c); c += b_fit;
  p.part_scal = reinterpret_cast<double*>(c); c += b_scal;
  p.part_flux = reinterpret_cast<double*>(c); c += b_flux;
  p.truth = reinterpret_cast<double*>(c); c += b_truth;
  if (!p.fit) p.fit = reinterpret_cast<double*>(c);
  return ECCKD_OK;
}

static int launch_flux_and_error(ecckd_ctx* ctx, const IntervalParams& p) {
  // (Capping the resident CTAs per SM so that the upward sweep's re-reads hit in L2 was measured and is
  // slower: 2.65 ms unrestricted, 3.2 / 3.6 / 4.3 / 5.6 ms at 5 / 4 / 3 / 2 CTAs per SM.)
  const size_t smem = (size_t)2 * (p.nlay + 1) * kWarps * sizeof(double) + (p.sw ? 0 : TRANS_TAB_BYTES);
  if (p.sw) interval_flux_kernel<true><<<p.nitems, kTile, smem, ctx->stream>>>(p);
  else interval_flux_kernel<false><<<p.nitems, kTile, smem, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  interval_error_kernel<<<p.nint, 512, (size_t)(3 * p.nlay + 2 + kSeg * 2 * (p.nlay + 1)) * sizeof(double), ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// Layer emissivity and Planck-gradient factor of radiative_transfer_lw (radiative_transfer_lw.cpp:41-43) with the
// table exponential and a Newton reciprocal (as rt_lw_staged_kernel; no shuffles, so any thread may call it)
__device__ __forceinline__ void lw_layer_tab(double tau, const TransConst256& K, const unsigned char* tab, double& e,
                                             double& f) {
  e = trans256_one_minus_exp(tau, K, tab);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(tau));
  r = fma(r, fma(-tau, r, 1.0), r);
  r = fma(r, fma(-tau, r, 1.0), r);
  f = (e > 1.0e-5) ? (1.0 - (e * (1.0 / ECCKD_LW_DIFFUSIVITY)) * r) : (0.5 * e);
}

// ------------------------------------------------------------------ forward cost of optimize_lut
struct CkdCostParams {
  TransConst256 tc;
  int sw, nprof, nlay, ng, nband, nemis, nalbedo, up_in_bb, direct_only;
  const double* pressure_hl;       // [nprof][nlay+1]
  const double* planck_hl;         // LW [nprof][nlay+1][ng]
  const double* surf_emissivity;   // LW [nprof][nemis]
  const double* surf_planck;       // LW [nprof][ng]
  const double* cos_sza;           // SW [nprof]
  const double* ssi;               // SW [ng] (already scaled by tsi_scaling)
  const double* albedo;            // SW [nalbedo]
  const double* optical_depth;     // [nprof][nlay][ng]
  const double* flux_dn;           // [nprof][nlay+1][nband]
  const double* flux_up;           // [nprof][nlay+1][nband]
  const double* hr;                // [nprof][nlay][nband]
  const double* spectral_flux_dn_surf;  // [nprof][ng] or NULL
  const double* spectral_flux_up_toa;   // [nprof][ng] or NULL
  const double* spectral_boundary_weights;  // SW [ng] or NULL
  const double* layer_weight;      // [nprof][nlay] or NULL (then from pressure_weight_power)
  const double* rel_dn;            // [nprof][nlay+1][ng] or NULL
  const double* rel_up;
  const int32_t* band_mapping;     // [ng] or NULL
  double flux_weight, flux_profile_weight, broadband_weight, spectral_boundary_weight, negative_od_penalty,
      pressure_weight_power;
  double* cost_profile;            // [nprof]
  double* penalty_profile;         // [nprof]
  double* cost_band_profile;       // [nprof][nband] or NULL
};

// CTA = profile, thread = g point for the sweep; the (tiny) cost terms follow the reference's
// evaluation order with sequential sums.
__device__ __forceinline__ void ckd_cost_body(const CkdCostParams& p, const int ip) {
  extern __shared__ double sm[];
  const int nlay = p.nlay, nlev = nlay + 1, ng = p.ng, nband = p.nband;
  double* fdo = sm;                          // [nlev][ng]
  double* fuo = fdo + (size_t)nlev * ng;     // [nlev][ng]
  double* fd = fuo + (size_t)nlev * ng;      // [nlev][nband]
  double* fu = fd + (size_t)nlev * nband;    // [nlev][nband]
  double* hrf = fu + (size_t)nlev * nband;   // [nlay][nband]
  double* lw = hrf + (size_t)nlay * nband;   // [nlay]
  double* iw = lw + nlay;                    // [nlay]
  double* tmp = iw + nlay;                   // [max(nband, nlev, ng)]
  double* red = tmp + max(max(nband, nlev), ng);  // [8]
  const double* p_hl = p.pressure_hl + (size_t)ip * nlev;
  const double* od = p.optical_depth + (size_t)ip * nlay * ng;
  const double hr_weight = 3600.0 * 24.0;
  double* s_tab = red + 8 + 3 * nlev + nband + 2;  // 2^(j/256) table of the longwave layer terms
  if (!p.sw) trans_tab_init(s_tab);
  const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_tab);
  __syncthreads();

  // layer weights (solve_adept.cpp:133-145)
  if (threadIdx.x == 0) {
    if (p.layer_weight) {
      for (int l = 0; l < nlay; ++l) lw[l] = p.layer_weight[(size_t)ip * nlay + l];
    } else {
      double s = 0.0;
      for (int l = 0; l < nlay; ++l) {
        double v;
        if (p.pressure_weight_power == 0.5) v = sqrt(p_hl[l + 1]) - sqrt(p_hl[l]);
        else if (p.pressure_weight_power == 1.0) v = p_hl[l + 1] - p_hl[l];
        else v = pow(p_hl[l + 1], p.pressure_weight_power) - pow(p_hl[l], p.pressure_weight_power);
        lw[l] = v;
        s += v;
      }
      for (int l = 0; l < nlay; ++l) lw[l] /= s;
    }
    for (int l = 0; l + 1 < nlay; ++l) iw[l] = p.flux_profile_weight * 0.5 * (lw[l] + lw[l + 1]);
  }

  // forward fluxes per g point, with the negative optical depths penalised and clamped
  // (solve_adept.cpp:107-116)
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    double pen = 0.0;
    if (!p.sw) {
      const double* pl = p.planck_hl + (size_t)ip * nlev * ng + g;
      double f_dn = 0.0, b_top = pl[0];
      fdo[g] = 0.0;
      for (int l = 0; l < nlay; ++l) {
        double tau = od[(size_t)l * ng + g];
        if (tau < 0.0) { pen += tau * tau; tau = 0.0; }
        double e, f;
        lw_layer_tab(tau, p.tc, tab, e, f);
        const double b_bot = pl[(size_t)(l + 1) * ng];
        f_dn = f_dn * (1.0 - e) + b_top * (e - f) + b_bot * f;
        fdo[(size_t)(l + 1) * ng + g] = f_dn;
        b_top = b_bot;
      }
      const int ie = p.band_mapping ? p.band_mapping[g] : g;
      const double es = p.surf_emissivity[(size_t)ip * p.nemis + ie];
      double f_up = p.surf_planck[(size_t)ip * ng + g] * es + (1.0 - es) * f_dn;
      fuo[(size_t)nlay * ng + g] = f_up;
      double b_bot = b_top;
      for (int l = nlay - 1; l >= 0; --l) {
        const double tau = fmax(od[(size_t)l * ng + g], 0.0);
        double e, f;
        lw_layer_tab(tau, p.tc, tab, e, f);
        const double b_t = pl[(size_t)l * ng];
        f_up = f_up * (1.0 - e) + b_bot * (e - f) + b_t * f;
        fuo[(size_t)l * ng + g] = f_up;
        b_bot = b_t;
      }
    } else {
      const double mu0 = p.cos_sza[ip];
      const double minus_sec_sza = -1.0 / mu0;
      double f = mu0 * p.ssi[g];
      fdo[g] = f;
      for (int l = 0; l < nlay; ++l) {
        double tau = od[(size_t)l * ng + g];
        if (tau < 0.0) { pen += tau * tau; tau = 0.0; }
        f = f * exp(minus_sec_sza * tau);
        fdo[(size_t)(l + 1) * ng + g] = f;
      }
      if (p.direct_only) {
        for (int l = 0; l < nlev; ++l) fuo[(size_t)l * ng + g] = 0.0;
      } else {
        f = f * p.albedo[p.band_mapping[g]];
        fuo[(size_t)nlay * ng + g] = f;
        for (int l = nlay - 1; l >= 0; --l) {
          f = f * exp(-2.0 * fmax(od[(size_t)l * ng + g], 0.0));
          fuo[(size_t)l * ng + g] = f;
        }
      }
    }
    tmp[g] = pen;
    if (p.rel_dn) {
      const double* rd = p.rel_dn + (size_t)ip * nlev * ng + g;
      const double* ru = p.rel_up + (size_t)ip * nlev * ng + g;
      for (int l = 0; l < nlev; ++l) {
        fdo[(size_t)l * ng + g] -= rd[(size_t)l * ng];
        fuo[(size_t)l * ng + g] -= ru[(size_t)l * ng];
      }
    }
  }
  // penalty: per-g partial sums added in g order
  __syncthreads();
  if (threadIdx.x == 0) {
    double pen = 0.0;
    for (int g = 0; g < ng; ++g) pen += tmp[g];
    p.penalty_profile[ip] = p.negative_od_penalty * pen;
  }
  __syncthreads();

  // g points -> bands.  The g points of a band are usually contiguous (iband_per_g is non-decreasing):
  // then only that range is walked; either way the sum runs over the band's g points in ascending order.
  int* s_rng = reinterpret_cast<int*>(red + 8 + 3 * nlev);  // [nband][2]: first g, last g (or -1: not contiguous)
  if (p.band_mapping) {
    for (int ib = threadIdx.x; ib < nband; ib += blockDim.x) {
      int first = -1, last = -1, cnt = 0;
      for (int g = 0; g < ng; ++g) {
        if (p.band_mapping[g] == ib) { if (first < 0) first = g; last = g; ++cnt; }
      }
      const bool contig = cnt > 0 && last - first + 1 == cnt;
      s_rng[2 * ib] = contig ? first : (cnt == 0 ? 0 : -1);
      s_rng[2 * ib + 1] = contig ? last : (cnt == 0 ? -1 : 0);
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < nlev * nband; j += blockDim.x) {
    const int l = j / nband, ib = j - l * nband;
    if (!p.band_mapping) {
      fd[j] = fdo[(size_t)l * ng + ib];
      fu[j] = fuo[(size_t)l * ng + ib];
    } else {
      double sd = 0.0, su = 0.0;
      const int g0 = s_rng[2 * ib], g1 = s_rng[2 * ib + 1];
      if (g0 >= 0) {
        for (int g = g0; g <= g1; ++g) { sd += fdo[(size_t)l * ng + g]; su += fuo[(size_t)l * ng + g]; }
      } else {
        for (int g = 0; g < ng; ++g) {
          if (p.band_mapping[g] == ib) { sd += fdo[(size_t)l * ng + g]; su += fuo[(size_t)l * ng + g]; }
        }
      }
      fd[j] = sd;
      fu[j] = su;
    }
  }
  __syncthreads();
  // heating rates (heating_rate.h:38-48; the SW passes an empty upwelling, calc_cost_function_sw.cpp:195)
  for (int j = threadIdx.x; j < nlay * nband; j += blockDim.x) {
    const int l = j / nband, ib = j - l * nband;
    const double conv = -(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p_hl[l + 1] - p_hl[l]);
    const size_t a = (size_t)l * nband + ib;
    hrf[j] = p.sw ? conv * (fd[a + nband] - fd[a]) : conv * (fd[a + nband] - fd[a] - fu[a + nband] + fu[a]);
  }
  __syncthreads();
  const double* t_dn = p.flux_dn + (size_t)ip * nlev * nband;
  const double* t_up = p.flux_up + (size_t)ip * nlev * nband;
  const double* t_hr = p.hr + (size_t)ip * nlay * nband;
  // per-band terms
  for (int ib = threadIdx.x; ib < nband; ib += blockDim.x) {
    double s = 0.0;
    for (int l = 0; l < nlay; ++l) {
      const double d = hrf[(size_t)l * nband + ib] - t_hr[(size_t)l * nband + ib];
      s += lw[l] * d * d;
    }
    const double dd = fd[(size_t)nlay * nband + ib] - t_dn[(size_t)nlay * nband + ib];
    const double du = fu[ib] - t_up[ib];
    double local = hr_weight * hr_weight * s + p.flux_weight * (dd * dd + (p.sw ? 20.0 : 1.0) * du * du);
    if (p.flux_profile_weight > 0.0) {
      double t = 0.0;
      for (int l = 1; l < nlay; ++l) {
        const double a = fd[(size_t)l * nband + ib] - t_dn[(size_t)l * nband + ib];
        const double b = fu[(size_t)l * nband + ib] - t_up[(size_t)l * nband + ib];
        t += iw[l - 1] * (a * a + b * b);
      }
      local += t;
    }
    tmp[ib] = local;
    if (p.cost_band_profile) p.cost_band_profile[(size_t)ip * nband + ib] = local;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
    for (int ib = 0; ib < nband; ++ib) c += tmp[ib];
    red[0] = c;
  }
  __syncthreads();
  // broadband terms: errors summed over bands, per layer / level
  for (int l = threadIdx.x; l < nlev; l += blockDim.x) {
    double e = 0.0, a = 0.0, b = 0.0;
    for (int ib = 0; ib < nband; ++ib) {
      if (l < nlay) e += hrf[(size_t)l * nband + ib] - t_hr[(size_t)l * nband + ib];
      a += fd[(size_t)l * nband + ib] - t_dn[(size_t)l * nband + ib];
      b += fu[(size_t)l * nband + ib] - t_up[(size_t)l * nband + ib];
    }
    red[8 + 3 * l + 0] = e;
    red[8 + 3 * l + 1] = a;
    red[8 + 3 * l + 2] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cost = red[0];
    const double* bb = red + 8;
    if (!p.sw || p.broadband_weight > 0.0) {
      double s = 0.0;
      for (int l = 0; l < nlay; ++l) s += lw[l] * (bb[3 * l] * bb[3 * l]);
      const double ed = bb[3 * nlay + 1], eu = bb[2];
      if (!p.sw) {
        cost = (cost * (1.0 - p.broadband_weight)) / nband + p.broadband_weight * hr_weight * hr_weight * s +
               p.broadband_weight * p.flux_weight * (ed * ed + eu * eu);
      } else {
        cost = (cost * (1.0 - p.broadband_weight)) / nband + p.broadband_weight * hr_weight * hr_weight * s;
        cost += p.broadband_weight * p.flux_weight * (ed * ed);
        if (p.up_in_bb) cost += p.broadband_weight * p.flux_weight * (eu * eu);
      }
      if (p.flux_profile_weight > 0.0) {
        if (!p.sw) {
          double t = 0.0;
          for (int l = 1; l < nlay; ++l) t += iw[l - 1] * (bb[3 * l + 1] * bb[3 * l + 1] + bb[3 * l + 2] * bb[3 * l + 2]);
          cost += p.broadband_weight * t;
        } else {
          double td = 0.0, tu = 0.0;
          for (int l = 1; l < nlay; ++l) {
            td += iw[l - 1] * (bb[3 * l + 1] * bb[3 * l + 1]);
            tu += iw[l - 1] * (bb[3 * l + 2] * bb[3 * l + 2]);
          }
          cost += p.broadband_weight * td;
          if (p.up_in_bb) cost += p.broadband_weight * tu;
        }
      }
    }
    // spectral boundary terms on the g-point fluxes
    if (!p.sw) {
      if (p.spectral_boundary_weight > 0.0 && p.spectral_flux_dn_surf && p.spectral_flux_up_toa) {
        double s = 0.0;
        for (int g = 0; g < ng; ++g) {
          const double a = fdo[(size_t)nlay * ng + g] - p.spectral_flux_dn_surf[(size_t)ip * ng + g];
          const double b = fuo[g] - p.spectral_flux_up_toa[(size_t)ip * ng + g];
          s += a * a + b * b;
        }
        cost += p.spectral_boundary_weight * s;
      }
    } else if (p.spectral_boundary_weights && p.spectral_flux_dn_surf) {
      double s = 0.0;
      for (int g = 0; g < ng; ++g) {
        const double a = fdo[(size_t)nlay * ng + g] - p.spectral_flux_dn_surf[(size_t)ip * ng + g];
        s += p.spectral_boundary_weights[g] * a * a;
      }
      cost += s;
    }
    p.cost_profile[ip] = cost;
  }
}

__global__ void ckd_cost_kernel(const CkdCostParams p) { ckd_cost_body(p, (int)blockIdx.x); }

// Several scenes (LblFluxes objects of the training set, solve_adept.cpp:96-199) in one launch:
// CTA = (scene, profile); scene_start[s] = first CTA of scene s.  The scene's parameters are copied
// to shared memory once per CTA.
__global__ void ckd_cost_scenes_kernel(const CkdCostParams* __restrict__ scenes, const int* __restrict__ scene_start,
                                       int nscene) {
  __shared__ CkdCostParams sp;
  __shared__ int s_first;
  if (threadIdx.x == 0) {
    int lo = 0, hi = nscene;  // last scene with scene_start <= blockIdx.x
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (scene_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    s_first = scene_start[lo];
    sp = scenes[lo];
  }
  __syncthreads();
  ckd_cost_body(sp, (int)blockIdx.x - s_first);
}

// per scene: penalties, then profile costs, in profile order; per-band costs of the scene
__global__ void ckd_cost_scene_sums_kernel(const CkdCostParams* __restrict__ scenes, int nband,
                                           double* __restrict__ cost_scene, double* __restrict__ penalty_scene,
                                           double* __restrict__ band_scene) {
  const CkdCostParams& p = scenes[blockIdx.x];
  if (threadIdx.x == 0) {
    double pen = 0.0, c = 0.0;
    for (int i = 0; i < p.nprof; ++i) pen += p.penalty_profile[i];
    for (int i = 0; i < p.nprof; ++i) c += p.cost_profile[i];
    penalty_scene[blockIdx.x] = pen;
    cost_scene[blockIdx.x] = pen + c;
  }
  if (band_scene && p.cost_band_profile) {
    for (int ib = threadIdx.x; ib < nband; ib += blockDim.x) {
      double t = 0.0;
      for (int i = 0; i < p.nprof; ++i) t += p.cost_band_profile[(size_t)i * nband + ib];
      band_scene[(size_t)blockIdx.x * nband + ib] = t;
    }
  }
}

// totals over the scenes in scene order (`cost +=` of solve_adept.cpp:157,183)
__global__ void ckd_cost_scene_total_kernel(int nscene, int nband, const double* __restrict__ cost_scene,
                                            const double* __restrict__ band_scene, double* __restrict__ cost_total,
                                            double* __restrict__ cost_fn_per_band) {
  if (threadIdx.x == 0 && cost_total) {
    double c = 0.0;
    for (int s = 0; s < nscene; ++s) c += cost_scene[s];
    *cost_total = c;
  }
  if (cost_fn_per_band && band_scene) {
    for (int ib = threadIdx.x; ib < nband; ib += blockDim.x) {
      double t = cost_fn_per_band[ib];
      for (int s = 0; s < nscene; ++s) t += band_scene[(size_t)s * nband + ib];
      cost_fn_per_band[ib] = t;
    }
  }
}

// cost = sum over profiles in profile order (solve_adept.cpp:96-199 `cost +=`), penalties first per scene
__global__ void ckd_cost_total_kernel(int nprof, int nband, const double* __restrict__ cost_profile,
                                      const double* __restrict__ penalty_profile,
                                      const double* __restrict__ cost_band_profile, double* __restrict__ cost_total,
                                      double* __restrict__ penalty_total, double* __restrict__ cost_fn_per_band) {
  if (threadIdx.x == 0) {
    double pen = 0.0, c = 0.0;
    for (int i = 0; i < nprof; ++i) pen += penalty_profile[i];
    for (int i = 0; i < nprof; ++i) c += cost_profile[i];
    if (penalty_total) *penalty_total = pen;
    if (cost_total) *cost_total = pen + c;
  }
  if (cost_fn_per_band && cost_band_profile) {
    for (int ib = threadIdx.x; ib < nband; ib += blockDim.x) {
      double s = cost_fn_per_band[ib];
      for (int i = 0; i < nprof; ++i) s += cost_band_profile[(size_t)i * nband + ib];
      cost_fn_per_band[ib] = s;
    }
  }
}

// checks the dimensions of a scene and returns the dynamic shared memory its CTAs need
// ---------------------------------------------------------------------------------------------
// Reverse mode of the longwave forward cost (first part): the gradient of
// calc_cost_function_ckd_lw (calc_cost_function_lw.cpp:115-232) plus the negative-optical-depth penalty
// (solve_adept.cpp:107-116) with respect to the optical depth of every (profile, layer, g point) - what the
// Adept tape of solve_adept.cpp:91, 201-203 delivers up to the look-up-table interpolation.  Written by hand:
//   forward sweep (as ckd_cost_body) -> band sums -> heating rates -> residuals;
//   dJ/d(band flux) from the quadratic terms (spectral and broadband parts, boundary and interior
//   fluxes, heating rates through their difference stencil), dJ/d(g-point flux) = its band's + the
//   spectral boundary term;
//   adjoint of the upward recurrence walked downwards (level 0 -> surface), the surface reflection
//   (1 - emissivity), adjoint of the downward recurrence walked upwards; per layer the adjoints of the
//   transmittance and of the two sources give dJ/de and dJ/df, and with de/dtau = D (1 - e),
//   df/dtau = (e / tau - de/dtau) / (D tau)  [e > 1e-5]  or  de/dtau / 2, the derivative in tau.
// Clamped (negative) optical depths only carry the penalty's derivative.  CTA = profile, thread = g point.
__global__ void ckd_cost_grad_kernel(const CkdCostParams p, double* __restrict__ grad) {
  extern __shared__ double sm[];
  const int ip = blockIdx.x;
  const int nlay = p.nlay, nlev = nlay + 1, ng = p.ng, nband = p.nband;
  double* fdo = sm;                          // [nlev][ng] forward flux down (relative-to flux subtracted)
  double* fuo = fdo + (size_t)nlev * ng;     // [nlev][ng] ... up; later dJ/de partial sums
  double* gd = fuo + (size_t)nlev * ng;      // [nlev][nband] band flux down -> dJ/d(band flux down)
  double* gu = gd + (size_t)nlev * nband;    // [nlev][nband]
  double* hrf = gu + (size_t)nlev * nband;   // [nlay][nband] heating-rate residual -> dJ/d(heating rate)
  double* lw = hrf + (size_t)nlay * nband;   // [nlay]
  double* iw = lw + nlay;                    // [nlay]
  double* sdn = iw + nlay;                   // [nlev] band sums of the residuals
  double* sup = sdn + nlev;                  // [nlev]
  double* shr = sup + nlev;                  // [nlay]
  double* s_tab = shr + nlay;                // 2^(j/256) table
  const double* p_hl = p.pressure_hl + (size_t)ip * nlev;
  const double* od = p.optical_depth + (size_t)ip * nlay * ng;
  double* g_out = grad + (size_t)ip * nlay * ng;
  const double hr_weight = 3600.0 * 24.0;
  const double D = ECCKD_LW_DIFFUSIVITY;
  trans_tab_init(s_tab);
  const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_tab);
  if (threadIdx.x == 0) {  // layer weights (solve_adept.cpp:133-145), as ckd_cost_body
    if (p.layer_weight) {
      for (int l = 0; l < nlay; ++l) lw[l] = p.layer_weight[(size_t)ip * nlay + l];
    } else {
      double s = 0.0;
      for (int l = 0; l < nlay; ++l) {
        double v;
        if (p.pressure_weight_power == 0.5) v = sqrt(p_hl[l + 1]) - sqrt(p_hl[l]);
        else if (p.pressure_weight_power == 1.0) v = p_hl[l + 1] - p_hl[l];
        else v = pow(p_hl[l + 1], p.pressure_weight_power) - pow(p_hl[l], p.pressure_weight_power);
        lw[l] = v;
        s += v;
      }
      for (int l = 0; l < nlay; ++l) lw[l] /= s;
    }
    for (int l = 0; l + 1 < nlay; ++l) iw[l] = p.flux_profile_weight * 0.5 * (lw[l] + lw[l + 1]);
  }
  __syncthreads();
  // ---- forward fluxes per g point
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    if (p.sw) {  // radiative_transfer_sw.cpp:24-77 (direct beam, Lambertian return without Rayleigh), as ckd_cost_body
      const double mu0 = p.cos_sza[ip];
      const double minus_sec_sza = -1.0 / mu0;
      double f = mu0 * p.ssi[g];
      fdo[g] = f;
      for (int l = 0; l < nlay; ++l) {
        f = f * exp(minus_sec_sza * fmax(od[(size_t)l * ng + g], 0.0));
        fdo[(size_t)(l + 1) * ng + g] = f;
      }
      if (p.direct_only) {
        for (int l = 0; l < nlev; ++l) fuo[(size_t)l * ng + g] = 0.0;
      } else {
        f = f * p.albedo[p.band_mapping[g]];
        fuo[(size_t)nlay * ng + g] = f;
        for (int l = nlay - 1; l >= 0; --l) {
          f = f * exp(-2.0 * fmax(od[(size_t)l * ng + g], 0.0));
          fuo[(size_t)l * ng + g] = f;
        }
      }
    } else {
    const double* pl = p.planck_hl + (size_t)ip * nlev * ng + g;
    double f_dn = 0.0, b_top = pl[0];
    fdo[g] = 0.0;
    for (int l = 0; l < nlay; ++l) {
      const double tau = fmax(od[(size_t)l * ng + g], 0.0);
      double e, f;
      lw_layer_tab(tau, p.tc, tab, e, f);
      const double b_bot = pl[(size_t)(l + 1) * ng];
      f_dn = f_dn * (1.0 - e) + b_top * (e - f) + b_bot * f;
      fdo[(size_t)(l + 1) * ng + g] = f_dn;
      b_top = b_bot;
    }
    const int ie = p.band_mapping ? p.band_mapping[g] : g;
    const double es = p.surf_emissivity[(size_t)ip * p.nemis + ie];
    double f_up = p.surf_planck[(size_t)ip * ng + g] * es + (1.0 - es) * f_dn;
    fuo[(size_t)nlay * ng + g] = f_up;
    double b_bot = b_top;
    for (int l = nlay - 1; l >= 0; --l) {
      const double tau = fmax(od[(size_t)l * ng + g], 0.0);
      double e, f;
      lw_layer_tab(tau, p.tc, tab, e, f);
      const double b_t = pl[(size_t)l * ng];
      f_up = f_up * (1.0 - e) + b_bot * (e - f) + b_t * f;
      fuo[(size_t)l * ng + g] = f_up;
      b_bot = b_t;
    }
    }
    if (p.rel_dn) {
      const double* rd = p.rel_dn + (size_t)ip * nlev * ng + g;
      const double* ru = p.rel_up + (size_t)ip * nlev * ng + g;
      for (int l = 0; l < nlev; ++l) {
        fdo[(size_t)l * ng + g] -= rd[(size_t)l * ng];
        fuo[(size_t)l * ng + g] -= ru[(size_t)l * ng];
      }
    }
  }
  __syncthreads();
  // ---- band flux residuals, heating-rate residuals
  const double* t_dn = p.flux_dn + (size_t)ip * nlev * nband;
  const double* t_up = p.flux_up + (size_t)ip * nlev * nband;
  const double* t_hr = p.hr + (size_t)ip * nlay * nband;
  for (int j = threadIdx.x; j < nlev * nband; j += blockDim.x) {
    const int l = j / nband, ib = j - l * nband;
    double sd = 0.0, su = 0.0;
    if (!p.band_mapping) {
      sd = fdo[(size_t)l * ng + ib];
      su = fuo[(size_t)l * ng + ib];
    } else {
      for (int g = 0; g < ng; ++g)
        if (p.band_mapping[g] == ib) { sd += fdo[(size_t)l * ng + g]; su += fuo[(size_t)l * ng + g]; }
    }
    gd[j] = sd;
    gu[j] = su;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nlay * nband; j += blockDim.x) {
    const int l = j / nband, ib = j - l * nband;
    const double conv = -(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p_hl[l + 1] - p_hl[l]);
    const size_t a = (size_t)l * nband + ib;
    hrf[j] = (p.sw ? conv * (gd[a + nband] - gd[a]) : conv * (gd[a + nband] - gd[a] - gu[a + nband] + gu[a])) - t_hr[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nlev * nband; j += blockDim.x) {  // flux residuals in place
    gd[j] -= t_dn[j];
    gu[j] -= t_up[j];
  }
  __syncthreads();
  for (int l = threadIdx.x; l < nlev; l += blockDim.x) {  // band sums of the residuals (band order)
    double a = 0.0, b = 0.0, c = 0.0;
    for (int ib = 0; ib < nband; ++ib) {
      a += gd[(size_t)l * nband + ib];
      b += gu[(size_t)l * nband + ib];
      if (l < nlay) c += hrf[(size_t)l * nband + ib];
    }
    sdn[l] = a;
    sup[l] = b;
    if (l < nlay) shr[l] = c;
  }
  __syncthreads();
  // ---- dJ/d(heating rate), dJ/d(band flux) of the quadratic terms (calc_cost_function_lw.cpp:194-222)
  // (shortwave, calc_cost_function_sw.cpp:209-269: the spectral part is only divided by nband when there is a
  // broadband part; the TOA upwelling error counts 20-fold in the spectral part; the upwelling enters the broadband
  // part only when every albedo is positive)
  const double bw = p.broadband_weight;
  const double c1 = (!p.sw || bw > 0.0) ? (1.0 - bw) / nband : 1.0;
  const double up_bb = (!p.sw || p.up_in_bb) ? bw : 0.0;
  for (int j = threadIdx.x; j < nlay * nband; j += blockDim.x) {
    const int l = j / nband;
    hrf[j] = 2.0 * hr_weight * hr_weight * lw[l] * (c1 * hrf[j] + bw * shr[l]);
  }
  for (int j = threadIdx.x; j < nlev * nband; j += blockDim.x) {
    const int l = j / nband;
    double wd = 0.0, wu = 0.0, wus = 1.0;  // weights of this level's squared flux errors
    if (l == nlay) wd = p.flux_weight;
    if (l == 0) { wu = p.flux_weight; wus = p.sw ? 20.0 : 1.0; }
    if (l >= 1 && l <= nlay - 1 && p.flux_profile_weight > 0.0) wd = wu = iw[l - 1];
    gd[j] = 2.0 * wd * (c1 * gd[j] + bw * sdn[l]);
    gu[j] = 2.0 * wu * (wus * c1 * gu[j] + up_bb * sup[l]);
  }
  __syncthreads();
  // heating rate -> fluxes (heating_rate.h:38-48): hr(l) = conv (dn(l+1) - dn(l) - up(l+1) + up(l))
  for (int j = threadIdx.x; j < nlev * nband; j += blockDim.x) {
    const int l = j / nband, ib = j - l * nband;
    double a = 0.0;  // dJ/d dn(l) from the heating rates of the layers above and below
    if (l >= 1) a += (-(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p_hl[l] - p_hl[l - 1])) * hrf[(size_t)(l - 1) * nband + ib];
    if (l < nlay) a -= (-(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p_hl[l + 1] - p_hl[l])) * hrf[(size_t)l * nband + ib];
    gd[j] += a;
    if (!p.sw) gu[j] -= a;  // heating_rate.h:38-42: the shortwave passes an empty upwelling
  }
  __syncthreads();
  // ---- reverse sweeps per g point
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    if (p.sw) {
      const int ib = p.band_mapping ? p.band_mapping[g] : g;
      const double* rd = p.rel_dn ? p.rel_dn + (size_t)ip * nlev * ng + g : nullptr;
      const double* ru = p.rel_up ? p.rel_up + (size_t)ip * nlev * ng + g : nullptr;
      const double inv_mu0 = 1.0 / p.cos_sza[ip];
      double a = 0.0;
      // up(l) = up(l+1) exp(-2 tau(l)): d up(l) / d tau(l) = -2 up(l); walked from the top down
      if (!p.direct_only) {
        a = gu[ib];
        for (int l = 0; l < nlay; ++l) {
          const double up_l = fuo[(size_t)l * ng + g] + (ru ? ru[(size_t)l * ng] : 0.0);  // raw flux_up(l)
          g_out[(size_t)l * ng + g] = -2.0 * a * up_l;
          a = a * exp(-2.0 * fmax(od[(size_t)l * ng + g], 0.0)) + gu[(size_t)(l + 1) * nband + ib];
        }
      } else {
        for (int l = 0; l < nlay; ++l) g_out[(size_t)l * ng + g] = 0.0;
      }
      // up(nlay) = albedo dn(nlay); dn(l+1) = dn(l) exp(-tau(l) / mu0): d dn(l+1) / d tau(l) = -dn(l+1) / mu0
      double d = gd[(size_t)nlay * nband + ib] + (p.direct_only ? 0.0 : a * p.albedo[p.band_mapping[g]]);
      if (p.spectral_boundary_weights && p.spectral_flux_dn_surf)
        d += 2.0 * p.spectral_boundary_weights[g] * (fdo[(size_t)nlay * ng + g] - p.spectral_flux_dn_surf[(size_t)ip * ng + g]);
      for (int l = nlay - 1; l >= 0; --l) {
        const double tau_raw = od[(size_t)l * ng + g];
        const double dn_next = fdo[(size_t)(l + 1) * ng + g] + (rd ? rd[(size_t)(l + 1) * ng] : 0.0);  // raw flux_dn(l+1)
        double gt = g_out[(size_t)l * ng + g] - d * dn_next * inv_mu0;
        if (tau_raw < 0.0) gt = 2.0 * p.negative_od_penalty * tau_raw;
        g_out[(size_t)l * ng + g] = gt;
        d = d * exp(-inv_mu0 * fmax(tau_raw, 0.0)) + gd[(size_t)l * nband + ib];
      }
      continue;
    }
    const double* pl = p.planck_hl + (size_t)ip * nlev * ng + g;
    const int ib = p.band_mapping ? p.band_mapping[g] : g;
    const double* rd = p.rel_dn ? p.rel_dn + (size_t)ip * nlev * ng + g : nullptr;
    const double* ru = p.rel_up ? p.rel_up + (size_t)ip * nlev * ng + g : nullptr;
    const bool sb = p.spectral_boundary_weight > 0.0 && p.spectral_flux_dn_surf && p.spectral_flux_up_toa;
    // adjoint of the upward recurrence: level 0 down to the surface
    double a = gu[ib];
    if (sb) a += 2.0 * p.spectral_boundary_weight * (fuo[g] - p.spectral_flux_up_toa[(size_t)ip * ng + g]);
    for (int l = 0; l < nlay; ++l) {
      const double tau = fmax(od[(size_t)l * ng + g], 0.0);
      double e, f;
      lw_layer_tab(tau, p.tc, tab, e, f);
      const double up_next = fuo[(size_t)(l + 1) * ng + g] + (ru ? ru[(size_t)(l + 1) * ng] : 0.0);  // raw flux_up(l+1)
      const double b_t = pl[(size_t)l * ng], b_b = pl[(size_t)(l + 1) * ng];
      // up(l) = up(l+1) (1 - e) + b_b (e - f) + b_t f
      fuo[(size_t)l * ng + g] = a * (b_b - up_next);  // dJ/de so far (fuo(l) is no longer needed)
      g_out[(size_t)l * ng + g] = a * (b_t - b_b);     // dJ/df so far
      a = a * (1.0 - e) + gu[(size_t)(l + 1) * nband + ib];
    }
    // surface: up(nlay) = surf_planck emissivity + (1 - emissivity) dn(nlay)
    const int ie = p.band_mapping ? p.band_mapping[g] : g;
    const double es = p.surf_emissivity[(size_t)ip * p.nemis + ie];
    double d = gd[(size_t)nlay * nband + ib] + a * (1.0 - es);
    if (sb) d += 2.0 * p.spectral_boundary_weight * (fdo[(size_t)nlay * ng + g] - p.spectral_flux_dn_surf[(size_t)ip * ng + g]);
    // adjoint of the downward recurrence: surface up to level 0
    for (int l = nlay - 1; l >= 0; --l) {
      const double tau_raw = od[(size_t)l * ng + g];
      const double tau = fmax(tau_raw, 0.0);
      double e, f;
      lw_layer_tab(tau, p.tc, tab, e, f);
      const double dn_l = fdo[(size_t)l * ng + g] + (rd ? rd[(size_t)l * ng] : 0.0);  // raw flux_dn(l)
      const double b_t = pl[(size_t)l * ng], b_b = pl[(size_t)(l + 1) * ng];
      // dn(l+1) = dn(l) (1 - e) + b_t (e - f) + b_b f
      const double de = fuo[(size_t)l * ng + g] + d * (b_t - dn_l);
      const double df = g_out[(size_t)l * ng + g] + d * (b_b - b_t);
      const double dedt = D * (1.0 - e);
      double dfdt;
      if (e > 1.0e-5) {
        dfdt = (e / tau - dedt) / (D * tau);  // f = 1 - e / (D tau)
      } else {
        dfdt = 0.5 * dedt;
      }
      double gt = de * dedt + df * dfdt;
      if (tau_raw < 0.0) gt = 2.0 * p.negative_od_penalty * tau_raw;  // clamped: only the penalty depends on it
      g_out[(size_t)l * ng + g] = gt;
      d = d * (1.0 - e) + gd[(size_t)l * nband + ib];
    }
  }
}

static int ckd_cost_smem(ecckd_ctx* ctx, CkdCostParams& p, const char* who, size_t* smem_out) {
  const int nlev = p.nlay + 1;
  trans_const256_init(&p.tc, ECCKD_LW_DIFFUSIVITY);
  ECCKD_REQUIRE(p.nprof > 0 && p.nlay > 0 && p.ng > 0 && p.nband > 0, ECCKD_PARAMETER_ERROR, "%s: bad dimensions", who);
  ECCKD_REQUIRE(p.band_mapping || p.nband == p.ng, ECCKD_PARAMETER_ERROR,
                "%s: without band_mapping the truth arrays are per g point (nband must equal ng)", who);
  const size_t ndbl = (size_t)2 * nlev * p.ng + (size_t)2 * nlev * p.nband + (size_t)p.nlay * p.nband + 2 * p.nlay +
                      std::max(std::max(p.nband, nlev), p.ng) + 8 + 3 * nlev + p.nband + 2 + 256;
  const size_t smem = ndbl * sizeof(double);
  ECCKD_REQUIRE(smem + sizeof(CkdCostParams) + 64 <= ctx->smem_optin, ECCKD_PARAMETER_ERROR,
                "%s: %d g points x %d levels need %zu B of shared memory per profile (limit %zu)", who, p.ng, nlev, smem,
                ctx->smem_optin);
  *smem_out = smem;
  return ECCKD_OK;
}

static int launch_ckd_cost(ecckd_ctx* ctx, CkdCostParams& p, double* d_cost_profile, double* d_cost_total,
                           double* d_penalty_total, double* d_cost_fn_per_band, const char* who) {
  size_t smem = 0;
  ECCKD_TRY(ckd_cost_smem(ctx, p, who, &smem));
  ECCKD_CUDA_OK(cudaFuncSetAttribute(ckd_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* base = nullptr;
  const size_t b_prof = ((size_t)p.nprof * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_band = d_cost_fn_per_band ? (size_t)p.nprof * p.nband * sizeof(double) : 0;
  ECCKD_TRY(ws_reserve(ctx, 2, 2 * b_prof + b_band, &base));
  char* c = static_cast<char*>(base);
  p.penalty_profile = reinterpret_cast<double*>(c);
  double* own_cost = reinterpret_cast<double*>(c + b_prof);
  p.cost_profile = d_cost_profile ? d_cost_profile : own_cost;
  p.cost_band_profile = d_cost_fn_per_band ? reinterpret_cast<double*>(c + 2 * b_prof) : nullptr;
  const int threads = std::min(1024, std::max(128, (p.ng + 31) / 32 * 32));
  ckd_cost_kernel<<<p.nprof, threads, smem, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  ckd_cost_total_kernel<<<1, 64, 0, ctx->stream>>>(p.nprof, p.nband, p.cost_profile, p.penalty_profile,
                                                   p.cost_band_profile, d_cost_total, d_penalty_total,
                                                   d_cost_fn_per_band);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

}  // namespace ecckd

using namespace ecckd;

// ------------------------------------------------------------------ C ABI
ECCKD_API int ecckd_interval_bounds_to_indices(int64_t npoints, int nint, const double* bound1, const double* bound2,
                                               int64_t* i1, int64_t* i2) {
  ECCKD_REQUIRE(npoints > 0 && nint >= 0 && bound1 && bound2 && i1 && i2, ECCKD_PARAMETER_ERROR,
                "ecckd_interval_bounds_to_indices: bad argument");
  for (int k = 0; k < nint; ++k) {
    // find_g_points.cpp:284-315
    int64_t ib1 = (int64_t)std::ceil(bound1[k] * (double)(npoints - 1));
    int64_t ib2 = (int64_t)std::floor(bound2[k] * (double)(npoints - 1));
    ECCKD_REQUIRE(!(ib1 < 0 || ib2 >= npoints), ECCKD_PROCESSING_ERROR,
                  "requested bounds %g-%g corresponding to indices %lld-%lld outside valid range 0-%lld", bound1[k],
                  bound2[k], (long long)ib1, (long long)ib2, (long long)(npoints - 1));
    ECCKD_REQUIRE(!(bound2[k] < bound1[k]), ECCKD_PROCESSING_ERROR, "requested bounds out of order: %g-%g", bound1[k],
                  bound2[k]);
    ECCKD_REQUIRE(!(ib2 + 1 < ib1), ECCKD_PROCESSING_ERROR, "requested indices out of order: %lld-%lld", (long long)ib1,
                  (long long)ib2);
    if (ib2 < ib1) ib2 = ib1;
    i1[k] = ib1;
    i2[k] = ib2;
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_interval_error_batch_lw_dev(ecckd_ctx* ctx, int averaging_method, int nlay, int64_t n, int64_t stride,
                                                const double* d_pressure_hl, const double* d_planck_hl,
                                                const double* d_surf_emissivity, const double* d_surf_planck,
                                                const double* d_bg_optical_depth, const double* d_metric,
                                                const double* d_flux_dn_surf, const double* d_flux_up_toa,
                                                const double* d_hr, double flux_weight, const double* d_layer_weight,
                                                int nint, const int64_t* h_i1, const int64_t* h_i2, double* d_error,
                                                double* d_optical_depth_fit) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_interval_error_batch_lw";
  ECCKD_REQUIRE(ctx && d_pressure_hl && d_planck_hl && d_surf_emissivity && d_surf_planck && d_bg_optical_depth &&
                    d_metric && d_flux_dn_surf && d_flux_up_toa && d_hr && d_layer_weight && d_error,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE(averaging_method == ECCKD_AVG_LINEAR || averaging_method == ECCKD_AVG_TRANSMISSION ||
                    averaging_method == ECCKD_AVG_TRANSMISSION_2 || averaging_method == ECCKD_AVG_SQUARE_ROOT ||
                    averaging_method == ECCKD_AVG_LOGARITHMIC,
                ECCKD_PARAMETER_ERROR, "Averaging method %d not understood", averaging_method);  // find_g_points.cpp:111-113
  ECCKD_REQUIRE(nlay > 0 && n > 0 && stride >= n, ECCKD_PARAMETER_ERROR, "%s: bad dimensions", who);
  IntervalParams p = {};
  p.sw = 0; p.method = averaging_method; p.nlay = nlay; p.n = n; p.stride = stride;
  trans_const256_init(&p.tc, ECCKD_LW_DIFFUSIVITY);
  p.pressure_hl = d_pressure_hl; p.planck_hl = d_planck_hl; p.surf_emissivity = d_surf_emissivity;
  p.surf_planck = d_surf_planck; p.bg_od = d_bg_optical_depth; p.metric = d_metric; p.flux_dn_surf = d_flux_dn_surf;
  p.flux_up_toa = d_flux_up_toa; p.hr = d_hr; p.layer_weight = d_layer_weight; p.flux_weight = flux_weight;
  p.od_scale = 1.0; p.error = d_error; p.fit = d_optical_depth_fit;
  ECCKD_TRY(setup_batch(ctx, p, nint, h_i1, h_i2, n, who));
  interval_fit_partial_kernel<<<p.nitems, kTile, 0, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  interval_fit_finish_kernel<<<p.nint, 512, finish_smem(p.nlay), ctx->stream>>>(p, 0);
  ECCKD_LAUNCH_CHECK(ctx);
  return launch_flux_and_error(ctx, p);
}

ECCKD_API int ecckd_interval_error_batch_sw_dev(
    ecckd_ctx* ctx, int averaging_method, int nlay, int64_t n, int64_t stride, double cos_sza,
    const double* d_pressure_hl, const double* d_ssi, double albedo, const double* d_bg_optical_depth,
    const double* d_metric, const double* d_flux_dn_surf, const double* d_flux_up_toa, const double* d_hr,
    double flux_weight, const double* d_layer_weight, const double* d_flux_dn_surf_low,
    const double* d_flux_up_toa_low, const double* d_hr_low, const double* d_flux_dn_surf_high,
    const double* d_flux_up_toa_high, const double* d_hr_high, double min_scaling, double max_scaling, int nint,
    const int64_t* h_i1, const int64_t* h_i2, double* d_error, double* d_optical_depth_fit) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_interval_error_batch_sw";
  ECCKD_REQUIRE(ctx && d_pressure_hl && d_ssi && d_bg_optical_depth && d_metric && d_layer_weight && d_error,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  const bool tt = averaging_method == ECCKD_AVG_TOTAL_TRANSMISSION;
  ECCKD_REQUIRE(averaging_method == ECCKD_AVG_LINEAR || averaging_method == ECCKD_AVG_TRANSMISSION ||
                    averaging_method == ECCKD_AVG_TRANSMISSION_2 || averaging_method == ECCKD_AVG_SQUARE_ROOT ||
                    averaging_method == ECCKD_AVG_LOGARITHMIC || tt,
                ECCKD_PARAMETER_ERROR, "Averaging method %d not understood", averaging_method);  // find_g_points.cpp:168-170
  if (tt) {
    ECCKD_REQUIRE(d_flux_dn_surf_low && d_flux_up_toa_low && d_hr_low && d_flux_dn_surf_high && d_flux_up_toa_high &&
                      d_hr_high,
                  ECCKD_PARAMETER_ERROR, "%s: total-transmission needs the low/high reference sets", who);
  } else {
    ECCKD_REQUIRE(d_flux_dn_surf && d_flux_up_toa && d_hr, ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  }
  ECCKD_REQUIRE(nlay > 0 && n > 0 && stride >= n && cos_sza > 0.0, ECCKD_PARAMETER_ERROR, "%s: bad dimensions", who);
  IntervalParams p = {};
  p.sw = 1; p.method = averaging_method; p.nlay = nlay; p.n = n; p.stride = stride; p.cos_sza = cos_sza; p.albedo = albedo;
  p.pressure_hl = d_pressure_hl; p.ssi = d_ssi; p.bg_od = d_bg_optical_depth; p.metric = d_metric;
  p.layer_weight = d_layer_weight; p.flux_weight = flux_weight; p.od_scale = 1.0; p.error = d_error;
  p.fit = d_optical_depth_fit;
  p.flux_dn_surf = tt ? d_flux_dn_surf_low : d_flux_dn_surf;
  p.flux_up_toa = tt ? d_flux_up_toa_low : d_flux_up_toa;
  p.hr = tt ? d_hr_low : d_hr;
  ECCKD_TRY(setup_batch(ctx, p, nint, h_i1, h_i2, n, who));
  interval_fit_partial_kernel<<<p.nitems, kTile, 0, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  if (!tt) {
    interval_fit_finish_kernel<<<p.nint, 512, finish_smem(p.nlay), ctx->stream>>>(p, 0);
    ECCKD_LAUNCH_CHECK(ctx);
    return launch_flux_and_error(ctx, p);
  }
  // total-transmission (find_g_points.cpp:334-384): one fit, errors against the low and high sets
  const size_t smem = (size_t)2 * (nlay + 1) * kWarps * sizeof(double);
  interval_tt_sweep_kernel<<<p.nitems, kTile, smem, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  interval_fit_finish_kernel<<<p.nint, 512, finish_smem(p.nlay), ctx->stream>>>(p, 1);
  ECCKD_LAUNCH_CHECK(ctx);
  p.od_scale = min_scaling;
  p.error_mode = 1;
  ECCKD_TRY(launch_flux_and_error(ctx, p));
  // truth sums of the high set, then the second evaluation
  p.flux_dn_surf = d_flux_dn_surf_high; p.flux_up_toa = d_flux_up_toa_high; p.hr = d_hr_high;
  interval_fit_partial_kernel<<<p.nitems, kTile, 0, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  interval_fit_finish_kernel<<<p.nint, 512, finish_smem(p.nlay), ctx->stream>>>(p, 2);
  ECCKD_LAUNCH_CHECK(ctx);
  p.od_scale = max_scaling;
  p.error_mode = 2;
  return launch_flux_and_error(ctx, p);
}

// Broadband solvers: one interval covering the whole array, fit = grey optical depth.
static int bb_solver(ecckd_ctx* ctx, IntervalParams& p, const double* d_grey_od, double* d_flux_dn, double* d_flux_up,
                     const char* who) {
  const int nlev = p.nlay + 1;
  const int64_t i1 = 0, i2 = p.n - 1;
  p.od_scale = 1.0;
  p.fit = const_cast<double*>(d_grey_od);
  ECCKD_TRY(setup_batch(ctx, p, 1, &i1, &i2, p.n, who));
  void* d_bb = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)2 * nlev * sizeof(double), &d_bb));
  p.bb_flux = static_cast<double*>(d_bb);
  p.error = nullptr;
  if (p.sw) {  // level 0 of the SW needs sum(ssi)
    p.flux_dn_surf = p.ssi; p.flux_up_toa = p.ssi; p.hr = p.bg_od; p.metric = p.bg_od; p.method = ECCKD_AVG_LINEAR;
    interval_fit_partial_kernel<<<p.nitems, kTile, 0, ctx->stream>>>(p);
    ECCKD_LAUNCH_CHECK(ctx);
    interval_fit_finish_kernel<<<1, 512, finish_smem(p.nlay), ctx->stream>>>(p, 2);
    ECCKD_LAUNCH_CHECK(ctx);
  } else {
    p.truth = nullptr;
  }
  ECCKD_TRY(launch_flux_and_error(ctx, p));
  ECCKD_CUDA_OK(cudaMemcpyAsync(d_flux_dn, p.bb_flux, nlev * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_flux_up)
    ECCKD_CUDA_OK(cudaMemcpyAsync(d_flux_up, p.bb_flux + nlev, nlev * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  return ECCKD_OK;
}

ECCKD_API int ecckd_radiative_transfer_lw_bb_dev(ecckd_ctx* ctx, int nlay, int64_t n, int64_t stride,
                                                 const double* d_planck, const double* d_spectral_od,
                                                 const double* d_grey_od, const double* d_surf_emissivity,
                                                 const double* d_surf_planck, double* d_flux_dn, double* d_flux_up) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_radiative_transfer_lw_bb";
  ECCKD_REQUIRE(ctx && d_planck && d_spectral_od && d_grey_od && d_surf_emissivity && d_surf_planck && d_flux_dn &&
                    d_flux_up,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE(nlay > 0 && n > 0 && stride >= n, ECCKD_PARAMETER_ERROR, "%s: bad dimensions", who);
  IntervalParams p = {};
  p.sw = 0; p.nlay = nlay; p.n = n; p.stride = stride; p.planck_hl = d_planck; p.bg_od = d_spectral_od;
  trans_const256_init(&p.tc, ECCKD_LW_DIFFUSIVITY);
  p.surf_emissivity = d_surf_emissivity; p.surf_planck = d_surf_planck;
  return bb_solver(ctx, p, d_grey_od, d_flux_dn, d_flux_up, who);
}

ECCKD_API int ecckd_radiative_transfer_sw_bb_dev(ecckd_ctx* ctx, int nlay, int64_t n, int64_t stride, double cos_sza,
                                                 const double* d_ssi, const double* d_spectral_od,
                                                 const double* d_grey_od, double albedo, double* d_flux_dn,
                                                 double* d_flux_up) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_radiative_transfer_sw_bb";
  ECCKD_REQUIRE(ctx && d_ssi && d_spectral_od && d_grey_od && d_flux_dn, ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE(nlay > 0 && n > 0 && stride >= n && cos_sza > 0.0, ECCKD_PARAMETER_ERROR, "%s: bad dimensions", who);
  IntervalParams p = {};
  p.sw = 1; p.nlay = nlay; p.n = n; p.stride = stride; p.ssi = d_ssi; p.bg_od = d_spectral_od; p.cos_sza = cos_sza;
  // radiative_transfer_direct_sw_bb has no upwelling; the norayleigh variant multiplies by the albedo as given
  p.albedo = d_flux_up ? albedo : 0.0;
  if (d_flux_up && !(albedo > 0.0)) p.albedo = 0.0;
  return bb_solver(ctx, p, d_grey_od, d_flux_dn, d_flux_up, who);
}

ECCKD_API int ecckd_cost_function_ckd_lw_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_pressure_hl, const double* d_planck_hl,
    const double* d_surf_emissivity, const double* d_surf_planck, const double* d_optical_depth,
    const double* d_flux_dn, const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    double spectral_boundary_weight, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_cost_profile, double* d_cost_total, double* d_penalty_total) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_cost_function_ckd_lw";
  ECCKD_REQUIRE(ctx && d_pressure_hl && d_planck_hl && d_surf_emissivity && d_surf_planck && d_optical_depth &&
                    d_flux_dn && d_flux_up && d_hr,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE((d_relative_ckd_flux_dn == nullptr) == (d_relative_ckd_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                "%s: relative-to fluxes must both be given or both be NULL", who);
  CkdCostParams p = {};
  p.sw = 0; p.nprof = nprof; p.nlay = nlay; p.ng = ng; p.nband = nband; p.nemis = d_band_mapping ? nband : ng;
  p.up_in_bb = 1;
  p.pressure_hl = d_pressure_hl; p.planck_hl = d_planck_hl; p.surf_emissivity = d_surf_emissivity;
  p.surf_planck = d_surf_planck; p.optical_depth = d_optical_depth; p.flux_dn = d_flux_dn; p.flux_up = d_flux_up;
  p.hr = d_hr; p.spectral_flux_dn_surf = d_spectral_flux_dn_surf; p.spectral_flux_up_toa = d_spectral_flux_up_toa;
  p.layer_weight = d_layer_weight; p.rel_dn = d_relative_ckd_flux_dn; p.rel_up = d_relative_ckd_flux_up;
  p.band_mapping = d_band_mapping; p.flux_weight = flux_weight; p.flux_profile_weight = flux_profile_weight;
  p.broadband_weight = broadband_weight; p.spectral_boundary_weight = spectral_boundary_weight;
  p.negative_od_penalty = negative_od_penalty; p.pressure_weight_power = pressure_weight_power;
  return launch_ckd_cost(ctx, p, d_cost_profile, d_cost_total, d_penalty_total, nullptr, who);
}

ECCKD_API int ecckd_cost_function_ckd_lw_grad_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_pressure_hl, const double* d_planck_hl,
    const double* d_surf_emissivity, const double* d_surf_planck, const double* d_optical_depth,
    const double* d_flux_dn, const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    double spectral_boundary_weight, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_grad_optical_depth) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_cost_function_ckd_lw_grad";
  ECCKD_REQUIRE(ctx && d_pressure_hl && d_planck_hl && d_surf_emissivity && d_surf_planck && d_optical_depth &&
                    d_flux_dn && d_flux_up && d_hr && d_grad_optical_depth,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE((d_relative_ckd_flux_dn == nullptr) == (d_relative_ckd_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                "%s: relative-to fluxes must both be given or both be NULL", who);
  ECCKD_REQUIRE(nprof > 0 && nlay > 0 && ng > 0 && nband > 0 && (d_band_mapping || nband == ng), ECCKD_PARAMETER_ERROR,
                "%s: bad dimensions", who);
  CkdCostParams p = {};
  trans_const256_init(&p.tc, ECCKD_LW_DIFFUSIVITY);
  p.sw = 0; p.nprof = nprof; p.nlay = nlay; p.ng = ng; p.nband = nband; p.nemis = d_band_mapping ? nband : ng;
  p.up_in_bb = 1;
  p.pressure_hl = d_pressure_hl; p.planck_hl = d_planck_hl; p.surf_emissivity = d_surf_emissivity;
  p.surf_planck = d_surf_planck; p.optical_depth = d_optical_depth; p.flux_dn = d_flux_dn; p.flux_up = d_flux_up;
  p.hr = d_hr; p.spectral_flux_dn_surf = d_spectral_flux_dn_surf; p.spectral_flux_up_toa = d_spectral_flux_up_toa;
  p.layer_weight = d_layer_weight; p.rel_dn = d_relative_ckd_flux_dn; p.rel_up = d_relative_ckd_flux_up;
  p.band_mapping = d_band_mapping; p.flux_weight = flux_weight; p.flux_profile_weight = flux_profile_weight;
  p.broadband_weight = broadband_weight; p.spectral_boundary_weight = spectral_boundary_weight;
  p.negative_od_penalty = negative_od_penalty; p.pressure_weight_power = pressure_weight_power;
  const int nlev = nlay + 1;
  const size_t ndbl = (size_t)2 * nlev * ng + (size_t)2 * nlev * nband + (size_t)nlay * nband + 2 * nlay + 2 * nlev + nlay + 256;
  const size_t smem = ndbl * sizeof(double);
  ECCKD_REQUIRE(smem + sizeof(CkdCostParams) + 64 <= ctx->smem_optin, ECCKD_PARAMETER_ERROR,
                "%s: %d g points x %d levels need %zu B of shared memory per profile (limit %zu)", who, ng, nlev, smem,
                ctx->smem_optin);
  ECCKD_CUDA_OK(cudaFuncSetAttribute(ckd_cost_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = std::min(1024, std::max(128, (ng + 31) / 32 * 32));
  ckd_cost_grad_kernel<<<nprof, threads, smem, ctx->stream>>>(p, d_grad_optical_depth);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_cost_function_ckd_sw_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_cos_sza, const double* d_pressure_hl,
    const double* d_ssi, int nalbedo, const double* h_albedo, const double* d_optical_depth, const double* d_flux_dn,
    const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    const double* d_spectral_boundary_weights, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_cost_fn_per_band, double* d_cost_profile, double* d_cost_total,
    double* d_penalty_total) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_cost_function_ckd_sw";
  (void)d_spectral_flux_up_toa;
  ECCKD_REQUIRE(ctx && d_cos_sza && d_pressure_hl && d_ssi && h_albedo && nalbedo > 0 && d_optical_depth && d_flux_dn &&
                    d_flux_up && d_hr,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE((d_relative_ckd_flux_dn == nullptr) == (d_relative_ckd_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                "%s: relative-to fluxes must both be given or both be NULL", who);
  bool all_le0 = true, all_gt0 = true;  // calc_cost_function_sw.cpp:149,250
  for (int k = 0; k < nalbedo; ++k) {
    if (!(h_albedo[k] <= 0.0)) all_le0 = false;
    if (!(h_albedo[k] > 0.0)) all_gt0 = false;
  }
  ECCKD_REQUIRE(all_le0 || d_band_mapping, ECCKD_PARAMETER_ERROR,
                "%s: a non-zero albedo is indexed by band_mapping (calc_cost_function_sw.cpp:156)", who);
  void* d_alb = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)nalbedo * sizeof(double), &d_alb));
  ECCKD_TRY(upload_small(ctx, d_alb, h_albedo, (size_t)nalbedo * sizeof(double)));
  CkdCostParams p = {};
  p.sw = 1; p.nprof = nprof; p.nlay = nlay; p.ng = ng; p.nband = nband; p.nalbedo = nalbedo;
  p.up_in_bb = all_gt0 ? 1 : 0; p.direct_only = all_le0 ? 1 : 0;
  p.cos_sza = d_cos_sza; p.pressure_hl = d_pressure_hl; p.ssi = d_ssi; p.albedo = static_cast<double*>(d_alb);
  p.optical_depth = d_optical_depth; p.flux_dn = d_flux_dn; p.flux_up = d_flux_up; p.hr = d_hr;
  p.spectral_flux_dn_surf = d_spectral_flux_dn_surf; p.spectral_boundary_weights = d_spectral_boundary_weights;
  p.layer_weight = d_layer_weight; p.rel_dn = d_relative_ckd_flux_dn; p.rel_up = d_relative_ckd_flux_up;
  p.band_mapping = d_band_mapping; p.flux_weight = flux_weight; p.flux_profile_weight = flux_profile_weight;
  p.broadband_weight = broadband_weight; p.negative_od_penalty = negative_od_penalty;
  p.pressure_weight_power = pressure_weight_power;
  return launch_ckd_cost(ctx, p, d_cost_profile, d_cost_total, d_penalty_total, d_cost_fn_per_band, who);
}

ECCKD_API int ecckd_cost_function_ckd_sw_grad_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_cos_sza, const double* d_pressure_hl,
    const double* d_ssi, int nalbedo, const double* h_albedo, const double* d_optical_depth, const double* d_flux_dn,
    const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf, double flux_weight,
    double flux_profile_weight, double broadband_weight, const double* d_spectral_boundary_weights,
    const double* d_layer_weight, double pressure_weight_power, double negative_od_penalty,
    const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up, const int32_t* d_band_mapping,
    double* d_grad_optical_depth) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_cost_function_ckd_sw_grad";
  ECCKD_REQUIRE(ctx && d_cos_sza && d_pressure_hl && d_ssi && h_albedo && nalbedo > 0 && d_optical_depth && d_flux_dn &&
                    d_flux_up && d_hr && d_grad_optical_depth,
                ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  ECCKD_REQUIRE((d_relative_ckd_flux_dn == nullptr) == (d_relative_ckd_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                "%s: relative-to fluxes must both be given or both be NULL", who);
  ECCKD_REQUIRE(nprof > 0 && nlay > 0 && ng > 0 && nband > 0 && (d_band_mapping || nband == ng), ECCKD_PARAMETER_ERROR,
                "%s: bad dimensions", who);
  bool all_le0 = true, all_gt0 = true;  // calc_cost_function_sw.cpp:149,250
  for (int k = 0; k < nalbedo; ++k) {
    if (!(h_albedo[k] <= 0.0)) all_le0 = false;
    if (!(h_albedo[k] > 0.0)) all_gt0 = false;
  }
  ECCKD_REQUIRE(all_le0 || d_band_mapping, ECCKD_PARAMETER_ERROR,
                "%s: a non-zero albedo is indexed by band_mapping (calc_cost_function_sw.cpp:156)", who);
  void* d_alb = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)nalbedo * sizeof(double), &d_alb));
  ECCKD_TRY(upload_small(ctx, d_alb, h_albedo, (size_t)nalbedo * sizeof(double)));
  CkdCostParams p = {};
  trans_const256_init(&p.tc, ECCKD_LW_DIFFUSIVITY);
  p.sw = 1; p.nprof = nprof; p.nlay = nlay; p.ng = ng; p.nband = nband; p.nalbedo = nalbedo;
  p.up_in_bb = all_gt0 ? 1 : 0; p.direct_only = all_le0 ? 1 : 0;
  p.cos_sza = d_cos_sza; p.pressure_hl = d_pressure_hl; p.ssi = d_ssi; p.albedo = static_cast<double*>(d_alb);
  p.optical_depth = d_optical_depth; p.flux_dn = d_flux_dn; p.flux_up = d_flux_up; p.hr = d_hr;
  p.spectral_flux_dn_surf = d_spectral_flux_dn_surf; p.spectral_boundary_weights = d_spectral_boundary_weights;
  p.layer_weight = d_layer_weight; p.rel_dn = d_relative_ckd_flux_dn; p.rel_up = d_relative_ckd_flux_up;
  p.band_mapping = d_band_mapping; p.flux_weight = flux_weight; p.flux_profile_weight = flux_profile_weight;
  p.broadband_weight = broadband_weight; p.negative_od_penalty = negative_od_penalty;
  p.pressure_weight_power = pressure_weight_power;
  const int nlev = nlay + 1;
  const size_t ndbl = (size_t)2 * nlev * ng + (size_t)2 * nlev * nband + (size_t)nlay * nband + 2 * nlay + 2 * nlev + nlay + 256;
  const size_t smem = ndbl * sizeof(double);
  ECCKD_REQUIRE(smem + sizeof(CkdCostParams) + 64 <= ctx->smem_optin, ECCKD_PARAMETER_ERROR,
                "%s: %d g points x %d levels need %zu B of shared memory per profile (limit %zu)", who, ng, nlev, smem,
                ctx->smem_optin);
  ECCKD_CUDA_OK(cudaFuncSetAttribute(ckd_cost_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = std::min(1024, std::max(128, (ng + 31) / 32 * 32));
  ckd_cost_grad_kernel<<<nprof, threads, smem, ctx->stream>>>(p, d_grad_optical_depth);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// ------------------------------------------------------------------ several scenes in one launch
ECCKD_API int ecckd_cost_function_ckd_scenes_dev(ecckd_ctx* ctx, int is_sw, int nscene, const ecckd_cost_scene* scenes,
                                                 int nlay, int ng, int nband, double flux_weight,
                                                 double flux_profile_weight, double broadband_weight,
                                                 double spectral_boundary_weight, double pressure_weight_power,
                                                 double negative_od_penalty, const int32_t* d_band_mapping,
                                                 double* d_cost_fn_per_band, double* d_cost_scene,
                                                 double* d_penalty_scene, double* d_cost_total) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_cost_function_ckd_scenes";
  ECCKD_REQUIRE(ctx && scenes && nscene > 0, ECCKD_PARAMETER_ERROR, "%s: NULL argument", who);
  std::vector<CkdCostParams> ps((size_t)nscene);
  std::vector<int> start((size_t)nscene + 1, 0);
  std::vector<double> albedos;
  std::vector<size_t> alb_off((size_t)nscene, 0);
  size_t smem = 0;
  for (int k = 0; k < nscene; ++k) {
    const ecckd_cost_scene& sc = scenes[k];
    CkdCostParams& p = ps[k];
    p = CkdCostParams{};
    ECCKD_REQUIRE(sc.d_pressure_hl && sc.d_optical_depth && sc.d_flux_dn && sc.d_flux_up && sc.d_hr, ECCKD_PARAMETER_ERROR,
                  "%s: scene %d: NULL argument", who, k);
    ECCKD_REQUIRE((sc.d_relative_ckd_flux_dn == nullptr) == (sc.d_relative_ckd_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                  "%s: scene %d: relative-to fluxes must both be given or both be NULL", who, k);
    p.sw = is_sw ? 1 : 0; p.nprof = sc.nprof; p.nlay = nlay; p.ng = ng; p.nband = nband;
    p.pressure_hl = sc.d_pressure_hl; p.optical_depth = sc.d_optical_depth; p.flux_dn = sc.d_flux_dn;
    p.flux_up = sc.d_flux_up; p.hr = sc.d_hr; p.spectral_flux_dn_surf = sc.d_spectral_flux_dn_surf;
    p.layer_weight = sc.d_layer_weight; p.rel_dn = sc.d_relative_ckd_flux_dn; p.rel_up = sc.d_relative_ckd_flux_up;
    p.band_mapping = d_band_mapping; p.flux_weight = flux_weight; p.flux_profile_weight = flux_profile_weight;
    p.broadband_weight = broadband_weight; p.negative_od_penalty = negative_od_penalty;
    p.pressure_weight_power = pressure_weight_power;
    if (!is_sw) {
      ECCKD_REQUIRE(sc.d_planck_hl && sc.d_surf_emissivity && sc.d_surf_planck, ECCKD_PARAMETER_ERROR,
                    "%s: scene %d: NULL longwave argument", who, k);
      p.nemis = d_band_mapping ? nband : ng;
      p.up_in_bb = 1;
      p.planck_hl = sc.d_planck_hl; p.surf_emissivity = sc.d_surf_emissivity; p.surf_planck = sc.d_surf_planck;
      p.spectral_flux_up_toa = sc.d_spectral_flux_up_toa; p.spectral_boundary_weight = spectral_boundary_weight;
    } else {
      ECCKD_REQUIRE(sc.d_cos_sza && sc.d_ssi && sc.h_albedo && sc.nalbedo > 0, ECCKD_PARAMETER_ERROR,
                    "%s: scene %d: NULL shortwave argument", who, k);
      bool all_le0 = true, all_gt0 = true;  // calc_cost_function_sw.cpp:149,250
      for (int j = 0; j < sc.nalbedo; ++j) {
        if (!(sc.h_albedo[j] <= 0.0)) all_le0 = false;
        if (!(sc.h_albedo[j] > 0.0)) all_gt0 = false;
      }
      ECCKD_REQUIRE(all_le0 || d_band_mapping, ECCKD_PARAMETER_ERROR,
                    "%s: a non-zero albedo is indexed by band_mapping (calc_cost_function_sw.cpp:156)", who);
      p.nalbedo = sc.nalbedo; p.up_in_bb = all_gt0 ? 1 : 0; p.direct_only = all_le0 ? 1 : 0;
      p.cos_sza = sc.d_cos_sza; p.ssi = sc.d_ssi; p.spectral_boundary_weights = sc.d_spectral_boundary_weights;
      alb_off[k] = albedos.size();
      albedos.insert(albedos.end(), sc.h_albedo, sc.h_albedo + sc.nalbedo);
    }
    size_t sm_k = 0;
    ECCKD_TRY(ckd_cost_smem(ctx, p, who, &sm_k));
    smem = std::max(smem, sm_k);
    start[k + 1] = start[k] + sc.nprof;
  }
  const int ntot = start[nscene];
  // scratch: per-profile penalties, costs (where the caller does not want them) and per-band costs,
  // per-scene band sums, the albedos, the scene table
  const bool bands = is_sw && d_cost_fn_per_band;
  const size_t b_prof = ((size_t)ntot * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_band = bands ? (((size_t)ntot * nband * sizeof(double) + 255) & ~(size_t)255) : 0;
  const size_t b_sband = bands ? (((size_t)nscene * nband * sizeof(double) + 255) & ~(size_t)255) : 0;
  const size_t b_alb = (albedos.size() * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_scal = (2 * (size_t)nscene * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_tab = ((size_t)nscene * sizeof(CkdCostParams) + 255) & ~(size_t)255;
  const size_t b_start = (((size_t)nscene + 1) * sizeof(int) + 255) & ~(size_t)255;
  void* base = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 2, 2 * b_prof + b_band + b_sband + b_alb + b_scal + b_tab + b_start, &base));
  char* c = static_cast<char*>(base);
  double* d_pen = reinterpret_cast<double*>(c);
  double* d_cost = reinterpret_cast<double*>(c + b_prof);
  double* d_band = bands ? reinterpret_cast<double*>(c + 2 * b_prof) : nullptr;
  double* d_sband = bands ? reinterpret_cast<double*>(c + 2 * b_prof + b_band) : nullptr;
  double* d_alb = reinterpret_cast<double*>(c + 2 * b_prof + b_band + b_sband);
  double* d_scal = reinterpret_cast<double*>(c + 2 * b_prof + b_band + b_sband + b_alb);
  CkdCostParams* d_tab = reinterpret_cast<CkdCostParams*>(c + 2 * b_prof + b_band + b_sband + b_alb + b_scal);
  int* d_start = reinterpret_cast<int*>(c + 2 * b_prof + b_band + b_sband + b_alb + b_scal + b_tab);
  for (int k = 0; k < nscene; ++k) {
    CkdCostParams& p = ps[k];
    p.penalty_profile = d_pen + start[k];
    p.cost_profile = scenes[k].d_cost_profile ? scenes[k].d_cost_profile : d_cost + start[k];
    p.cost_band_profile = bands ? d_band + (size_t)start[k] * nband : nullptr;
    if (is_sw) p.albedo = d_alb + alb_off[k];
  }
  if (!albedos.empty()) ECCKD_TRY(upload_small(ctx, d_alb, albedos.data(), albedos.size() * sizeof(double)));
  ECCKD_TRY(upload_small(ctx, d_tab, ps.data(), ps.size() * sizeof(CkdCostParams)));
  ECCKD_TRY(upload_small(ctx, d_start, start.data(), start.size() * sizeof(int)));
  ECCKD_CUDA_OK(cudaFuncSetAttribute(ckd_cost_scenes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = std::min(1024, std::max(128, (ng + 31) / 32 * 32));
  ckd_cost_scenes_kernel<<<ntot, threads, smem, ctx->stream>>>(d_tab, d_start, nscene);
  ECCKD_LAUNCH_CHECK(ctx);
  double* cs = d_cost_scene ? d_cost_scene : d_scal;
  double* pn = d_penalty_scene ? d_penalty_scene : d_scal + nscene;
  ckd_cost_scene_sums_kernel<<<nscene, 64, 0, ctx->stream>>>(d_tab, nband, cs, pn, d_sband);
  ECCKD_LAUNCH_CHECK(ctx);
  if (d_cost_total || bands) {
    ckd_cost_scene_total_kernel<<<1, 64, 0, ctx->stream>>>(nscene, nband, cs, d_sband, d_cost_total,
                                                           bands ? d_cost_fn_per_band : nullptr);
    ECCKD_LAUNCH_CHECK(ctx);
  }
  return ECCKD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Background (prior) term of the optimize_lut cost function, CkdModel<true>::calc_background_cost_function
// (ckd_model.cpp:838-877).  A gas's slice of the state vector is [nx][ng], g point fastest (:851-855).
// background_grad_kernel: one thread per (row i, g point): the row of inv_background_shape times the g point's column
// of delta_x in ascending j (coalesced over g, the matrix row broadcast), scaled by 1/error(g)^2 (:857-858), added to
// the gradient (:861); the product delta_x * gradient_local goes to a scratch array.  background_cost_kernel: one
// thread per g point adds its column of products in ascending i (the dot_product of :859), thread 0 then adds the
// per-g halves in ascending g (:864): fixed order, no atomics.
namespace ecckd {
__global__ void background_grad_kernel(int nx, int ng, const double* __restrict__ shape, const double* __restrict__ err,
                                       const double* __restrict__ dx, double* __restrict__ grad,
                                       double* __restrict__ prod) {
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (ig >= ng) return;
  const double* row = shape + (size_t)i * nx;
  double s = 0.0;
  for (int j = 0; j < nx; ++j) s += row[j] * dx[(size_t)j * ng + ig];
  const double e = err[ig];
  const double gl = (1.0 / (e * e)) * s;
  const size_t k = (size_t)i * ng + ig;
  if (grad) grad[k] += gl;
  prod[k] = dx[k] * gl;
}
__global__ void __launch_bounds__(1024) background_cost_kernel(int nx, int ng, const double* __restrict__ prod,
                                                                double* __restrict__ cost) {
  extern __shared__ double s_half[];  // [ng]
  for (int ig = threadIdx.x; ig < ng; ig += blockDim.x) {
    double d = 0.0;
    for (int i = 0; i < nx; ++i) d += prod[(size_t)i * ng + ig];
    s_half[ig] = 0.5 * d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
    for (int ig = 0; ig < ng; ++ig) c += s_half[ig];
    *cost = c;
  }
}
// the Rayleigh part (:868-873): 0.5 * sum_g delta_x(g) * (inv_background(g) * delta_x(g)) in ascending g
__global__ void background_rayleigh_kernel(int ng, const double* __restrict__ inv_b, const double* __restrict__ dx,
                                           double* __restrict__ cost) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double d = 0.0;
  for (int ig = 0; ig < ng; ++ig) d += dx[ig] * (inv_b[ig] * dx[ig]);
  *cost = 0.5 * d;
}
}  // namespace ecckd

ECCKD_API int ecckd_background_cost_rayleigh_dev(ecckd_ctx* ctx, int ng, const double* d_rayleigh_inv_background,
                                                 const double* d_delta_x, double* h_cost) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_rayleigh_inv_background && d_delta_x && h_cost && ng > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_background_cost_rayleigh: bad argument");
  void* d_tmp = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 4, sizeof(double), &d_tmp));
  ecckd::background_rayleigh_kernel<<<1, 32, 0, ctx->stream>>>(ng, d_rayleigh_inv_background, d_delta_x,
                                                                static_cast<double*>(d_tmp));
  ECCKD_LAUNCH_CHECK(ctx);
  ECCKD_CUDA_OK(cudaMemcpyAsync(h_cost, d_tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return ECCKD_OK;
}

ECCKD_API int ecckd_background_cost_gas_dev(ecckd_ctx* ctx, int nx, int ng, const double* d_inv_background_shape,
                                            const double* d_background_error, const double* d_delta_x,
                                            double* d_gradient, double* h_cost) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_inv_background_shape && d_background_error && d_delta_x && h_cost, ECCKD_PARAMETER_ERROR,
                "ecckd_background_cost_gas: NULL argument");
  ECCKD_REQUIRE(nx > 0 && ng > 0 && nx <= 65535 && (size_t)ng * sizeof(double) <= 48 * 1024, ECCKD_PARAMETER_ERROR,
                "ecckd_background_cost_gas: bad dimensions");
  void* d_tmp = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 4, ((size_t)nx * ng + 1) * sizeof(double), &d_tmp));
  double* d_prod = static_cast<double*>(d_tmp);
  double* d_cost = d_prod + (size_t)nx * ng;
  const int bs = ng >= 128 ? 128 : 32 * ((ng + 31) / 32);
  ecckd::background_grad_kernel<<<dim3((unsigned)((ng + bs - 1) / bs), (unsigned)nx), bs, 0, ctx->stream>>>(
      nx, ng, d_inv_background_shape, d_background_error, d_delta_x, d_gradient, d_prod);
  ecckd::background_cost_kernel<<<1, std::min(1024, 32 * ((ng + 31) / 32)), (size_t)ng * sizeof(double), ctx->stream>>>(
      nx, ng, d_prod, d_cost);
  ctx->launches += 1;
  ECCKD_LAUNCH_CHECK(ctx);
  ECCKD_CUDA_OK(cudaMemcpyAsync(h_cost, d_cost, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return ECCKD_OK;
}
