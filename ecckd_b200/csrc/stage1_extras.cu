// stage1_extras.cu - small per-LUT products of create_look_up_table that are not reductions
// over the plan: the base g-point split (create_look_up_table.cpp:162-224) and the Rayleigh
// molar scattering coefficients (rayleigh_scattering.h:23-43, ckd_model.h:368-385).
#include <algorithm>

#include "common.cuh"

namespace ecckd {

// rayleigh_scattering.h:28-40 (Bucholtz 1995), m2 mol-1
__device__ __forceinline__ double rayleigh_coeff(double wavenumber_cm_1) {
  const double wl = 10000.0 / wavenumber_cm_1;
  const double s = (wl < 0.5) ? 3.01577e-32 * pow(wl, -(3.55212 + 1.35579 * wl + 0.11563 / wl))
                              : 4.01061e-32 * pow(wl, -(3.99668 + 0.00110298 * wl + 0.0271393 / wl));
  return s * 6.02214076e23;  // AVOGADRO_CONSTANT, constants.h:26
}

__global__ void rayleigh_coeff_kernel(int64_t n, const double* __restrict__ wn, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rayleigh_coeff(wn[i]);
}

// ckd_model.h:368-385: thread = g point, bins summed in ascending order (the matrix-vector
// products gpoint_fraction ** (...) of the reference).
__global__ void rayleigh_gpoint_kernel(int ng, int nbin, const double* __restrict__ wn1, const double* __restrict__ wn2,
                                       const double* __restrict__ gpoint_fraction,
                                       const double* __restrict__ ssi_intervals, double* __restrict__ out) {
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  if (ig >= ng) return;
  const double REFERENCE_COS_SZA = 0.5;  // calc_cost_function_sw.h:20
  const double molar_column = 1.0e5 / (ECCKD_ACCEL_GRAVITY * 0.001 * ECCKD_MOLAR_MASS_DRY_AIR);
  double num = 0.0, den = 0.0;
  for (int i = 0; i < nbin; ++i) {
    const double scat = rayleigh_coeff(0.5 * (wn1[i] + wn2[i]));
    const double trans_hr = exp(-(molar_column * scat) / REFERENCE_COS_SZA);
    const double gf = gpoint_fraction[(size_t)ig * nbin + i];
    num += gf * (ssi_intervals[i] * trans_hr);
    den += gf * ssi_intervals[i];
  }
  const double od = -log(fmax(1.0e-14, num / den)) * REFERENCE_COS_SZA;
  out[ig] = od / molar_column;
}

// create_look_up_table.cpp:197-212: renumber the g points above the base one and cut the base
// g point of the band at the given wavenumbers.  bounds[nidx+2] in shared memory.
__global__ void split_base_kernel(int64_t nwav, const double* __restrict__ wn, int32_t* __restrict__ g_point, int ig,
                                  int nidx, const double* __restrict__ bounds) {
  extern __shared__ double s_b[];
  for (int k = threadIdx.x; k < nidx + 2; k += blockDim.x) s_b[k] = bounds[k];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwav) return;
  const int32_t g = g_point[i];
  int32_t out = g > ig ? g + nidx : g;
  if (g == ig) {
    const double w = wn[i];
    // ascending new_ig, later matches overwrite (they cannot overlap: the bounds ascend)
    for (int k = 0; k <= nidx; ++k) {
      if (w >= s_b[k] && w < s_b[k + 1]) out = ig + k;
    }
  }
  g_point[i] = out;
}

// sum(ssi(find(new_g_point == new_ig))) for new_ig = ig..ig+nidx: block partials in a fixed
// geometry, added in block order by the second kernel.
constexpr int kSumBlock = 256;
constexpr int kSumPerBlock = 8192;

__global__ void __launch_bounds__(kSumBlock) masked_sum_partial_kernel(int64_t nwav, const int32_t* __restrict__ g_point,
                                                                       const double* __restrict__ ssi, int ig, int ntarget,
                                                                       double* __restrict__ partial) {
  __shared__ double s_w[kSumBlock / 32];
  const int64_t a = (int64_t)blockIdx.x * kSumPerBlock, b = min(a + kSumPerBlock, nwav);
  for (int t = 0; t < ntarget; ++t) {
    double s = 0.0;
    for (int64_t i = a + threadIdx.x; i < b; i += kSumBlock) {
      if (g_point[i] == ig + t) s += ssi[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
      for (int w = 0; w < kSumBlock / 32; ++w) r += s_w[w];
      partial[(size_t)blockIdx.x * ntarget + t] = r;
    }
    __syncthreads();
  }
}

__global__ void masked_sum_final_kernel(int nblocks, int ntarget, const double* __restrict__ partial,
                                        double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntarget) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * ntarget + t];
  out[t] = s;
}

}  // namespace ecckd

using namespace ecckd;

ECCKD_API int ecckd_rayleigh_molar_scattering_coeff_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wavenumber_cm_1,
                                                        double* d_scat_coeff) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_wavenumber_cm_1 && d_scat_coeff && nwav > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_rayleigh_molar_scattering_coeff: bad argument");
  rayleigh_coeff_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(nwav, d_wavenumber_cm_1, d_scat_coeff);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_rayleigh_molar_scat_gpoints_dev(ecckd_ctx* ctx, int ng, int nbin, const double* d_wavenumber1,
                                                    const double* d_wavenumber2, const double* d_gpoint_fraction,
                                                    const double* d_ssi_intervals, double* d_rayleigh_molar_scat) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_wavenumber1 && d_wavenumber2 && d_gpoint_fraction && d_ssi_intervals && d_rayleigh_molar_scat &&
                    ng > 0 && nbin > 0,
                ECCKD_PARAMETER_ERROR, "ecckd_rayleigh_molar_scat_gpoints: bad argument");
  rayleigh_gpoint_kernel<<<(ng + 63) / 64, 64, 0, ctx->stream>>>(ng, nbin, d_wavenumber1, d_wavenumber2,
                                                                   d_gpoint_fraction, d_ssi_intervals,
                                                                   d_rayleigh_molar_scat);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_split_base_g_points_dev(ecckd_ctx* ctx, int nband, const double* h_band_wn1, const double* h_band_wn2,
                                            int nbound, const double* h_base_wavenumber_boundary, int64_t nwav,
                                            const double* d_wavenumber_cm_1, const double* d_ssi, int cap, int* ng_io,
                                            int32_t* h_band_number, int32_t* d_g_point, double* h_solar_irradiance) {
  ECCKD_DEVICE(ctx);
  const char* who = "ecckd_split_base_g_points";
  ECCKD_REQUIRE(ctx && h_band_wn1 && h_band_wn2 && (nbound == 0 || h_base_wavenumber_boundary) && d_wavenumber_cm_1 &&
                    ng_io && h_band_number && d_g_point && nband > 0 && nwav > 0,
                ECCKD_PARAMETER_ERROR, "%s: bad argument", who);
  ECCKD_REQUIRE((d_ssi == nullptr) == (h_solar_irradiance == nullptr), ECCKD_PARAMETER_ERROR,
                "%s: ssi and solar_irradiance must both be given (shortwave) or both be NULL", who);
  int ng = *ng_io;
  for (int iband = 0; iband < nband; ++iband) {
    std::vector<double> bounds;
    bounds.push_back(h_band_wn1[iband]);
    for (int k = 0; k < nbound; ++k) {  // create_look_up_table.cpp:164-165
      if (h_base_wavenumber_boundary[k] > h_band_wn1[iband] && h_base_wavenumber_boundary[k] < h_band_wn2[iband])
        bounds.push_back(h_base_wavenumber_boundary[k]);
    }
    const int nidx = (int)bounds.size() - 1;
    if (nidx == 0) continue;
    bounds.push_back(h_band_wn2[iband]);
    int ig = -1;
    for (int k = 0; k < ng; ++k)
      if (h_band_number[k] == iband) { ig = k; break; }
    ECCKD_REQUIRE(ig >= 0, ECCKD_PROCESSING_ERROR, "%s: band %d has no g point", who, iband);
    const int new_ng = ng + nidx;
    ECCKD_REQUIRE(new_ng <= cap, ECCKD_PROCESSING_ERROR, "%s: %d g points exceed the capacity %d", who, new_ng, cap);
    // band numbers / solar irradiance of the unchanged g points (:183-194; see the oracle for the
    // last-band note)
    for (int k = ng - 1; k > ig; --k) {
      h_band_number[k + nidx] = h_band_number[k];
      if (h_solar_irradiance) h_solar_irradiance[k + nidx] = h_solar_irradiance[k];
    }
    for (int k = ig + 1; k <= ig + nidx; ++k) h_band_number[k] = iband;
    void* d_b = nullptr;
    ECCKD_TRY(ws_reserve(ctx, 3, bounds.size() * sizeof(double), &d_b));
    ECCKD_TRY(upload_small(ctx, d_b, bounds.data(), bounds.size() * sizeof(double)));
    split_base_kernel<<<(unsigned)((nwav + 255) / 256), 256, bounds.size() * sizeof(double), ctx->stream>>>(
        nwav, d_wavenumber_cm_1, d_g_point, ig, nidx, static_cast<double*>(d_b));
    ECCKD_LAUNCH_CHECK(ctx);
    if (d_ssi) {
      const int nt = nidx + 1;
      const int nblocks = (int)((nwav + kSumPerBlock - 1) / kSumPerBlock);
      void* d_p = nullptr;
      ECCKD_TRY(ws_reserve(ctx, 2, ((size_t)nblocks + 1) * nt * sizeof(double), &d_p));
      double* part = static_cast<double*>(d_p);
      double* d_out = part + (size_t)nblocks * nt;
      masked_sum_partial_kernel<<<nblocks, kSumBlock, 0, ctx->stream>>>(nwav, d_g_point, d_ssi, ig, nt, part);
      ECCKD_LAUNCH_CHECK(ctx);
      masked_sum_final_kernel<<<(nt + 63) / 64, 64, 0, ctx->stream>>>(nblocks, nt, part, d_out);
      ECCKD_LAUNCH_CHECK(ctx);
      ECCKD_CUDA_OK(cudaMemcpyAsync(h_solar_irradiance + ig, d_out, nt * sizeof(double), cudaMemcpyDeviceToHost,
                                    ctx->stream));
      ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    ng = new_ng;
  }
  *ng_io = ng;
  return ECCKD_OK;
}
