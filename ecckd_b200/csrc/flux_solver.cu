// flux_solver.cu - stage 2 of the hot path: CKD optical depths from the
// look-up table and the forward flux solvers that run_ckd and the forward pass
// of optimize_lut evaluate (reference: ckd_model.cpp:923-1145,
// radiative_transfer_{lw,sw}.cpp, heating_rate.h, run_ckd.cpp:119-367).
//
// Layout: every 3-D array is [column][layer or half-level][g] with g fastest,
// exactly as the reference's Array3D.  Threads map to g (fastest) so each warp
// reads/writes contiguous 256-512 B rows; one thread walks the layers of one
// column x g-point, carrying the flux recurrence in registers.  All kernels are
// HBM-bound streams: 32 B per column x layer x g for the LW solver.
#include <algorithm>

#include "common.cuh"
#include "async_copy.cuh"
#include "trans_core.cuh"

struct ecckd_ckd_model {
  ecckd_ctx* ctx = nullptr;
  int ngas = 0, nt = 0, np = 0, ng = 0, nlut = 0;
  bool is_sw = false;
  std::vector<int> conc_dependence, nconc;
  std::vector<double> reference_vmr;
  std::vector<double*> d_molar_abs;  // per gas
  std::vector<double*> d_packed;     // per gas, longwave models: [nconc|1][nt-1][np-1][ng][4] the four table corners
                                     // (it,ip) (it,ip+1) (it+1,ip) (it+1,ip+1) of every interpolation cell side by side
  std::vector<double*> d_vmr_lut;    // per gas (LUT only)
  std::vector<double> log_vmr0, d_log_c;  // per gas (LUT only), host-computed as in ckd_model.cpp:1003-1005
  double* d_log_pressure = nullptr;
  double* d_temperature = nullptr;
  double* d_temperature_planck = nullptr;
  double* d_planck_lut = nullptr;
  double* d_solar_irradiance = nullptr;
  double* d_rayleigh = nullptr;
  double log_p_0 = 0, d_log_p = 0, d_t = 0;          // ckd_model.cpp:948-952
  double planck_t0 = 0, planck_dt = 0;               // ckd_model.cpp:1123-1124
  double solar_sum = 0;                              // sum(solar_irradiance), run_ckd.cpp:98
  void* d_gasdesc = nullptr;                         // device array of GasDev
  double* d_ones = nullptr;                          // [ng] surface emissivity of run_ckd
};

namespace ecckd {

struct GasDev {
  const double* molar_abs;
  const double* vmr;  // per-call: [ncol][nlay] or NULL
  double reference_vmr;
  double log_vmr0;
  double d_log_c;
  int32_t conc_dependence;
  int32_t nconc;
};

// interpolation state of one (column, layer): ckd_model.cpp:962-980
struct Interp {
  int ip0, it0;
  double pw0, pw1, tw0, tw1, simple_weight;
};

__device__ __forceinline__ Interp make_interp(double p0, double p1, double t_fl, double log_p_0, double d_log_p,
                                              double d_t, int nt, int np, const double* __restrict__ temperature) {
  Interp it;
  const double log_pressure_fl = log(0.5 * (p1 + p0));
  double pindex0 = (log_pressure_fl - log_p_0) / d_log_p;
  pindex0 = fmax(0.0, fmin(pindex0, np - 1.0001));
  it.ip0 = (int)pindex0;
  it.pw1 = pindex0 - it.ip0;
  it.pw0 = 1.0 - it.pw1;
  const double t_0 = it.pw0 * temperature[it.ip0] + it.pw1 * temperature[it.ip0 + 1];
  double tindex0 = (t_fl - t_0) / d_t;
  tindex0 = fmax(0.0, fmin(tindex0, nt - 1.0001));
  it.it0 = (int)tindex0;
  it.tw1 = tindex0 - it.it0;
  it.tw0 = 1.0 - it.tw1;
  it.simple_weight = (1.0 / (ECCKD_ACCEL_GRAVITY * 0.001 * ECCKD_MOLAR_MASS_DRY_AIR)) * (p1 - p0);
  return it;
}

// optical depth of one gas at one (column, layer, g): ckd_model.cpp:982-1088
__device__ __forceinline__ double gas_od(const GasDev& gs, const Interp& it, double vmr, bool have_vmr, int nt, int np,
                                         int ng, int g) {
  const size_t sP = (size_t)ng, sT = (size_t)np * ng;
  double weight = 0.0;
  if (have_vmr) {
    weight = (gs.conc_dependence == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - gs.reference_vmr)
                                                                : it.simple_weight * vmr;
  }
  if (gs.conc_dependence == ECCKD_CONC_LUT) {
    double cindex0 = (log(vmr) - gs.log_vmr0) / gs.d_log_c;
    cindex0 = fmax(0.0, fmin(cindex0, gs.nconc - 1.0001));
    const int ic0 = (int)cindex0;
    const double cw1 = cindex0 - ic0;
    const double cw0 = 1.0 - cw1;
    const size_t sC = (size_t)nt * sT;
    const double* k = gs.molar_abs + ic0 * sC + it.it0 * sT + it.ip0 * sP + g;
    const double a = it.tw0 * (it.pw0 * k[0] + it.pw1 * k[sP]) + it.tw1 * (it.pw0 * k[sT] + it.pw1 * k[sT + sP]);
    k += sC;
    const double b = it.tw0 * (it.pw0 * k[0] + it.pw1 * k[sP]) + it.tw1 * (it.pw0 * k[sT] + it.pw1 * k[sT + sP]);
    return weight * (cw0 * a + cw1 * b);
  }
  const double* k = gs.molar_abs + it.it0 * sT + it.ip0 * sP + g;
  const double v = it.tw0 * (it.pw0 * k[0] + it.pw1 * k[sP]) + it.tw1 * (it.pw0 * k[sT] + it.pw1 * k[sT + sP]);
  return ((gs.conc_dependence == ECCKD_CONC_NONE) ? it.simple_weight : weight) * v;
}

// CkdModel::calc_planck_function for one temperature and g: ckd_model.cpp:1124-1142
__device__ __forceinline__ double planck_interp(double t, double t0, double d_t, int nlut, int ng, int g,
                                                const double* __restrict__ lut) {
  const double tindex0 = (t - t0) / d_t;
  if (tindex0 >= 0) {
    const int it0 = min((int)tindex0, nlut - 2);
    const double w1 = tindex0 - it0;
    const double w0 = 1.0 - w1;
    return w0 * lut[(size_t)it0 * ng + g] + w1 * lut[(size_t)(it0 + 1) * ng + g];
  }
  return (t / t0) * lut[g];
}

// LW layer emissivity and Planck-gradient factor: radiative_transfer_lw.cpp:41-43
__device__ __forceinline__ void lw_layer(double tau, double& e, double& f) {
  e = 1.0 - exp(-ECCKD_LW_DIFFUSIVITY * tau);
  f = (e > 1.0e-5) ? (1.0 - e * (1.0 / ECCKD_LW_DIFFUSIVITY) / tau) : (0.5 * e);
}

// ------------------------------------------------------------------ kernels
// thread = (column*nlay + layer, g)
__global__ void ckd_od_kernel(GasDev gs, int nt, int np, int ng, double log_p_0, double d_log_p, double d_t,
                              const double* __restrict__ temperature, int ncol, int nlay,
                              const double* __restrict__ pressure_hl, const double* __restrict__ temperature_fl,
                              double* __restrict__ od, int accumulate) {
  const int g = blockIdx.y * blockDim.x + threadIdx.x;  // grid.x runs over rows (no 65535 limit), grid.y over g
  const int64_t cl = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (g >= ng || cl >= (int64_t)ncol * nlay) return;
  const int icol = (int)(cl / nlay), ip = (int)(cl - (int64_t)icol * nlay);
  const double p0 = pressure_hl[(int64_t)icol * (nlay + 1) + ip];
  const double p1 = pressure_hl[(int64_t)icol * (nlay + 1) + ip + 1];
  const Interp it = make_interp(p0, p1, temperature_fl[cl], log_p_0, d_log_p, d_t, nt, np, temperature);
  const bool have = gs.vmr != nullptr;
  const double v = gas_od(gs, it, have ? gs.vmr[cl] : 0.0, have, nt, np, ng, g);
  double* o = od + cl * ng + g;
  *o = accumulate ? (*o + v) : v;
}

__global__ void clamp_zero_kernel(int64_t n, double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmax(x[i], 0.0);  // run_ckd.cpp:317
}

__global__ void ckd_planck_kernel(int nlut, int ng, double t0, double d_t, const double* __restrict__ lut, int64_t n,
                                  const double* __restrict__ temperature, double* __restrict__ planck) {
  const int g = blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (g >= ng || i >= n) return;
  planck[i * ng + g] = planck_interp(temperature[i], t0, d_t, nlut, ng, g, lut);
}

__global__ void rayleigh_od_kernel(int ncol, int nlay, int ng, const double* __restrict__ pressure_hl,
                                   const double* __restrict__ molar_scat, double* __restrict__ od) {
  const int g = blockIdx.y * blockDim.x + threadIdx.x;  // grid.x runs over rows (no 65535 limit), grid.y over g
  const int64_t cl = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (g >= ng || cl >= (int64_t)ncol * nlay) return;
  const int icol = (int)(cl / nlay), ip = (int)(cl - (int64_t)icol * nlay);
  const double dp = pressure_hl[(int64_t)icol * (nlay + 1) + ip + 1] - pressure_hl[(int64_t)icol * (nlay + 1) + ip];
  // ckd_model.h:246-250: dp * (1/(g*0.001*M)) * sigma
  od[cl * ng + g] = (dp * (1.0 / (ECCKD_ACCEL_GRAVITY * 0.001 * ECCKD_MOLAR_MASS_DRY_AIR))) * molar_scat[g];
}

__global__ void temperature_fl_kernel(int ncol, int nlay, const double* __restrict__ t_hl,
                                      const double* __restrict__ p_hl, double* __restrict__ t_fl) {
  const int64_t cl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cl >= (int64_t)ncol * nlay) return;
  const int icol = (int)(cl / nlay), i = (int)(cl - (int64_t)icol * nlay);
  const double* t = t_hl + (int64_t)icol * (nlay + 1);
  const double* p = p_hl + (int64_t)icol * (nlay + 1);
  t_fl[cl] = (t[i] * p[i] + t[i + 1] * p[i + 1]) / (p[i] + p[i + 1]);  // run_ckd.cpp:119-122
}

// radiative_transfer_lw.cpp:36-59; thread = (column, spectral index), one warp-wide block per
// 32 spectral points of a column so that 32 blocks (1024 independent recurrences, ~64 KB of loads in
// flight) are resident per SM.  The upward sweep re-reads optical depth and Planck function: parking
// the per-layer terms in shared memory instead was measured at half the speed (27 KB per warp leaves
// 8 warps per SM, too few loads in flight), see profiles/README.md.
__global__ void rt_lw_kernel(int ncol, int nlay, int64_t n, const double* __restrict__ planck,
                             const double* __restrict__ od, const double* __restrict__ surf_emissivity,
                             const double* __restrict__ surf_planck, double* __restrict__ flux_dn,
                             double* __restrict__ flux_up) {
  const unsigned nchunk = (unsigned)((n + blockDim.x - 1) / blockDim.x);  // 1-D grid: block = (column, spectral chunk)
  const int icol = (int)(blockIdx.x / nchunk);
  const int64_t i = (int64_t)(blockIdx.x - (unsigned)icol * nchunk) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* pl = planck + (int64_t)icol * (nlay + 1) * n + i;
  const double* tau = od + (int64_t)icol * nlay * n + i;
  double* dn = flux_dn + (int64_t)icol * (nlay + 1) * n + i;
  double* up = flux_up + (int64_t)icol * (nlay + 1) * n + i;
  double f_dn = 0.0;
  dn[0] = 0.0;
  double b_top = pl[0];
#pragma unroll 4
  for (int l = 0; l < nlay; ++l) {
    double e, f;
    lw_layer(tau[(int64_t)l * n], e, f);
    const double b_bot = pl[(int64_t)(l + 1) * n];
    f_dn = f_dn * (1.0 - e) + b_top * (e - f) + b_bot * f;
    dn[(int64_t)(l + 1) * n] = f_dn;
    b_top = b_bot;
  }
  const double es = surf_emissivity[i];
  double f_up = surf_planck[(int64_t)icol * n + i] * es + (1.0 - es) * f_dn;
  up[(int64_t)nlay * n] = f_up;
  double b_bot = b_top;
#pragma unroll 4
  for (int l = nlay - 1; l >= 0; --l) {
    double e, f;
    lw_layer(tau[(int64_t)l * n], e, f);
    const double b_t = pl[(int64_t)l * n];
    f_up = f_up * (1.0 - e) + b_bot * (e - f) + b_t * f;
    up[(int64_t)l * n] = f_up;
    b_bot = b_t;
  }
}

// The same solver with the column staged in shared memory (the path taken when n is a multiple of
// 32 and the arrays are 16-byte aligned).  One warp owns one (column, 32 spectral points) item at
// a time: it pulls the item's optical depths and Planck functions into shared memory with bulk
// asynchronous copies (two copies when n == 32, where a column is contiguous; one 256-byte copy per
// layer row otherwise), walks down writing flux_dn and overwriting each layer's two inputs IN PLACE
// with the two numbers the way up needs - the layer transmittance 1 - e and the source
// B(l+1) (e - f) + B(l) f - and walks up with one FMA per layer.  Every input is read from HBM
// once (32 B per column x layer x g, the algorithmic figure; the register-only kernel above re-reads
// both inputs on the way up: 48 B), and the layer terms are evaluated once instead of twice, four
// layers at a time (they are independent; only the flux recurrence is serial).
// 1 - exp(-1.66 tau) is trans_core.cuh's table evaluation (2.3e-16 absolute error) and the division
// by tau a reciprocal refined by two Newton steps: the fluxes differ from the kernel above by a
// few ulp, nine orders of magnitude inside the 1e-6 W m-2 tolerance.
struct RtLwStagedParams {
  int ncol, nlay;
  int64_t n;
  const double* planck;
  const double* od;
  const double* surf_emissivity;
  const double* surf_planck;
  double* flux_dn;
  double* flux_up;
  double* bb_dn;  // [ncol][nlay+1] sums over the n points, or NULL; only with n == 32 (lane = spectral point)
  double* bb_up;
  TransConst tc;  // kd = LW_DIFFUSIVITY
};

__device__ __forceinline__ double warp_sum_xor(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void lw_layer_fast(double tau, const TransConst& K, const TransLane& tl, double& e, double& f) {
  e = trans_one_minus_exp(tau, K, tl);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(tau));  // ~20 bits
  r = fma(r, fma(-tau, r, 1.0), r);
  r = fma(r, fma(-tau, r, 1.0), r);
  f = (e > 1.0e-5) ? (1.0 - (e * (1.0 / ECCKD_LW_DIFFUSIVITY)) * r) : (0.5 * e);
}

template <bool UNIT>
__global__ void __launch_bounds__(32) rt_lw_staged_kernel(const __grid_constant__ RtLwStagedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int nlay = p.nlay;
  const int64_t n = p.n;
  double* s_tau = reinterpret_cast<double*>(smem_raw);  // [nlay][32]  -> 1 - e
  double* s_pl = s_tau + (size_t)nlay * 32;             // [nlay+1][32] -> source of the upward sweep in rows 0..nlay-1
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned bar = sbase + (unsigned)((2 * nlay + 1) * 32 * sizeof(double));
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const TransLane tl = trans_lane_init();
  const int64_t nchunk = n >> 5;
  const int64_t nitems = (int64_t)p.ncol * nchunk;
  unsigned parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t icol = item / nchunk;
    const int64_t i0 = (item - icol * nchunk) << 5;
    const double* g_tau = p.od + icol * nlay * n + i0;
    const double* g_pl = p.planck + icol * (nlay + 1) * n + i0;
    if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)((2 * nlay + 1) * 256));
    __syncwarp();
    if (UNIT) {
      if (lane == 0) bulk_g2s(sbase, g_tau, (unsigned)(nlay * 256), bar);
      if (lane == 1) bulk_g2s(sbase + (unsigned)(nlay * 256), g_pl, (unsigned)((nlay + 1) * 256), bar);
    } else {
      for (int r = lane; r < 2 * nlay + 1; r += 32) {
        const double* src = r < nlay ? g_tau + (int64_t)r * n : g_pl + (int64_t)(r - nlay) * n;
        bulk_g2s(sbase + (unsigned)(r * 256), src, 256u, bar);
      }
    }
    double* dn = p.flux_dn + icol * (nlay + 1) * n + i0 + lane;
    double* up = p.flux_up + icol * (nlay + 1) * n + i0 + lane;
    const double es = p.surf_emissivity[i0 + lane];
    const double bs = p.surf_planck[icol * n + i0 + lane];
    mbar_wait(bar, parity);
    parity ^= 1u;
    // broadband sums (run_ckd.cpp:355-356) while the fluxes are in registers: with n == 32 the lanes of
    // the warp are the g points of the column (fixed order: xor shuffle tree)
    const bool bb = UNIT && p.bb_dn != nullptr;
    double* bdn = bb ? p.bb_dn + icol * (nlay + 1) : nullptr;
    double* bup = bb ? p.bb_up + icol * (nlay + 1) : nullptr;
    double f_dn = 0.0;
    dn[0] = 0.0;
    if (bb && lane == 0) bdn[0] = 0.0;
    double b_top = s_pl[lane];
    int l = 0;
    const int slot4 = warp_sum4_slot(lane);
    for (; l + 4 <= nlay; l += 4) {
      double e[4], f[4], b[4], fd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lw_layer_fast(s_tau[(l + j) * 32 + lane], p.tc, tl, e[j], f[j]);
        b[j] = s_pl[(l + j + 1) * 32 + lane];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f_dn = f_dn * (1.0 - e[j]) + b_top * (e[j] - f[j]) + b[j] * f[j];
        dn[(int64_t)(l + j + 1) * n] = f_dn;
        fd[j] = f_dn;
        s_tau[(l + j) * 32 + lane] = 1.0 - e[j];
        s_pl[(l + j) * 32 + lane] = b[j] * (e[j] - f[j]) + b_top * f[j];
        b_top = b[j];
      }
      if (bb) {  // the four levels' broadband sums together
        const double sdn = warp_sum4_transposed(fd, lane);
        if ((lane & 7) == 0) bdn[l + 1 + slot4] = sdn;
      }
    }
    for (; l < nlay; ++l) {
      double e, f;
      lw_layer_fast(s_tau[l * 32 + lane], p.tc, tl, e, f);
      const double b_bot = s_pl[(l + 1) * 32 + lane];
      f_dn = f_dn * (1.0 - e) + b_top * (e - f) + b_bot * f;
      dn[(int64_t)(l + 1) * n] = f_dn;
      if (bb) {
        const double sdn = warp_sum_xor(f_dn);
        if (lane == 0) bdn[l + 1] = sdn;
      }
      s_tau[l * 32 + lane] = 1.0 - e;
      s_pl[l * 32 + lane] = b_bot * (e - f) + b_top * f;
      b_top = b_bot;
    }
    double f_up = bs * es + (1.0 - es) * f_dn;
    up[(int64_t)nlay * n] = f_up;
    if (bb) {
      const double sup = warp_sum_xor(f_up);
      if (lane == 0) bup[nlay] = sup;
    }
    l = nlay - 1;
    for (; l >= 3; l -= 4) {  // four levels at a time: one transposed reduction for their broadband sums
      double fu[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f_up = f_up * s_tau[(l - j) * 32 + lane] + s_pl[(l - j) * 32 + lane];
        up[(int64_t)(l - j) * n] = f_up;
        fu[j] = f_up;
      }
      if (bb) {
        const double sup = warp_sum4_transposed(fu, lane);
        if ((lane & 7) == 0) bup[l - slot4] = sup;
      }
    }
    for (; l >= 0; --l) {
      f_up = f_up * s_tau[l * 32 + lane] + s_pl[l * 32 + lane];
      up[(int64_t)l * n] = f_up;
      if (bb) {
        const double sup = warp_sum_xor(f_up);
        if (lane == 0) bup[l] = sup;
      }
    }
    // the next item's bulk copies (async proxy) overwrite what this one wrote through the generic proxy
    fence_proxy_async();
    __syncwarp();
  }
}

// radiative_transfer_sw.cpp:36-42 and :61-76
__global__ void rt_sw_kernel(int ncol, int nlay, int64_t n, const double* __restrict__ cos_sza,
                             const double* __restrict__ ssi, const double* __restrict__ od,
                             const double* __restrict__ albedo, double* __restrict__ flux_dn,
                             double* __restrict__ flux_up) {
  const unsigned nchunk = (unsigned)((n + blockDim.x - 1) / blockDim.x);
  const int icol = (int)(blockIdx.x / nchunk);
  const int64_t i = (int64_t)(blockIdx.x - (unsigned)icol * nchunk) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mu0 = cos_sza[icol];
  const double minus_sec_sza = -1.0 / mu0;
  const double* tau = od + (int64_t)icol * nlay * n + i;
  double* dn = flux_dn + (int64_t)icol * (nlay + 1) * n + i;
  double f = mu0 * ssi[i];
  dn[0] = f;
#pragma unroll 4
  for (int l = 0; l < nlay; ++l) {
    f = f * exp(minus_sec_sza * tau[(int64_t)l * n]);
    dn[(int64_t)(l + 1) * n] = f;
  }
  if (flux_up) {
    double* up = flux_up + (int64_t)icol * (nlay + 1) * n + i;
    f = f * albedo[i];
    up[(int64_t)nlay * n] = f;
#pragma unroll 4
    for (int l = nlay - 1; l >= 0; --l) {
      f = f * exp(-2.0 * tau[(int64_t)l * n]);
      up[(int64_t)l * n] = f;
    }
  }
}

// The shortwave solvers with the column's optical depths staged in shared memory, like
// rt_lw_staged_kernel: tau is read from HBM once (24 B per column x layer x g with the upward
// sweep instead of 32), the two exponentials of a layer are evaluated together on the way down
// (exp(-tau/mu0) for the direct beam, exp(-2 tau) parked in place of tau for the way up) with
// the table exponential of trans_core.cuh (exp_table: below 1 ulp relative, so that the deep,
// tiny direct fluxes keep their relative accuracy), four layers at a time.
struct RtSwStagedParams {
  int ncol, nlay;
  int64_t n;
  const double* cos_sza;  // [ncol]
  const double* ssi;      // [n]
  const double* od;
  const double* albedo;   // [n] or NULL (direct beam only)
  double* flux_dn;
  double* flux_up;
};

__global__ void __launch_bounds__(32) rt_sw_staged_kernel(const __grid_constant__ RtSwStagedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int nlay = p.nlay;
  const int64_t n = p.n;
  double* s_tau = reinterpret_cast<double*>(smem_raw);  // [nlay][32] -> exp(-2 tau)
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned bar = sbase + (unsigned)(nlay * 32 * sizeof(double));
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const TransLane tl = trans_lane_init();
  const int64_t nchunk = n >> 5;
  const int64_t nitems = (int64_t)p.ncol * nchunk;
  const bool contiguous = (n == 32);
  unsigned parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t icol = item / nchunk;
    const int64_t i0 = (item - icol * nchunk) << 5;
    const double* g_tau = p.od + icol * nlay * n + i0;
    if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)(nlay * 256));
    __syncwarp();
    if (contiguous) {
      if (lane == 0) bulk_g2s(sbase, g_tau, (unsigned)(nlay * 256), bar);
    } else {
      for (int r = lane; r < nlay; r += 32) bulk_g2s(sbase + (unsigned)(r * 256), g_tau + (int64_t)r * n, 256u, bar);
    }
    double* dn = p.flux_dn + icol * (nlay + 1) * n + i0 + lane;
    const double mu0 = p.cos_sza[icol];
    const double minus_sec_sza = -1.0 / mu0;
    double f = mu0 * p.ssi[i0 + lane];
    const double alb = p.albedo ? p.albedo[i0 + lane] : 0.0;
    dn[0] = f;
    mbar_wait(bar, parity);
    parity ^= 1u;
    int l = 0;
    for (; l + 4 <= nlay; l += 4) {
      double t[4], u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double tau = s_tau[(l + j) * 32 + lane];
        // arguments below -700 give exp() < 1e-304: clamp (exp_table has no underflow handling)
        t[j] = exp_table(fmax(minus_sec_sza * tau, -700.0), tl);
        u[j] = p.albedo ? exp_table(fmax(-2.0 * tau, -700.0), tl) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f = f * t[j];
        dn[(int64_t)(l + j + 1) * n] = f;
        s_tau[(l + j) * 32 + lane] = u[j];
      }
    }
    for (; l < nlay; ++l) {
      const double tau = s_tau[l * 32 + lane];
      f = f * exp_table(fmax(minus_sec_sza * tau, -700.0), tl);
      dn[(int64_t)(l + 1) * n] = f;
      s_tau[l * 32 + lane] = p.albedo ? exp_table(fmax(-2.0 * tau, -700.0), tl) : 0.0;
    }
    if (p.albedo) {
      double* up = p.flux_up + icol * (nlay + 1) * n + i0 + lane;
      f = f * alb;
      up[(int64_t)nlay * n] = f;
#pragma unroll 6
      for (l = nlay - 1; l >= 0; --l) {
        f = f * s_tau[l * 32 + lane];
        up[(int64_t)l * n] = f;
      }
    }
    fence_proxy_async();  // the next item's bulk copies overwrite what was written through the generic proxy
    __syncwarp();
  }
}

// heating_rate.h:38-48
__global__ void heating_rate_kernel(int ncol, int nlay, int64_t n, const double* __restrict__ pressure_hl,
                                    const double* __restrict__ flux_dn, const double* __restrict__ flux_up,
                                    double* __restrict__ hr) {
  const unsigned nchunk = (unsigned)((n + blockDim.x - 1) / blockDim.x);
  const int64_t cl = blockIdx.x / nchunk;
  const int64_t i = (int64_t)(blockIdx.x - (unsigned)cl * nchunk) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int icol = (int)(cl / nlay), l = (int)(cl - (int64_t)icol * nlay);
  const double* p = pressure_hl + (int64_t)icol * (nlay + 1);
  const double conv = -(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (p[l + 1] - p[l]);
  const int64_t a = ((int64_t)icol * (nlay + 1) + l) * n + i;
  double v;
  if (flux_up) v = conv * (flux_dn[a + n] - flux_dn[a] - flux_up[a + n] + flux_up[a]);
  else v = conv * (flux_dn[a + n] - flux_dn[a]);
  hr[cl * n + i] = v;
}

// sum(flux, 2): run_ckd.cpp:355-356,367.  One warp per (column, level) row.
__global__ void sum_g_kernel(int64_t nrows, int ng, const double* __restrict__ flux, double* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  // fixed order: lane partial sums over g = lane, lane+32, ... then a shuffle tree
  double s = 0.0;
  for (int g = lane; g < ng; g += 32) s += flux[row * ng + g];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// The same sums for even ng and 16-byte aligned rows: eight lanes per row, four consecutive rows per
// warp (one contiguous KB at ng = 32), 16-byte loads, two row quads in flight per warp.  Fixed order:
// lane j of a row adds g = 2j, 2j+1, 2j+16, 2j+17, ... in that order, then a three-step shuffle tree.
// (One lane per g point and one row per warp, as above, ran at a third of the HBM rate.)
__global__ void __launch_bounds__(256) sum_g8_kernel(int64_t nrows, int ng, const double* __restrict__ flux,
                                                     double* __restrict__ out) {
  const int lane = threadIdx.x & 31, sub = lane & 7, rq = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp * 8; r0 < nrows; r0 += nwarps * 8) {
    double s[2] = {0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t row = r0 + h * 4 + rq;
      if (row < nrows) {
        const double2* src = reinterpret_cast<const double2*>(flux + row * ng);
        for (int g2 = sub; g2 < (ng >> 1); g2 += 8) {
          const double2 v = src[g2];
          s[h] += v.x;
          s[h] += v.y;
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
      const int64_t row = r0 + h * 4 + rq;
      if (sub == 0 && row < nrows) out[row] = s[h];
    }
  }
}

static void launch_sum_g(ecckd_ctx* ctx, int64_t nrows, int ng, const double* flux, double* out) {
  if ((ng & 1) == 0 && (reinterpret_cast<uintptr_t>(flux) & 15u) == 0) {
    const int64_t want = (nrows + 63) / 64;  // 8 warps x 8 rows per block and trip
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->num_sms * 16));
    sum_g8_kernel<<<blocks, 256, 0, ctx->stream>>>(nrows, ng, flux, out);
  } else {
    sum_g_kernel<<<(unsigned)((nrows * 32 + 255) / 256), 256, 0, ctx->stream>>>(nrows, ng, flux, out);
  }
}

// Fused run_ckd core (run_ckd.cpp:119-132,261-367): block = one column,
// thread = g point.  Interpolation state per (layer, gas) is computed once per
// column into shared memory, then each thread walks the layers.
struct RunCkdParams {
  const GasDev* gases;
  int ngas, nt, np, ng, nlut, is_sw;
  double log_p_0, d_log_p, d_t, planck_t0, planck_dt, tsi_scaling;
  const double* temperature;       // [nt][np]
  const double* planck_lut;        // [nlut][ng]
  const double* solar_irradiance;  // [ng]
  const double* rayleigh;          // [ng]
  int ncol, nlay;
  const double* pressure_hl;
  const double* temperature_hl;
  double* od;
  double* planck_hl;
  double* planck_surf;
  double* flux_dn;
  double* flux_up;
  double* bb_flux_dn;
  double* bb_flux_up;
  double* scratch_od;  // [ncol][nlay][ng] when od == NULL
  int vec2;            // ng even and the optical-depth array 16-byte aligned: two g points per thread
};

// g-independent part of one gas's optical depth at one (column, layer)
struct GasLayer {
  double weight;   // what multiplies the interpolated absorption (ckd_model.cpp:980-999)
  double cindex0;  // LUT gases: clamped concentration index (ic0 + cw1); -1 otherwise
};
constexpr int RUN_CKD_MAX_GAS = 16;
__host__ __device__ inline size_t run_ckd_col_smem(int nlay, int ngas) {
  return (((size_t)nlay * sizeof(Interp) + (size_t)nlay * ngas * sizeof(GasLayer)) + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(1024, 1) run_ckd_kernel(const RunCkdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ const double* s_molar_abs[RUN_CKD_MAX_GAS];
  if (threadIdx.y == 0 && threadIdx.z == 0 && threadIdx.x < p.ngas) s_molar_abs[threadIdx.x] = p.gases[threadIdx.x].molar_abs;
  // blockDim.y columns per block: neighbouring columns walk the layers at the same pace and read the
  // same LUT rows (same pressure index, a handful of temperature indices), so the gathers of the
  // block's warps hit in L1 instead of each going to L2.  A column past the end repeats the last one.
  // blockDim.z warps share a column and take every blockDim.z-th layer: the per-column interpolation
  // state in shared memory caps the resident columns at ~20 per SM, and one warp per column left the
  // gathers latency-bound.
  const int icol = min((int)(blockIdx.x * blockDim.y + threadIdx.y), p.ncol - 1);
  const int nlay = p.nlay, ng = p.ng, ngas = p.ngas;
  const size_t col_bytes = run_ckd_col_smem(nlay, ngas);
  unsigned char* my = smem_raw + threadIdx.y * col_bytes;
  Interp* s_it = reinterpret_cast<Interp*>(my);                                    // [nlay]
  GasLayer* s_gl = reinterpret_cast<GasLayer*>(s_it + nlay);                       // [nlay][ngas]
  const double* p_hl = p.pressure_hl + (int64_t)icol * (nlay + 1);
  const double* t_hl = p.temperature_hl + (int64_t)icol * (nlay + 1);
  // everything of ckd_model.cpp:962-1009 that does not depend on the g point, once per (layer, gas):
  // interpolation indices and weights, the concentration weight, and for a look-up-table gas its
  // concentration index (one log() per layer instead of one per layer and g point)
  for (int l = threadIdx.z * blockDim.x + threadIdx.x; l < nlay; l += blockDim.x * blockDim.z) {
    const double t_fl = (t_hl[l] * p_hl[l] + t_hl[l + 1] * p_hl[l + 1]) / (p_hl[l] + p_hl[l + 1]);
    const Interp it = make_interp(p_hl[l], p_hl[l + 1], t_fl, p.log_p_0, p.d_log_p, p.d_t, p.nt, p.np, p.temperature);
    s_it[l] = it;
    for (int ig = 0; ig < ngas; ++ig) {
      const GasDev gs = p.gases[ig];
      const bool have = gs.vmr != nullptr;
      const double vmr = have ? gs.vmr[(int64_t)icol * nlay + l] : 0.0;
      double weight = 0.0;
      if (have) {
        weight = (gs.conc_dependence == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - gs.reference_vmr)
                                                                    : it.simple_weight * vmr;
      }
      GasLayer gl;
      gl.cindex0 = -1.0;
      if (gs.conc_dependence == ECCKD_CONC_LUT) {
        double cindex0 = (log(vmr) - gs.log_vmr0) / gs.d_log_c;
        cindex0 = fmax(0.0, fmin(cindex0, gs.nconc - 1.0001));
        gl.cindex0 = cindex0;
        gl.weight = weight;
      } else {
        gl.weight = (gs.conc_dependence == ECCKD_CONC_NONE) ? it.simple_weight : weight;
      }
      s_gl[l * ngas + ig] = gl;
    }
  }
  __syncthreads();

  double* od_col = (p.od ? p.od : p.scratch_od) + (int64_t)icol * nlay * ng;
  const size_t sP = (size_t)ng, sT = (size_t)p.np * ng, sC = (size_t)p.nt * sT;
  auto bilinear = [&](const Interp& it, double k00, double k01, double k10, double k11) {
    return it.tw0 * (it.pw0 * k00 + it.pw1 * k01) + it.tw1 * (it.pw0 * k10 + it.pw1 * k11);  // as gas_od()
  };
  if (p.vec2) {
    // Two g points per thread (16-byte loads: one request fetches two LUT rows' worth for a warp of
    // ng = 32), the (layer, g pair) items of the column dealt out over its threads; every layer is
    // independent.  Same arithmetic per element as the scalar loop below.
    const int hg = ng >> 1;
    const int tid_c = threadIdx.z * blockDim.x + threadIdx.x, nthr_c = blockDim.x * blockDim.z;
    for (int t = tid_c; t < nlay * hg; t += nthr_c) {
      const int l = t / hg, g = (t - l * hg) * 2;
      double tx = 0.0, ty = 0.0;
      const Interp it = s_it[l];
      const size_t o00 = it.it0 * sT + it.ip0 * sP + g;
      for (int ig = 0; ig < ngas; ++ig) {
        const GasLayer gl = s_gl[l * ngas + ig];
        const int ic0 = (int)gl.cindex0;  // -1 for the gases without a concentration axis
        const double* k = s_molar_abs[ig] + o00 + (ic0 >= 0 ? ic0 * sC : (size_t)0);
        const double2 k00 = *reinterpret_cast<const double2*>(k), k01 = *reinterpret_cast<const double2*>(k + sP),
                      k10 = *reinterpret_cast<const double2*>(k + sT), k11 = *reinterpret_cast<const double2*>(k + sT + sP);
        const double ax = bilinear(it, k00.x, k01.x, k10.x, k11.x), ay = bilinear(it, k00.y, k01.y, k10.y, k11.y);
        if (ic0 >= 0) {
          const double cw1 = gl.cindex0 - ic0;
          k += sC;
          const double2 c00 = *reinterpret_cast<const double2*>(k), c01 = *reinterpret_cast<const double2*>(k + sP),
                        c10 = *reinterpret_cast<const double2*>(k + sT), c11 = *reinterpret_cast<const double2*>(k + sT + sP);
          tx += gl.weight * ((1.0 - cw1) * ax + cw1 * bilinear(it, c00.x, c01.x, c10.x, c11.x));
          ty += gl.weight * ((1.0 - cw1) * ay + cw1 * bilinear(it, c00.y, c01.y, c10.y, c11.y));
        } else {
          tx += gl.weight * ax;
          ty += gl.weight * ay;
        }
      }
      *reinterpret_cast<double2*>(od_col + (int64_t)l * ng + g) = make_double2(fmax(tx, 0.0), fmax(ty, 0.0));
    }
  } else
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    // optical depth: sum over gases, clamp (run_ckd.cpp:265-317).  (Keeping the table values at ip0 + 1
    // in registers for the next layer was measured slower: 7.1 vs 6.3 ms for 65536 x 54 x 32.)
    for (int l = threadIdx.z; l < nlay; l += blockDim.z) {
      double tau = 0.0;
      const Interp it = s_it[l];
      const size_t o00 = it.it0 * sT + it.ip0 * sP + g;
      for (int ig = 0; ig < ngas; ++ig) {
        const GasLayer gl = s_gl[l * ngas + ig];
        const int ic0 = (int)gl.cindex0;  // -1 for the gases without a concentration axis
        const double* k = s_molar_abs[ig] + o00 + (ic0 >= 0 ? ic0 * sC : (size_t)0);
        const double a = bilinear(it, k[0], k[sP], k[sT], k[sT + sP]);
        if (ic0 >= 0) {
          const double cw1 = gl.cindex0 - ic0;
          k += sC;
          tau += gl.weight * ((1.0 - cw1) * a + cw1 * bilinear(it, k[0], k[sP], k[sT], k[sT + sP]));
        } else {
          tau += gl.weight * a;
        }
      }
      od_col[(int64_t)l * ng + g] = fmax(tau, 0.0);
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    const int64_t hl0 = (int64_t)icol * (nlay + 1) * ng + g;
    if (!p.is_sw) {
      // Planck function at half levels and surface (run_ckd.cpp:131-132)
      if (!(p.flux_dn && p.flux_up)) {  // no sweep here: the levels are shared out over the column's warps
        if (p.planck_hl) {
          for (int l = threadIdx.z; l <= nlay; l += blockDim.z)
            p.planck_hl[hl0 + (int64_t)l * ng] = planck_interp(t_hl[l], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
        }
        if (p.planck_surf && threadIdx.z == 0)
          p.planck_surf[(int64_t)icol * ng + g] = planck_interp(t_hl[nlay], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
        continue;
      }
      if (threadIdx.z != 0) continue;  // the sweeps below are serial in the layers: one warp per column
      double b_top = planck_interp(t_hl[0], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
      if (p.planck_hl) p.planck_hl[hl0] = b_top;
      const double b_surf = planck_interp(t_hl[nlay], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
      if (p.planck_surf) p.planck_surf[(int64_t)icol * ng + g] = b_surf;
      double f_dn = 0.0;
      p.flux_dn[hl0] = 0.0;
      for (int l = 0; l < nlay; ++l) {
        double e, f;
        lw_layer(od_col[(int64_t)l * ng + g], e, f);
        const double b_bot = planck_interp(t_hl[l + 1], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
        if (p.planck_hl) p.planck_hl[hl0 + (int64_t)(l + 1) * ng] = b_bot;
        f_dn = f_dn * (1.0 - e) + b_top * (e - f) + b_bot * f;
        p.flux_dn[hl0 + (int64_t)(l + 1) * ng] = f_dn;
        b_top = b_bot;
      }
      // surface emissivity 1 (run_ckd.cpp:344-345): up = B_s*1 + (1-1)*dn
      double f_up = b_surf * 1.0 + (1.0 - 1.0) * f_dn;
      p.flux_up[hl0 + (int64_t)nlay * ng] = f_up;
      double b_bot = b_top;
      for (int l = nlay - 1; l >= 0; --l) {
        double e, f;
        lw_layer(od_col[(int64_t)l * ng + g], e, f);
        const double b_t = planck_interp(t_hl[l], p.planck_t0, p.planck_dt, p.nlut, ng, g, p.planck_lut);
        f_up = f_up * (1.0 - e) + b_bot * (e - f) + b_t * f;
        p.flux_up[hl0 + (int64_t)l * ng] = f_up;
        b_bot = b_t;
      }
    } else if (p.flux_dn && threadIdx.z == 0) {
      // direct beam at mu0 = 0.5 through gas + Rayleigh (run_ckd.cpp:360-365)
      const double mu0 = 0.5;
      const double minus_sec = -1.0 / mu0;
      double f = mu0 * (p.tsi_scaling * p.solar_irradiance[g]);
      p.flux_dn[hl0] = f;
      const double sig = p.rayleigh[g];
      for (int l = 0; l < nlay; ++l) {
        const double ray = ((p_hl[l + 1] - p_hl[l]) * (1.0 / (ECCKD_ACCEL_GRAVITY * 0.001 * ECCKD_MOLAR_MASS_DRY_AIR))) * sig;
        f = f * exp(minus_sec * (od_col[(int64_t)l * ng + g] + ray));
        p.flux_dn[hl0 + (int64_t)(l + 1) * ng] = f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Total optical depth and Planck function of the run_ckd core (run_ckd.cpp:119-132, 265-317;
// ckd_model.cpp:962-1088, 1105-1145), layer-major with the look-up-table rows held in REGISTERS.
//
// The interpolation indices (ip0, it0) of ckd_model.cpp:962-980 belong to the (column, layer), not to
// the gas: every gas reads the same four corners of its own table.  At one layer the columns of a tile
// share a handful of (ip0, it0) pairs (pressure at a model level varies little between columns,
// temperature spans one or two 20 K table intervals), so a warp (lane = g point) takes one layer of a
// tile of 32 columns, groups the columns by (ip0, it0) with match.any, loads the 4 corner rows of every
// gas without a concentration axis ONCE per group (coalesced row reads into registers) and then
// produces the optical depth of each column of the group with 5 FMAs per gas and no table read at all.
// The gas with a concentration axis (H2O: ckd_model.cpp:1001-1045) reads its 8 rows per column, the
// next column's rows in flight while this one is summed.  The column-major kernel above gathers 36 rows
// per (column, layer); this one 4*NLIN per group + 8 per column.  Phase A (lane = column) computes
// everything that does not depend on g once per (column, layer).  The host orders the gases: the NLIN
// table gases first, then the ones with a concentration axis (the sum over gases is reordered
// accordingly; run_ckd.cpp:312 adds them in file order).
constexpr int TOD_MAXGAS = 8;
constexpr int TOD_TILE = 32;

struct __align__(16) TodCol {  // g-independent state of one (column, layer), in shared memory per warp
  double c00, c01, c10, c11;   // tw0*pw0, tw0*pw1, tw1*pw0, tw1*pw1
  double w[TOD_MAXGAS];        // what multiplies the interpolated absorption of each gas
  double cw1[TOD_MAXGAS];      // gases with a concentration axis: weight of level ic0 + 1 ...
  int ic0[TOD_MAXGAS];         // ... and the level index
};

struct TotalOdParams {
  const double* packed[TOD_MAXGAS];  // corner-packed tables (ecckd_ckd_model::d_packed)
  const double* vmr[TOD_MAXGAS];
  double reference_vmr[TOD_MAXGAS], log_vmr0[TOD_MAXGAS], d_log_c[TOD_MAXGAS];
  int conc_dependence[TOD_MAXGAS], nconc[TOD_MAXGAS];
  int ngas, nlin, nt, np, ng, nlut;
  double log_p_0, d_log_p, d_t, planck_t0, planck_dt;
  const double* temperature;  // [nt][np]
  const double* planck_lut;   // [nlut][ng]
  int ncol, nlay;
  const double* pressure_hl;
  const double* temperature_hl;
  double* od;           // [ncol][nlay][ng]
  double* planck_hl;    // [ncol][nlay+1][ng] or NULL
  double* planck_surf;  // [ncol][ng] or NULL
};

template <int NLIN, bool HAS_LUT, int NWARP, int MINB>
__global__ void __launch_bounds__(NWARP * 32, MINB) total_od_planck_kernel(const __grid_constant__ TotalOdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  TodCol* s_col = reinterpret_cast<TodCol*>(smem_raw) + warp * TOD_TILE;
  const int nlay = p.nlay, ng = p.ng, ngas = p.ngas;
  const int ntile = (p.ncol + TOD_TILE - 1) / TOD_TILE;
  const int64_t n_od = (int64_t)ntile * nlay;
  const int64_t n_units = n_od + (p.planck_hl || p.planck_surf ? (int64_t)ntile * (nlay + 1) : 0);
  // strides of the corner-packed tables, in elements (< 2^31)
  const unsigned qP = (unsigned)ng * 4, qT = (unsigned)(p.np - 1) * qP, qC = (unsigned)(p.nt - 1) * qT;
  for (int64_t u = (int64_t)blockIdx.x * nwarp + warp; u < n_units; u += (int64_t)gridDim.x * nwarp) {
    if (u >= n_od) {
      // ---- Planck function at one half level of the tile's columns (run_ckd.cpp:131-132;
      // ckd_model.cpp:1124-1142): index and weights per column by lane = column, rows by lane = g
      const int64_t v = u - n_od;
      const int tile = (int)(v / (nlay + 1)), lev = (int)(v - (int64_t)tile * (nlay + 1));
      const int c0 = tile * TOD_TILE, nc = min(TOD_TILE, p.ncol - c0);
      int my_off = 0;
      double my_w0 = 0.0, my_w1 = 0.0;
      if (lane < nc) {
        const double t = p.temperature_hl[(int64_t)(c0 + lane) * (nlay + 1) + lev];
        const double tindex0 = (t - p.planck_t0) / p.planck_dt;
        if (tindex0 >= 0) {
          const int it0 = min((int)tindex0, p.nlut - 2);
          my_off = it0 * ng;
          my_w1 = tindex0 - it0;
          my_w0 = 1.0 - my_w1;
        } else {  // below the table: (t / t0) * lut[0]
          my_w0 = t / p.planck_t0;
        }
      }
      const int64_t lev_stride = (int64_t)(nlay + 1) * ng;
      for (int g0 = 0; g0 < ng; g0 += 32) {
        const int g = min(g0 + lane, ng - 1);
        const bool store = g0 + lane < ng;
        const double* lut_g = p.planck_lut + g;
        double* out = p.planck_hl ? p.planck_hl + ((int64_t)c0 * (nlay + 1) + lev) * ng + g : nullptr;
        double* surf = (lev == nlay && p.planck_surf) ? p.planck_surf + (int64_t)c0 * ng + g : nullptr;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) {
          const int off = __shfl_sync(0xffffffffu, my_off, c);
          const double w0 = __shfl_sync(0xffffffffu, my_w0, c), w1 = __shfl_sync(0xffffffffu, my_w1, c);
          const double b = w0 * __ldg(lut_g + off) + w1 * __ldg(lut_g + off + ng);  // w1 == 0 below the table
          if (store && out) out[c * lev_stride] = b;
          if (store && surf) surf[c * ng] = b;
        }
      }
      continue;
    }
    // consecutive warps take consecutive layers of one tile: its columns' profile sectors are shared in L1/L2
    const int tile = (int)(u / nlay), l = (int)(u - (int64_t)tile * nlay);
    const int c0 = tile * TOD_TILE;
    // ---- phase A: lane = column
    int key = -1 - lane;  // columns past the end form groups of their own that nobody visits
    const bool valid = c0 + lane < p.ncol;
    if (valid) {
      const int64_t col = c0 + lane;
      const double* p_hl = p.pressure_hl + col * (nlay + 1);
      const double* t_hl = p.temperature_hl + col * (nlay + 1);
      const double p0 = p_hl[l], p1 = p_hl[l + 1];
      const double t_fl = (t_hl[l] * p0 + t_hl[l + 1] * p1) / (p0 + p1);
      const Interp it = make_interp(p0, p1, t_fl, p.log_p_0, p.d_log_p, p.d_t, p.nt, p.np, p.temperature);
      key = it.it0 * p.np + it.ip0;
      TodCol tc;
      tc.c00 = it.tw0 * it.pw0;
      tc.c01 = it.tw0 * it.pw1;
      tc.c10 = it.tw1 * it.pw0;
      tc.c11 = it.tw1 * it.pw1;
#pragma unroll
      for (int ig = 0; ig < TOD_MAXGAS; ++ig) {
        tc.w[ig] = 0.0;
        tc.cw1[ig] = 0.0;
        tc.ic0[ig] = 0;
        if (ig >= ngas) continue;
        const int cd = p.conc_dependence[ig];
        const bool have = p.vmr[ig] != nullptr;
        const double vmr = have ? p.vmr[ig][col * nlay + l] : 0.0;
        double weight = 0.0;  // ckd_model.cpp:980-999
        if (have) weight = (cd == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - p.reference_vmr[ig]) : it.simple_weight * vmr;
        if (cd == ECCKD_CONC_LUT) {
          double cindex0 = (log(vmr) - p.log_vmr0[ig]) / p.d_log_c[ig];  // :1003-1009
          cindex0 = fmax(0.0, fmin(cindex0, p.nconc[ig] - 1.0001));
          tc.ic0[ig] = (int)cindex0;
          tc.cw1[ig] = cindex0 - tc.ic0[ig];
          tc.w[ig] = weight;
        } else {
          tc.w[ig] = (cd == ECCKD_CONC_NONE) ? it.simple_weight : weight;
        }
      }
      s_col[lane] = tc;
    }
    const unsigned same = __match_any_sync(0xffffffffu, key);
    __syncwarp();
    // ---- phase B: lane = g point, groups of columns with the same four table corners
    const unsigned col_stride = (unsigned)nlay * ng;  // 32 * nlay * ng < 2^31 (checked by the host)
    for (int g0 = 0; g0 < ng; g0 += 32) {
      const int g = min(g0 + lane, ng - 1);
      const bool store = g0 + lane < ng;
      double* const od_unit = p.od + ((int64_t)c0 * nlay + l) * ng + g;
      unsigned todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        unsigned members = __shfl_sync(0xffffffffu, same, leader);
        const int gkey = __shfl_sync(0xffffffffu, key, leader);
        todo &= ~members;
        const int it0 = gkey / p.np, ip0 = gkey - it0 * p.np;
        const unsigned o00 = (unsigned)it0 * qT + (unsigned)ip0 * qP + (unsigned)g * 4;
        double K[NLIN > 0 ? NLIN : 1][4];
#pragma unroll
        for (int ig = 0; ig < NLIN; ++ig) {
          const double2* k = reinterpret_cast<const double2*>(p.packed[ig] + o00);
          const double2 a = __ldg(k), b = __ldg(k + 1);
          K[ig][0] = a.x;
          K[ig][1] = a.y;
          K[ig][2] = b.x;
          K[ig][3] = b.y;
        }
        const double* lut_base = HAS_LUT ? p.packed[NLIN] + o00 : nullptr;
        auto lut_rows = [&](int c, double2 (&L)[4]) {
          if (!HAS_LUT) return;
          const double2* k = reinterpret_cast<const double2*>(lut_base + (size_t)((unsigned)s_col[c].ic0[NLIN] * qC));
          const double2* k1 = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(k) + qC);
          L[0] = __ldg(k);
          L[1] = __ldg(k + 1);
          L[2] = __ldg(k1);
          L[3] = __ldg(k1 + 1);
        };
        auto column = [&](int c, const double2 (&L)[4]) {
          const TodCol& tc = s_col[c];
          const double c00 = tc.c00, c01 = tc.c01, c10 = tc.c10, c11 = tc.c11;
          double tau0 = 0.0, tau1 = 0.0;  // two chains
#pragma unroll
          for (int ig = 0; ig < NLIN; ++ig) {
            const double v = c00 * K[ig][0] + c01 * K[ig][1] + c10 * K[ig][2] + c11 * K[ig][3];
            if (ig & 1) tau1 += tc.w[ig] * v; else tau0 += tc.w[ig] * v;
          }
          if (HAS_LUT) {
            const double cw1 = tc.cw1[NLIN];
            const double a = c00 * L[0].x + c01 * L[0].y + c10 * L[1].x + c11 * L[1].y;
            const double b = c00 * L[2].x + c01 * L[2].y + c10 * L[3].x + c11 * L[3].y;
            tau1 += tc.w[NLIN] * ((1.0 - cw1) * a + cw1 * b);
          }
          double tau = tau0 + tau1;
          for (int ig = NLIN + (HAS_LUT ? 1 : 0); ig < ngas; ++ig) {  // further gases with a concentration axis
            const double cw1 = tc.cw1[ig];
            const double* k = p.packed[ig] + (size_t)((unsigned)tc.ic0[ig] * qC) + o00;
            const double a = c00 * k[0] + c01 * k[1] + c10 * k[2] + c11 * k[3];
            k += qC;
            const double b = c00 * k[0] + c01 * k[1] + c10 * k[2] + c11 * k[3];
            tau += tc.w[ig] * ((1.0 - cw1) * a + cw1 * b);
          }
          if (store) od_unit[(unsigned)c * col_stride] = fmax(tau, 0.0);  // run_ckd.cpp:317
        };
        // software pipeline over the members: the next column's LUT rows are loaded before this one is summed
        double2 La[4], Lb[4];
        int ca = __ffs(members) - 1;
        members &= members - 1;
        lut_rows(ca, La);
        while (true) {
          int cb = -1;
          if (members) {
            cb = __ffs(members) - 1;
            members &= members - 1;
            lut_rows(cb, Lb);
          }
          column(ca, La);
          if (cb < 0) break;
          ca = -1;
          if (members) {
            ca = __ffs(members) - 1;
            members &= members - 1;
            lut_rows(ca, La);
          }
          column(cb, Lb);
          if (ca < 0) break;
        }
      }
    }
    __syncwarp();  // s_col is rewritten by the next unit
  }
}


// ---------------------------------------------------------------------------------------------
// The same work with every long-latency load taken off the warp's critical path (ncu on the kernel
// above, 65536 x 54 x 32: 30 % of the stall samples in phase A - the (column, layer) inputs are
// strided loads followed by dependent arithmetic, once per unit with nothing to overlap -, 18 % in
// the separate Planck units, and the first use of the concentration-axis rows of every column stalls
// on their load, which was only one column ahead):
//   * the inputs of a warp's NEXT unit (pressures, temperatures, mixing ratios of its 32 columns)
//     travel to shared memory by per-thread asynchronous copies (cp.async, SASS LDGSTS) while the
//     current unit is evaluated;
//   * the Planck function of half level l is produced by the unit of layer l from the temperature it
//     has already loaded (the unit of the last layer adds the surface level): no Planck units;
//   * the columns of a unit are visited in one flat sequence ordered by (it0, ip0) group - the four
//     corner rows of the table gases are reloaded when the group changes -, and the rows of the gas
//     with a concentration axis come through a ring of asynchronous copies TOD_AHEAD columns deep.
constexpr int TOD_AHEAD = 2;              // columns whose concentration-axis rows are in flight beyond the pair being evaluated
constexpr int TOD_RING = TOD_AHEAD + 2;   // ring slots: the pair being read + the columns in flight
constexpr int TOD_NIN = 4 + TOD_MAXGAS;   // staged inputs per column: p(l), p(l+1), T(l), T(l+1), vmr of each gas
static_assert((TOD_RING & (TOD_RING - 1)) == 0, "ring slots: a power of two");

struct __align__(16) Tod2Col {  // g-independent state of one (column, layer), stored in VISIT order
  double c00, c01, c10, c11;    // tw0*pw0, tw0*pw1, tw1*pw0, tw1*pw1
  double w[TOD_MAXGAS];         // what multiplies the interpolated absorption of each gas
  double cw1[TOD_MAXGAS];       // gases with a concentration axis: weight of level ic0 + 1 ...
  unsigned coff[TOD_MAXGAS];    // ... and the element offset ic0 * qC of level ic0 in the packed table
  unsigned o00;                 // element offset of the (it0, ip0) corner block (without the g point)
  int col;                      // column of the tile
  int pad_[2];
};
struct Tod2Smem {  // per warp
  Tod2Col col[TOD_TILE];
  double in[TOD_NIN][TOD_TILE];
  double2 ring[TOD_RING][4][32];
};

template <int NLIN, bool HAS_LUT, int NWARP, int MINB>
__global__ void __launch_bounds__(NWARP * 32, MINB) total_od_planck2_kernel(const __grid_constant__ TotalOdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  Tod2Smem& S = reinterpret_cast<Tod2Smem*>(smem_raw)[warp];
  const unsigned s_in = smem_u32(&S.in[0][0]);
  const unsigned s_ring = smem_u32(&S.ring[0][0][0]);
  const int nlay = p.nlay, ng = p.ng, ngas = p.ngas;
  const int ntile = (p.ncol + TOD_TILE - 1) / TOD_TILE;
  const int64_t n_units = (int64_t)ntile * nlay;
  const int64_t stride = (int64_t)gridDim.x * nwarp;
  const bool want_planck = p.planck_hl != nullptr || p.planck_surf != nullptr;
  const unsigned qP = (unsigned)ng * 4, qT = (unsigned)(p.np - 1) * qP, qC = (unsigned)(p.nt - 1) * qT;

  // asynchronous copies of unit u's inputs into S.in (one column per lane)
  auto stage_inputs = [&](const int64_t u) {
    if (u < n_units) {
      const int tile = (int)(u / nlay), l = (int)(u - (int64_t)tile * nlay);
      const int64_t col = (int64_t)tile * TOD_TILE + lane;
      if (col < p.ncol) {
        const double* p_hl = p.pressure_hl + col * (nlay + 1) + l;
        const double* t_hl = p.temperature_hl + col * (nlay + 1) + l;
        cp_async8(s_in + (0 * TOD_TILE + lane) * 8, p_hl);
        cp_async8(s_in + (1 * TOD_TILE + lane) * 8, p_hl + 1);
        cp_async8(s_in + (2 * TOD_TILE + lane) * 8, t_hl);
        cp_async8(s_in + (3 * TOD_TILE + lane) * 8, t_hl + 1);
#pragma unroll
        for (int ig = 0; ig < TOD_MAXGAS; ++ig)
          if (ig < ngas && p.vmr[ig]) cp_async8(s_in + ((4 + ig) * TOD_TILE + lane) * 8, p.vmr[ig] + col * nlay + l);
      }
    }
    cp_async_commit();
  };

  int64_t u = (int64_t)blockIdx.x * nwarp + warp;
  stage_inputs(u);
  for (; u < n_units; u += stride) {
    const int tile = (int)(u / nlay), l = (int)(u - (int64_t)tile * nlay);
    const int c0 = tile * TOD_TILE;
    const int nc = min(TOD_TILE, p.ncol - c0);
    cp_async_wait<0>();
    // ---- phase A: lane = column (everything that does not depend on g: ckd_model.cpp:962-1009)
    int key = 0x7fffffff;
    const bool valid = lane < nc;
    Interp it = {};
    double t0 = 0.0, t1 = 0.0;
    if (valid) {
      const double p0 = S.in[0][lane], p1 = S.in[1][lane];
      t0 = S.in[2][lane];
      t1 = S.in[3][lane];
      const double t_fl = (t0 * p0 + t1 * p1) / (p0 + p1);
      it = make_interp(p0, p1, t_fl, p.log_p_0, p.d_log_p, p.d_t, p.nt, p.np, p.temperature);
      key = it.it0 * p.np + it.ip0;
    }
    // order of the visit: columns grouped by (it0, ip0), groups by their first column; a column's
    // record goes to its place in that order
    int pos = 0;
    {
      const unsigned same = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(same) - 1;
      unsigned leaders = __ballot_sync(0xffffffffu, valid && lane == leader);
      int base = 0;
      while (leaders) {  // warp-uniform
        const int L = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const unsigned members = __shfl_sync(0xffffffffu, same, L);
        if (leader == L) pos = base + __popc(members & ((1u << lane) - 1u));
        base += __popc(members);
      }
    }
    int pl_off = 0, pl_off1 = 0;
    double pl_w0 = 0.0, pl_w1 = 0.0, pl1_w0 = 0.0, pl1_w1 = 0.0;
    if (valid) {
      Tod2Col& tc = S.col[pos];
      tc.c00 = it.tw0 * it.pw0;
      tc.c01 = it.tw0 * it.pw1;
      tc.c10 = it.tw1 * it.pw0;
      tc.c11 = it.tw1 * it.pw1;
      tc.o00 = (unsigned)it.it0 * qT + (unsigned)it.ip0 * qP;
      tc.col = lane;
#pragma unroll
      for (int ig = 0; ig < TOD_MAXGAS; ++ig) {
        double w = 0.0, cw1 = 0.0;
        unsigned coff = 0u;
        if (ig < ngas) {
          const int cd = p.conc_dependence[ig];
          const bool have = p.vmr[ig] != nullptr;
          const double vmr = have ? S.in[4 + ig][lane] : 0.0;
          double weight = 0.0;  // ckd_model.cpp:980-999
          if (have) weight = (cd == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - p.reference_vmr[ig]) : it.simple_weight * vmr;
          if (cd == ECCKD_CONC_LUT) {
            double cindex0 = (log(vmr) - p.log_vmr0[ig]) / p.d_log_c[ig];  // :1003-1009
            cindex0 = fmax(0.0, fmin(cindex0, p.nconc[ig] - 1.0001));
            const int ic0 = (int)cindex0;
            coff = (unsigned)ic0 * qC;
            cw1 = cindex0 - ic0;
            w = weight;
          } else {
            w = (cd == ECCKD_CONC_NONE) ? it.simple_weight : weight;
          }
        }
        tc.w[ig] = w;
        tc.cw1[ig] = cw1;
        tc.coff[ig] = coff;
      }
      if (want_planck) {  // ckd_model.cpp:1124-1142 at half level l (and at the surface level for the last layer)
        auto planck_index = [&](const double t, int& off, double& w0, double& w1) {
          const double tindex0 = (t - p.planck_t0) / p.planck_dt;
          off = 0;
          w1 = 0.0;
          if (tindex0 >= 0) {
            const int it0 = min((int)tindex0, p.nlut - 2);
            off = it0 * ng;
            w1 = tindex0 - it0;
            w0 = 1.0 - w1;
          } else {  // below the table: (t / t0) * lut[0]
            w0 = t / p.planck_t0;
          }
        };
        planck_index(t0, pl_off, pl_w0, pl_w1);
        planck_index(t1, pl_off1, pl1_w0, pl1_w1);
      }
    }
    __syncwarp();
    // the next unit's inputs travel while this one is evaluated (S.in has been read)
    stage_inputs(u + stride);

    for (int g0 = 0; g0 < ng; g0 += 32) {
      const int g = min(g0 + lane, ng - 1);
      const bool store = g0 + lane < ng;
      // ---- Planck function of the unit's half level(s): rows by lane = g, index and weights from lane = column
      if (want_planck) {
        const double* lut_g = p.planck_lut + g;
        const int64_t lev_stride = (int64_t)(nlay + 1) * ng;
        double* out = p.planck_hl ? p.planck_hl + ((int64_t)c0 * (nlay + 1) + l) * ng + g : nullptr;
        const bool last = l == nlay - 1;
        double* surf = (last && p.planck_surf) ? p.planck_surf + (int64_t)c0 * ng + g : nullptr;
#pragma unroll 8
        for (int c = 0; c < nc; ++c) {
          const int off = __shfl_sync(0xffffffffu, pl_off, c);
          const double w0 = __shfl_sync(0xffffffffu, pl_w0, c), w1 = __shfl_sync(0xffffffffu, pl_w1, c);
          const double b = w0 * __ldg(lut_g + off) + w1 * __ldg(lut_g + off + ng);  // w1 == 0 below the table
          if (store && out) out[c * lev_stride] = b;
        }
        if (last) {
#pragma unroll 8
          for (int c = 0; c < nc; ++c) {
            const int off = __shfl_sync(0xffffffffu, pl_off1, c);
            const double w0 = __shfl_sync(0xffffffffu, pl1_w0, c), w1 = __shfl_sync(0xffffffffu, pl1_w1, c);
            const double b = w0 * __ldg(lut_g + off) + w1 * __ldg(lut_g + off + ng);
            if (store && out) out[c * lev_stride + ng] = b;
            if (store && surf) surf[c * ng] = b;
          }
        }
      }
      // ---- optical depth: lane = g point, columns in visit order
      const unsigned col_stride = (unsigned)nlay * ng;  // 32 * nlay * ng < 2^31 (checked by the host)
      double* const od_unit = p.od + ((int64_t)c0 * nlay + l) * ng + g;
      const unsigned g4 = (unsigned)g * 4;
      // asynchronous copies of the concentration-axis rows of the k-th column of the visit
      auto ring_issue = [&](const int k) {
        if (HAS_LUT && k < nc) {
          const double* k0 = p.packed[NLIN] + (size_t)(S.col[k].coff[NLIN] + S.col[k].o00 + g4);
          const unsigned dst = s_ring + (unsigned)(((k & (TOD_RING - 1)) * 4) * 32 + lane) * 16;
          cp_async16(dst, k0);
          cp_async16(dst + 32 * 16, k0 + 2);
          cp_async16(dst + 64 * 16, k0 + qC);
          cp_async16(dst + 96 * 16, k0 + qC + 2);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int k = 0; k < TOD_AHEAD; ++k) ring_issue(k);
      unsigned cur = 0xffffffffu;
      double K[NLIN > 0 ? NLIN : 1][4];
      // Two columns of the same (it0, ip0) group per trip, their loads issued together and their
      // arithmetic interleaved (a single column is a chain of dependent FP64 instructions behind a
      // shared-memory load: with three warps per scheduler that latency was the whole run time).
      // At the start of a trip at column k, the rows of columns k .. k + TOD_AHEAD - 1 are in flight.
#pragma unroll 1
      for (int k = 0; k < nc;) {
        const Tod2Col& ta = S.col[k];
        const unsigned o00 = ta.o00;
        const bool pair = k + 1 < nc && S.col[k + 1].o00 == o00;  // warp-uniform
        ring_issue(k + TOD_AHEAD);
        if (pair) ring_issue(k + TOD_AHEAD + 1);
        if (o00 != cur) {  // warp-uniform: a new (it0, ip0) group, four corner rows of every table gas
          cur = o00;
#pragma unroll
          for (int ig = 0; ig < NLIN; ++ig) {
            const double2* kk = reinterpret_cast<const double2*>(p.packed[ig] + (o00 + g4));
            const double2 a = __ldg(kk), b = __ldg(kk + 1);
            K[ig][0] = a.x;
            K[ig][1] = a.y;
            K[ig][2] = b.x;
            K[ig][3] = b.y;
          }
        }
        const Tod2Col& tb = S.col[pair ? k + 1 : k];
        const double a00 = ta.c00, a01 = ta.c01, a10 = ta.c10, a11 = ta.c11;
        const double b00 = tb.c00, b01 = tb.c01, b10 = tb.c10, b11 = tb.c11;
        double sa0 = 0.0, sa1 = 0.0, sb0 = 0.0, sb1 = 0.0;  // two chains per column
#pragma unroll
        for (int ig = 0; ig < NLIN; ++ig) {
          const double va = a00 * K[ig][0] + a01 * K[ig][1] + a10 * K[ig][2] + a11 * K[ig][3];
          const double vb = b00 * K[ig][0] + b01 * K[ig][1] + b10 * K[ig][2] + b11 * K[ig][3];
          if (ig & 1) {
            sa1 += ta.w[ig] * va;
            sb1 += tb.w[ig] * vb;
          } else {
            sa0 += ta.w[ig] * va;
            sb0 += tb.w[ig] * vb;
          }
        }
        if (HAS_LUT) {
          cp_async_wait<TOD_AHEAD>();  // the copies of this trip's columns have landed (a lane reads what it copied itself)
          const double2* Ra = &S.ring[k & (TOD_RING - 1)][0][lane];
          const double2* Rb = &S.ring[(pair ? k + 1 : k) & (TOD_RING - 1)][0][lane];
          const double2 A0 = Ra[0], A1 = Ra[32], A2 = Ra[64], A3 = Ra[96];
          const double2 B0 = Rb[0], B1 = Rb[32], B2 = Rb[64], B3 = Rb[96];
          const double cwa = ta.cw1[NLIN], cwb = tb.cw1[NLIN];
          const double xa = a00 * A0.x + a01 * A0.y + a10 * A1.x + a11 * A1.y;
          const double ya = a00 * A2.x + a01 * A2.y + a10 * A3.x + a11 * A3.y;
          const double xb = b00 * B0.x + b01 * B0.y + b10 * B1.x + b11 * B1.y;
          const double yb = b00 * B2.x + b01 * B2.y + b10 * B3.x + b11 * B3.y;
          sa1 += ta.w[NLIN] * ((1.0 - cwa) * xa + cwa * ya);
          sb1 += tb.w[NLIN] * ((1.0 - cwb) * xb + cwb * yb);
        }
        double tau_a = sa0 + sa1, tau_b = sb0 + sb1;
        for (int ig = NLIN + (HAS_LUT ? 1 : 0); ig < ngas; ++ig) {  // further gases with a concentration axis
          const double* ka = p.packed[ig] + (size_t)(ta.coff[ig] + o00 + g4);
          const double* kb = p.packed[ig] + (size_t)(tb.coff[ig] + o00 + g4);
          const double xa = a00 * ka[0] + a01 * ka[1] + a10 * ka[2] + a11 * ka[3];
          const double xb = b00 * kb[0] + b01 * kb[1] + b10 * kb[2] + b11 * kb[3];
          ka += qC;
          kb += qC;
          const double ya = a00 * ka[0] + a01 * ka[1] + a10 * ka[2] + a11 * ka[3];
          const double yb = b00 * kb[0] + b01 * kb[1] + b10 * kb[2] + b11 * kb[3];
          tau_a += ta.w[ig] * ((1.0 - ta.cw1[ig]) * xa + ta.cw1[ig] * ya);
          tau_b += tb.w[ig] * ((1.0 - tb.cw1[ig]) * xb + tb.cw1[ig] * yb);
        }
        if (store) {  // run_ckd.cpp:317
          od_unit[(unsigned)ta.col * col_stride] = fmax(tau_a, 0.0);
          if (pair) od_unit[(unsigned)tb.col * col_stride] = fmax(tau_b, 0.0);
        }
        k += pair ? 2 : 1;
      }
    }
    __syncwarp();  // S.col is rewritten by the next unit
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// total_od_planck3_kernel: the column loop of total_od_planck_kernel (table rows in registers, the
// concentration-axis rows of the next column in flight) with the two changes of the kernel above
// that paid: the inputs of a warp's next unit staged by asynchronous copies while the current one
// is evaluated, and the Planck function of half level l produced by the unit of layer l.
struct Tod3Smem {  // per warp
  TodCol col[TOD_TILE];
  double in[TOD_NIN][TOD_TILE];
};

template <int NLIN, bool HAS_LUT, int NWARP, int MINB>
__global__ void __launch_bounds__(NWARP * 32, MINB) total_od_planck3_kernel(const __grid_constant__ TotalOdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  Tod3Smem& S = reinterpret_cast<Tod3Smem*>(smem_raw)[warp];
  TodCol* s_col = S.col;
  const unsigned s_in = smem_u32(&S.in[0][0]);
  const int nlay = p.nlay, ng = p.ng, ngas = p.ngas;
  const int ntile = (p.ncol + TOD_TILE - 1) / TOD_TILE;
  const int64_t n_units = (int64_t)ntile * nlay;
  const int64_t stride = (int64_t)gridDim.x * nwarp;
  const bool want_planck = p.planck_hl != nullptr || p.planck_surf != nullptr;
  // strides of the corner-packed tables, in elements (< 2^31)
  const unsigned qP = (unsigned)ng * 4, qT = (unsigned)(p.np - 1) * qP, qC = (unsigned)(p.nt - 1) * qT;

  // asynchronous copies of unit u's inputs into S.in (one column per lane)
  auto stage_inputs = [&](const int64_t u) {
    if (u < n_units) {
      const int tile = (int)(u / nlay), l = (int)(u - (int64_t)tile * nlay);
      const int64_t col = (int64_t)tile * TOD_TILE + lane;
      if (col < p.ncol) {
        const double* p_hl = p.pressure_hl + col * (nlay + 1) + l;
        const double* t_hl = p.temperature_hl + col * (nlay + 1) + l;
        cp_async8(s_in + (0 * TOD_TILE + lane) * 8, p_hl);
        cp_async8(s_in + (1 * TOD_TILE + lane) * 8, p_hl + 1);
        cp_async8(s_in + (2 * TOD_TILE + lane) * 8, t_hl);
        cp_async8(s_in + (3 * TOD_TILE + lane) * 8, t_hl + 1);
#pragma unroll
        for (int ig = 0; ig < TOD_MAXGAS; ++ig)
          if (ig < ngas && p.vmr[ig]) cp_async8(s_in + ((4 + ig) * TOD_TILE + lane) * 8, p.vmr[ig] + col * nlay + l);
      }
    }
    cp_async_commit();
  };

  int64_t u = (int64_t)blockIdx.x * nwarp + warp;
  stage_inputs(u);
  for (; u < n_units; u += stride) {
    // consecutive warps take consecutive layers of one tile: its columns' profile sectors are shared in L1/L2
    const int tile = (int)(u / nlay), l = (int)(u - (int64_t)tile * nlay);
    const int c0 = tile * TOD_TILE;
    const int nc = min(TOD_TILE, p.ncol - c0);
    cp_async_wait<0>();
    // ---- phase A: lane = column
    int key = -1 - lane;  // columns past the end form groups of their own that nobody visits
    const bool valid = c0 + lane < p.ncol;
    int pl_off = 0, pl_off1 = 0;
    double pl_w0 = 0.0, pl_w1 = 0.0, pl1_w0 = 0.0, pl1_w1 = 0.0;
    if (valid) {
      const double p0 = S.in[0][lane], p1 = S.in[1][lane], t0 = S.in[2][lane], t1 = S.in[3][lane];
      const double t_fl = (t0 * p0 + t1 * p1) / (p0 + p1);
      const Interp it = make_interp(p0, p1, t_fl, p.log_p_0, p.d_log_p, p.d_t, p.nt, p.np, p.temperature);
      key = it.it0 * p.np + it.ip0;
      TodCol tc;
      tc.c00 = it.tw0 * it.pw0;
      tc.c01 = it.tw0 * it.pw1;
      tc.c10 = it.tw1 * it.pw0;
      tc.c11 = it.tw1 * it.pw1;
#pragma unroll
      for (int ig = 0; ig < TOD_MAXGAS; ++ig) {
        tc.w[ig] = 0.0;
        tc.cw1[ig] = 0.0;
        tc.ic0[ig] = 0;
        if (ig >= ngas) continue;
        const int cd = p.conc_dependence[ig];
        const bool have = p.vmr[ig] != nullptr;
        const double vmr = have ? S.in[4 + ig][lane] : 0.0;
        double weight = 0.0;  // ckd_model.cpp:980-999
        if (have) weight = (cd == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - p.reference_vmr[ig]) : it.simple_weight * vmr;
        if (cd == ECCKD_CONC_LUT) {
          double cindex0 = (log(vmr) - p.log_vmr0[ig]) / p.d_log_c[ig];  // :1003-1009
          cindex0 = fmax(0.0, fmin(cindex0, p.nconc[ig] - 1.0001));
          tc.ic0[ig] = (int)cindex0;
          tc.cw1[ig] = cindex0 - tc.ic0[ig];
          tc.w[ig] = weight;
        } else {
          tc.w[ig] = (cd == ECCKD_CONC_NONE) ? it.simple_weight : weight;
        }
      }
      s_col[lane] = tc;
      if (want_planck) {  // ckd_model.cpp:1124-1142 at half level l (and at the surface level for the last layer)
        auto planck_index = [&](const double t, int& off, double& w0, double& w1) {
          const double tindex0 = (t - p.planck_t0) / p.planck_dt;
          off = 0;
          w1 = 0.0;
          if (tindex0 >= 0) {
            const int it0 = min((int)tindex0, p.nlut - 2);
            off = it0 * ng;
            w1 = tindex0 - it0;
            w0 = 1.0 - w1;
          } else {  // below the table: (t / t0) * lut[0]
            w0 = t / p.planck_t0;
          }
        };
        planck_index(t0, pl_off, pl_w0, pl_w1);
        planck_index(t1, pl_off1, pl1_w0, pl1_w1);
      }
    }
    const unsigned same = __match_any_sync(0xffffffffu, key);
    __syncwarp();
    // the next unit's inputs travel while this one is evaluated (S.in has been read)
    stage_inputs(u + stride);
    // ---- Planck function of the unit's half level(s): rows by lane = g, index and weights from lane = column
    if (want_planck) {
      const int64_t lev_stride = (int64_t)(nlay + 1) * ng;
      const bool last = l == nlay - 1;
      for (int g0 = 0; g0 < ng; g0 += 32) {
        const int g = min(g0 + lane, ng - 1);
        const bool store = g0 + lane < ng;
        const double* lut_g = p.planck_lut + g;
        double* out = p.planck_hl ? p.planck_hl + ((int64_t)c0 * (nlay + 1) + l) * ng + g : nullptr;
        double* surf = (last && p.planck_surf) ? p.planck_surf + (int64_t)c0 * ng + g : nullptr;
#pragma unroll 8
        for (int c = 0; c < nc; ++c) {
          const int off = __shfl_sync(0xffffffffu, pl_off, c);
          const double w0 = __shfl_sync(0xffffffffu, pl_w0, c), w1 = __shfl_sync(0xffffffffu, pl_w1, c);
          const double b = w0 * __ldg(lut_g + off) + w1 * __ldg(lut_g + off + ng);  // w1 == 0 below the table
          if (store && out) out[c * lev_stride] = b;
        }
        if (last) {
#pragma unroll 8
          for (int c = 0; c < nc; ++c) {
            const int off = __shfl_sync(0xffffffffu, pl_off1, c);
            const double w0 = __shfl_sync(0xffffffffu, pl1_w0, c), w1 = __shfl_sync(0xffffffffu, pl1_w1, c);
            const double b = w0 * __ldg(lut_g + off) + w1 * __ldg(lut_g + off + ng);
            if (store && out) out[c * lev_stride + ng] = b;
            if (store && surf) surf[c * ng] = b;
          }
        }
      }
    }
    // ---- phase B: lane = g point, groups of columns with the same four table corners
    const unsigned col_stride = (unsigned)nlay * ng;  // 32 * nlay * ng < 2^31 (checked by the host)
    for (int g0 = 0; g0 < ng; g0 += 32) {
      const int g = min(g0 + lane, ng - 1);
      const bool store = g0 + lane < ng;
      double* const od_unit = p.od + ((int64_t)c0 * nlay + l) * ng + g;
      unsigned todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        unsigned members = __shfl_sync(0xffffffffu, same, leader);
        const int gkey = __shfl_sync(0xffffffffu, key, leader);
        todo &= ~members;
        const int it0 = gkey / p.np, ip0 = gkey - it0 * p.np;
        const unsigned o00 = (unsigned)it0 * qT + (unsigned)ip0 * qP + (unsigned)g * 4;
        double K[NLIN > 0 ? NLIN : 1][4];
#pragma unroll
        for (int ig = 0; ig < NLIN; ++ig) {
          const double2* k = reinterpret_cast<const double2*>(p.packed[ig] + o00);
          const double2 a = __ldg(k), b = __ldg(k + 1);
          K[ig][0] = a.x;
          K[ig][1] = a.y;
          K[ig][2] = b.x;
          K[ig][3] = b.y;
        }
        const double* lut_base = HAS_LUT ? p.packed[NLIN] + o00 : nullptr;
        auto lut_rows = [&](int c, double2 (&L)[4]) {
          if (!HAS_LUT) return;
          const double2* k = reinterpret_cast<const double2*>(lut_base + (size_t)((unsigned)s_col[c].ic0[NLIN] * qC));
          const double2* k1 = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(k) + qC);
          L[0] = __ldg(k);
          L[1] = __ldg(k + 1);
          L[2] = __ldg(k1);
          L[3] = __ldg(k1 + 1);
        };
        auto column = [&](int c, const double2 (&L)[4]) {
          const TodCol& tc = s_col[c];
          const double c00 = tc.c00, c01 = tc.c01, c10 = tc.c10, c11 = tc.c11;
          double tau0 = 0.0, tau1 = 0.0;  // two chains
#pragma unroll
          for (int ig = 0; ig < NLIN; ++ig) {
            const double v = c00 * K[ig][0] + c01 * K[ig][1] + c10 * K[ig][2] + c11 * K[ig][3];
            if (ig & 1) tau1 += tc.w[ig] * v; else tau0 += tc.w[ig] * v;
          }
          if (HAS_LUT) {
            const double cw1 = tc.cw1[NLIN];
            const double a = c00 * L[0].x + c01 * L[0].y + c10 * L[1].x + c11 * L[1].y;
            const double b = c00 * L[2].x + c01 * L[2].y + c10 * L[3].x + c11 * L[3].y;
            tau1 += tc.w[NLIN] * ((1.0 - cw1) * a + cw1 * b);
          }
          double tau = tau0 + tau1;
          for (int ig = NLIN + (HAS_LUT ? 1 : 0); ig < ngas; ++ig) {  // further gases with a concentration axis
            const double cw1 = tc.cw1[ig];
            const double* k = p.packed[ig] + (size_t)((unsigned)tc.ic0[ig] * qC) + o00;
            const double a = c00 * k[0] + c01 * k[1] + c10 * k[2] + c11 * k[3];
            k += qC;
            const double b = c00 * k[0] + c01 * k[1] + c10 * k[2] + c11 * k[3];
            tau += tc.w[ig] * ((1.0 - cw1) * a + cw1 * b);
          }
          if (store) od_unit[(unsigned)c * col_stride] = fmax(tau, 0.0);  // run_ckd.cpp:317
        };
        // software pipeline over the members: the next column's LUT rows are loaded before this one is summed
        double2 La[4], Lb[4];
        int ca = __ffs(members) - 1;
        members &= members - 1;
        lut_rows(ca, La);
        while (true) {
          int cb = -1;
          if (members) {
            cb = __ffs(members) - 1;
            members &= members - 1;
            lut_rows(cb, Lb);
          }
          column(ca, La);
          if (cb < 0) break;
          ca = -1;
          if (members) {
            ca = __ffs(members) - 1;
            members &= members - 1;
            lut_rows(ca, La);
          }
          column(cb, Lb);
          if (ca < 0) break;
        }
      }
    }
    __syncwarp();  // s_col is rewritten by the next unit
  }
  cp_async_wait<0>();
}


// ---------------------------------------------------------------------------------------------
// Adjoint of the look-up-table interpolation: the optical depth is linear in the table,
//   tau(cl, g) = sum_gas W_gas(cl) sum_corners c(cl, corner) [sum_conc cw] K_gas(node, g)   (ckd_model.cpp:962-1088),
// so dJ/dK_gas(node, g) = sum over the (column, layer) pairs cl whose interpolation cell touches the node of
// dJ/dtau(cl, g) times that coefficient - the transpose of the forward gather, done as a GATHER per table node so
// that every sum has a fixed order (no floating-point atomics): kernel 1 writes the g-independent record of every
// cl; kernel 2 runs one block per node (it, ip): its warps scan fixed contiguous ranges of cl in ascending order,
// add the contributions of the cl in the (up to) four adjacent cells into per-warp accumulators in shared memory
// (lane = g point) and the warps' partial sums are added in warp order.
constexpr int ADJ_MAXGAS = TOD_MAXGAS;
struct AdjParams {
  const double* vmr[ADJ_MAXGAS];
  double reference_vmr[ADJ_MAXGAS], log_vmr0[ADJ_MAXGAS], d_log_c[ADJ_MAXGAS];
  int conc_dependence[ADJ_MAXGAS], nconc[ADJ_MAXGAS], present[ADJ_MAXGAS], acc_off[ADJ_MAXGAS];
  double* grad[ADJ_MAXGAS];  // [nconc|1][nt][np][ng] or NULL
  int ngas, nt, np, ng, nacc;
  double log_p_0, d_log_p, d_t;
  const double* temperature;
  int ncol, nlay;
  const double* pressure_hl;
  const double* temperature_fl;
  const double* grad_od;  // [ncol][nlay][ng]
  // records, one per cl
  int* key;               // it0 * np + ip0
  double* corner;         // [4][ncl]: tw0 pw0, tw0 pw1, tw1 pw0, tw1 pw1
  double* w;              // [ngas][ncl] what multiplies the interpolated absorption of the gas (0: gas absent)
  double* cw1;            // [ngas][ncl]
  int* ic0;               // [ngas][ncl]
  int accumulate;
};

__global__ void od_adjoint_records_kernel(const AdjParams p) {
  const int64_t ncl = (int64_t)p.ncol * p.nlay;
  const int64_t cl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cl >= ncl) return;
  const int icol = (int)(cl / p.nlay), il = (int)(cl - (int64_t)icol * p.nlay);
  const double p0 = p.pressure_hl[(int64_t)icol * (p.nlay + 1) + il];
  const double p1 = p.pressure_hl[(int64_t)icol * (p.nlay + 1) + il + 1];
  const Interp it = make_interp(p0, p1, p.temperature_fl[cl], p.log_p_0, p.d_log_p, p.d_t, p.nt, p.np, p.temperature);
  p.key[cl] = it.it0 * p.np + it.ip0;
  p.corner[0 * ncl + cl] = it.tw0 * it.pw0;
  p.corner[1 * ncl + cl] = it.tw0 * it.pw1;
  p.corner[2 * ncl + cl] = it.tw1 * it.pw0;
  p.corner[3 * ncl + cl] = it.tw1 * it.pw1;
  for (int ig = 0; ig < p.ngas; ++ig) {
    double w = 0.0, cw1 = 0.0;
    int ic0 = 0;
    if (p.present[ig]) {  // as gas_od
      const bool have = p.vmr[ig] != nullptr;
      const double vmr = have ? p.vmr[ig][cl] : 0.0;
      const int cd = p.conc_dependence[ig];
      double weight = 0.0;
      if (have) weight = (cd == ECCKD_CONC_RELATIVE_LINEAR) ? it.simple_weight * (vmr - p.reference_vmr[ig]) : it.simple_weight * vmr;
      if (cd == ECCKD_CONC_LUT) {
        double cindex0 = (log(vmr) - p.log_vmr0[ig]) / p.d_log_c[ig];
        cindex0 = fmax(0.0, fmin(cindex0, p.nconc[ig] - 1.0001));
        ic0 = (int)cindex0;
        cw1 = cindex0 - ic0;
        w = weight;
      } else {
        w = (cd == ECCKD_CONC_NONE) ? it.simple_weight : weight;
      }
    }
    p.w[(int64_t)ig * ncl + cl] = w;
    p.cw1[(int64_t)ig * ncl + cl] = cw1;
    p.ic0[(int64_t)ig * ncl + cl] = ic0;
  }
}

constexpr int ADJ_WARPS = 8;
__global__ void __launch_bounds__(ADJ_WARPS * 32) od_adjoint_gather_kernel(const AdjParams p) {
  extern __shared__ double s_acc[];  // [ADJ_WARPS][nacc][32]
  const int node = blockIdx.x, it = node / p.np, ip = node - it * p.np;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ncl = (int64_t)p.ncol * p.nlay;
  const int64_t per = (ncl + ADJ_WARPS - 1) / ADJ_WARPS;
  const int64_t lo = min(ncl, (int64_t)warp * per), hi = min(ncl, lo + per);
  double* acc = s_acc + (size_t)warp * p.nacc * 32;
  for (int g0 = 0; g0 < p.ng; g0 += 32) {
    const int g = min(g0 + lane, p.ng - 1);
    for (int a = 0; a < p.nacc; ++a) acc[a * 32 + lane] = 0.0;
    __syncwarp();
    // 32 records are tested at a time (lane = record), the ones whose cell touches the node are then taken in
    // ascending order (lane = g point): the order of the additions is that of a serial scan
    for (int64_t base = lo; base < hi; base += 32) {
      const int64_t mine = base + lane;
      int corner_of = -1;
      if (mine < hi) {
        const int key = p.key[mine];
        const int it0 = key / p.np, ip0 = key - it0 * p.np;
        const int dt = it - it0, dp = ip - ip0;
        if ((unsigned)dt <= 1u && (unsigned)dp <= 1u) corner_of = dt * 2 + dp;
      }
      unsigned hits = __ballot_sync(0xffffffffu, corner_of >= 0);
      while (hits) {  // warp-uniform
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const int64_t cl = base + j;
        const int cn = __shfl_sync(0xffffffffu, corner_of, j);
        const double c = p.corner[(int64_t)cn * ncl + cl];
        const double gt = p.grad_od[cl * p.ng + g] * c;
        // the per-gas fields of the record: lane ig fetches gas ig's (one memory latency for all gases), then broadcast
        double w_l = 0.0, cw_l = 0.0;
        int ic_l = 0;
        if (lane < p.ngas) {
          w_l = p.w[(int64_t)lane * ncl + cl];
          cw_l = p.cw1[(int64_t)lane * ncl + cl];
          ic_l = p.ic0[(int64_t)lane * ncl + cl];
        }
        for (int ig = 0; ig < p.ngas; ++ig) {
          const double w = __shfl_sync(0xffffffffu, w_l, ig);
          const double cw1 = __shfl_sync(0xffffffffu, cw_l, ig);
          const int ic0 = __shfl_sync(0xffffffffu, ic_l, ig);
          if (!p.grad[ig]) continue;  // warp-uniform
          double* a = acc + (size_t)p.acc_off[ig] * 32 + lane;
          if (p.conc_dependence[ig] == ECCKD_CONC_LUT) {
            a[(size_t)ic0 * 32] += gt * (w * (1.0 - cw1));
            a[(size_t)(ic0 + 1) * 32] += gt * (w * cw1);
          } else {
            a[0] += gt * w;
          }
        }
      }
    }
    __syncthreads();
    // warps' partial sums in warp order, one (accumulator, g) per thread
    for (int j = threadIdx.x; j < p.nacc * 32; j += blockDim.x) {
      const int a = j >> 5, l = j & 31;
      if (g0 + l >= p.ng) continue;
      double sum = 0.0;
      for (int w = 0; w < ADJ_WARPS; ++w) sum += s_acc[((size_t)w * p.nacc + a) * 32 + l];
      // which gas / concentration level is accumulator a
      int ig = 0;
      while (ig + 1 < p.ngas && (!p.grad[ig] || a >= p.acc_off[ig] + (p.conc_dependence[ig] == ECCKD_CONC_LUT ? p.nconc[ig] : 1))) ++ig;
      if (!p.grad[ig]) continue;
      const int ic = a - p.acc_off[ig];
      double* dst = p.grad[ig] + (((size_t)ic * p.nt + it) * p.np + ip) * p.ng + g0 + l;
      *dst = p.accumulate ? (*dst + sum) : sum;
    }
    __syncthreads();
  }
}

}  // namespace ecckd

using namespace ecckd;

static int upload_new(const double* h, size_t n, double** d) {
  *d = nullptr;
  if (!h || n == 0) return ECCKD_OK;
  ECCKD_CUDA_OK(cudaMalloc(d, n * sizeof(double)));
  ECCKD_CUDA_OK(cudaMemcpy(*d, h, n * sizeof(double), cudaMemcpyHostToDevice));
  return ECCKD_OK;
}

ECCKD_API int ecckd_ckd_model_destroy(ecckd_ckd_model* m) {
  ECCKD_DEVICE(m ? m->ctx : nullptr);
  if (!m) return ECCKD_OK;
  if (m->ctx) cudaStreamSynchronize(m->ctx->stream);
  for (double* p : m->d_molar_abs) cudaFree(p);
  for (double* p : m->d_packed) cudaFree(p);
  for (double* p : m->d_vmr_lut) cudaFree(p);
  cudaFree(m->d_log_pressure);
  cudaFree(m->d_temperature);
  cudaFree(m->d_temperature_planck);
  cudaFree(m->d_planck_lut);
  cudaFree(m->d_solar_irradiance);
  cudaFree(m->d_rayleigh);
  cudaFree(m->d_gasdesc);
  cudaFree(m->d_ones);
  delete m;
  return ECCKD_OK;
}

ECCKD_API int ecckd_ckd_model_create(ecckd_ctx* ctx, int ngas, const ecckd_gas_table* gases, int nt, int np, int ng,
                                     const double* log_pressure, const double* temperature, int nlut,
                                     const double* temperature_planck, const double* planck_lut,
                                     const double* solar_irradiance, const double* rayleigh_molar_scat,
                                     ecckd_ckd_model** out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && gases && log_pressure && temperature && out, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_model_create: NULL argument");
  ECCKD_REQUIRE(ngas > 0 && nt >= 2 && np >= 2 && ng > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_model_create: need ngas>0, nt>=2, np>=2, ng>0");
  const bool is_sw = solar_irradiance != nullptr;
  ECCKD_REQUIRE(is_sw || (planck_lut && temperature_planck && nlut >= 2), ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_model_create: longwave model needs a Planck look-up table");
  ECCKD_REQUIRE(!is_sw || rayleigh_molar_scat, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_model_create: shortwave model needs rayleigh_molar_scat");
  *out = nullptr;
  ecckd_ckd_model* m = new ecckd_ckd_model();
  m->ctx = ctx;
  m->ngas = ngas;
  m->nt = nt;
  m->np = np;
  m->ng = ng;
  m->nlut = nlut;
  m->is_sw = is_sw;
  int rc = ECCKD_OK;
  for (int i = 0; i < ngas && rc == ECCKD_OK; ++i) {
    const ecckd_gas_table& g = gases[i];
    m->conc_dependence.push_back(g.conc_dependence);
    m->nconc.push_back(g.nconc);
    m->reference_vmr.push_back(g.reference_vmr);
    double *d_k = nullptr, *d_v = nullptr;
    size_t n = (size_t)nt * np * ng;
    double lv0 = 0.0, dlc = 1.0;
    if (g.conc_dependence == ECCKD_CONC_LUT) {
      if (g.nconc < 2 || !g.vmr_lut) {
        set_error("gas %d: a LUT gas needs nconc >= 2 mole fractions", i);
        rc = ECCKD_PARAMETER_ERROR;
        break;
      }
      n *= (size_t)g.nconc;
      rc = upload_new(g.vmr_lut, g.nconc, &d_v);
      lv0 = log(g.vmr_lut[0]);                  // ckd_model.cpp:1005
      dlc = log(g.vmr_lut[1] / g.vmr_lut[0]);   // ckd_model.cpp:1004
    } else if (g.conc_dependence < 0 || g.conc_dependence > 3) {
      set_error("gas %d: conc_dependence %d not understood", i, g.conc_dependence);
      rc = ECCKD_PARAMETER_ERROR;
      break;
    }
    if (rc == ECCKD_OK) rc = upload_new(g.molar_abs, n, &d_k);
    m->d_molar_abs.push_back(d_k);
    double* d_q = nullptr;
    if (rc == ECCKD_OK && !is_sw) {  // the layer-major optical-depth kernel reads whole interpolation cells
      const size_t nc = g.conc_dependence == ECCKD_CONC_LUT ? (size_t)g.nconc : 1;
      std::vector<double> q(nc * (nt - 1) * (np - 1) * ng * 4);
      size_t o = 0;
      for (size_t ic = 0; ic < nc; ++ic)
        for (int it = 0; it < nt - 1; ++it)
          for (int ip = 0; ip < np - 1; ++ip)
            for (int ig = 0; ig < ng; ++ig)
              for (int corner = 0; corner < 4; ++corner)
                q[o++] = g.molar_abs[((ic * nt + it + (corner >> 1)) * np + ip + (corner & 1)) * ng + ig];
      rc = upload_new(q.data(), q.size(), &d_q);
    }
    m->d_packed.push_back(d_q);
    m->d_vmr_lut.push_back(d_v);
    m->log_vmr0.push_back(lv0);
    m->d_log_c.push_back(dlc);
  }
  if (rc == ECCKD_OK) rc = upload_new(log_pressure, np, &m->d_log_pressure);
  if (rc == ECCKD_OK) rc = upload_new(temperature, (size_t)nt * np, &m->d_temperature);
  if (rc == ECCKD_OK && !is_sw) rc = upload_new(temperature_planck, nlut, &m->d_temperature_planck);
  if (rc == ECCKD_OK && !is_sw) rc = upload_new(planck_lut, (size_t)nlut * ng, &m->d_planck_lut);
  if (rc == ECCKD_OK && is_sw) rc = upload_new(solar_irradiance, ng, &m->d_solar_irradiance);
  if (rc == ECCKD_OK && is_sw) rc = upload_new(rayleigh_molar_scat, ng, &m->d_rayleigh);
  if (rc == ECCKD_OK) {
    cudaError_t e = cudaMalloc(&m->d_gasdesc, (size_t)ngas * sizeof(GasDev));
    if (e != cudaSuccess) {
      set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
      rc = ECCKD_UNEXPECTED_EXCEPTION;
    }
  }
  if (rc != ECCKD_OK) {
    ecckd_ckd_model_destroy(m);
    return rc;
  }
  m->log_p_0 = log_pressure[0];
  m->d_log_p = log_pressure[1] - log_pressure[0];
  m->d_t = temperature[(size_t)1 * np + 0] - temperature[0];
  if (!is_sw) {
    m->planck_t0 = temperature_planck[0];
    m->planck_dt = temperature_planck[1] - temperature_planck[0];
  } else {
    double s = 0.0;
    for (int g = 0; g < ng; ++g) s += solar_irradiance[g];
    m->solar_sum = s;
  }
  *out = m;
  return ECCKD_OK;
}

static GasDev make_gasdev(const ecckd_ckd_model* m, int igas, const double* d_vmr) {
  GasDev g;
  g.molar_abs = m->d_molar_abs[igas];
  g.vmr = d_vmr;
  g.reference_vmr = m->reference_vmr[igas];
  g.log_vmr0 = m->log_vmr0[igas];
  g.d_log_c = m->d_log_c[igas];
  g.conc_dependence = m->conc_dependence[igas];
  g.nconc = m->nconc[igas];
  return g;
}

static int check_vmr(const ecckd_ckd_model* m, int igas, const double* d_vmr) {
  // ckd_model.cpp:1011-1014, 1048-1051: THROW(1)
  if (m->conc_dependence[igas] != ECCKD_CONC_NONE && d_vmr == nullptr) {
    set_error("Concentration of gas %d not provided", igas);
    return 1;
  }
  return ECCKD_OK;
}

static dim3 g_block(int ng) { return dim3(ng <= 32 ? 32 : (ng <= 64 ? 64 : 128), ng <= 32 ? 8 : (ng <= 64 ? 4 : 2)); }

ECCKD_API int ecckd_ckd_calc_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, int igas, int ncol, int nlay,
                                               const double* d_pressure_hl, const double* d_temperature_fl,
                                               const double* d_vmr_fl, double* d_od, int accumulate) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && d_pressure_hl && d_temperature_fl && d_od, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_calc_optical_depth: NULL argument");
  ECCKD_REQUIRE(igas >= 0 && igas < m->ngas && ncol > 0 && nlay > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_calc_optical_depth: bad gas index or dimensions");
  ECCKD_TRY(check_vmr(m, igas, d_vmr_fl));
  const dim3 block = g_block(m->ng);
  const int64_t ncl = (int64_t)ncol * nlay;
  const dim3 grid((unsigned)((ncl + block.y - 1) / block.y), (m->ng + block.x - 1) / block.x);
  ckd_od_kernel<<<grid, block, 0, ctx->stream>>>(make_gasdev(m, igas, d_vmr_fl), m->nt, m->np, m->ng, m->log_p_0,
                                                 m->d_log_p, m->d_t, m->d_temperature, ncol, nlay, d_pressure_hl,
                                                 d_temperature_fl, d_od, accumulate);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// calc_total_optical_depth: src/ecckd/solve_adept.cpp:23-69.  Rayleigh optical depth (shortwave
// models), then every gas of the model: with its concentration when the training scene holds it,
// without when it does not and the gas has no concentration dependence, skipped otherwise.
ECCKD_API int ecckd_calc_total_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, const double* const* h_d_vmr_fl,
                                                 int ncol, int nlay, const double* d_pressure_hl,
                                                 const double* d_temperature_hl, double* d_od) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && h_d_vmr_fl && d_pressure_hl && d_temperature_hl && d_od && ncol > 0 && nlay > 0,
                ECCKD_PARAMETER_ERROR, "ecckd_calc_total_optical_depth: bad argument");
  const int64_t ncl = (int64_t)ncol * nlay;
  void* d_tfl = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, (size_t)ncl * sizeof(double), &d_tfl));
  temperature_fl_kernel<<<(unsigned)((ncl + 255) / 256), 256, 0, ctx->stream>>>(ncol, nlay, d_temperature_hl,
                                                                              d_pressure_hl, static_cast<double*>(d_tfl));
  ECCKD_LAUNCH_CHECK(ctx);
  if (m->is_sw) {
    ECCKD_TRY(ecckd_ckd_calc_rayleigh_optical_depth_dev(ctx, m, ncol, nlay, d_pressure_hl, d_od));
  } else {
    ECCKD_CUDA_OK(cudaMemsetAsync(d_od, 0, (size_t)ncl * m->ng * sizeof(double), ctx->stream));
  }
  for (int igas = 0; igas < m->ngas; ++igas) {
    if (!h_d_vmr_fl[igas] && m->conc_dependence[igas] != ECCKD_CONC_NONE) continue;  // :62-66 "not present in LBL file"
    ECCKD_TRY(ecckd_ckd_calc_optical_depth_dev(ctx, m, igas, ncol, nlay, d_pressure_hl, static_cast<double*>(d_tfl),
                                               h_d_vmr_fl[igas], d_od, 1));
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_calc_total_optical_depth_adjoint_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, const double* const* h_d_vmr_fl,
                                                         int ncol, int nlay, const double* d_pressure_hl,
                                                         const double* d_temperature_hl, const double* d_grad_od,
                                                         double* const* h_d_grad_molar_abs, int accumulate) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && h_d_vmr_fl && d_pressure_hl && d_temperature_hl && d_grad_od && h_d_grad_molar_abs && ncol > 0 &&
                    nlay > 0,
                ECCKD_PARAMETER_ERROR, "ecckd_calc_total_optical_depth_adjoint: bad argument");
  ECCKD_REQUIRE(m->ngas <= ADJ_MAXGAS, ECCKD_PARAMETER_ERROR, "ecckd_calc_total_optical_depth_adjoint: more than %d gases",
                ADJ_MAXGAS);
  const int64_t ncl = (int64_t)ncol * nlay;
  AdjParams p;
  memset(&p, 0, sizeof(p));
  int nacc = 0;
  for (int i = 0; i < m->ngas; ++i) {
    p.vmr[i] = h_d_vmr_fl[i];
    p.reference_vmr[i] = m->reference_vmr[i];
    p.log_vmr0[i] = m->log_vmr0[i];
    p.d_log_c[i] = m->d_log_c[i];
    p.conc_dependence[i] = m->conc_dependence[i];
    p.nconc[i] = m->nconc[i];
    // solve_adept.cpp:62-66: a gas the training scene does not hold is skipped unless it has no concentration dependence
    p.present[i] = (h_d_vmr_fl[i] || m->conc_dependence[i] == ECCKD_CONC_NONE) ? 1 : 0;
    p.grad[i] = p.present[i] ? h_d_grad_molar_abs[i] : nullptr;
    p.acc_off[i] = nacc;
    if (p.grad[i]) nacc += (m->conc_dependence[i] == ECCKD_CONC_LUT) ? m->nconc[i] : 1;
  }
  if (nacc == 0) return ECCKD_OK;
  p.ngas = m->ngas; p.nt = m->nt; p.np = m->np; p.ng = m->ng; p.nacc = nacc;
  p.log_p_0 = m->log_p_0; p.d_log_p = m->d_log_p; p.d_t = m->d_t; p.temperature = m->d_temperature;
  p.ncol = ncol; p.nlay = nlay; p.pressure_hl = d_pressure_hl; p.grad_od = d_grad_od; p.accumulate = accumulate ? 1 : 0;
  // scratch: layer temperatures, then the records
  const size_t b_t = ((size_t)ncl * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_key = ((size_t)ncl * sizeof(int) + 255) & ~(size_t)255;
  const size_t b_c = ((size_t)4 * ncl * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_w = ((size_t)m->ngas * ncl * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_i = ((size_t)m->ngas * ncl * sizeof(int) + 255) & ~(size_t)255;
  void* base = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, b_t + b_key + b_c + 2 * b_w + b_i, &base));
  char* c = static_cast<char*>(base);
  double* d_tfl = reinterpret_cast<double*>(c);
  p.temperature_fl = d_tfl;
  p.key = reinterpret_cast<int*>(c + b_t);
  p.corner = reinterpret_cast<double*>(c + b_t + b_key);
  p.w = reinterpret_cast<double*>(c + b_t + b_key + b_c);
  p.cw1 = reinterpret_cast<double*>(c + b_t + b_key + b_c + b_w);
  p.ic0 = reinterpret_cast<int*>(c + b_t + b_key + b_c + 2 * b_w);
  temperature_fl_kernel<<<(unsigned)((ncl + 255) / 256), 256, 0, ctx->stream>>>(ncol, nlay, d_temperature_hl, d_pressure_hl, d_tfl);
  ECCKD_LAUNCH_CHECK(ctx);
  od_adjoint_records_kernel<<<(unsigned)((ncl + 127) / 128), 128, 0, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  const size_t smem = (size_t)ADJ_WARPS * nacc * 32 * sizeof(double);
  ECCKD_REQUIRE(smem <= ctx->smem_optin, ECCKD_PARAMETER_ERROR,
                "ecckd_calc_total_optical_depth_adjoint: %d accumulators need %zu B of shared memory", nacc, smem);
  ECCKD_CUDA_OK(cudaFuncSetAttribute(od_adjoint_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  od_adjoint_gather_kernel<<<m->nt * m->np, ADJ_WARPS * 32, smem, ctx->stream>>>(p);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_ckd_calc_planck_function_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, int64_t n,
                                                 const double* d_temperature, double* d_planck) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && d_temperature && d_planck && n > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_calc_planck_function: bad argument");
  ECCKD_REQUIRE(!m->is_sw, ECCKD_PARAMETER_ERROR, "ecckd_ckd_calc_planck_function: shortwave model has no Planck table");
  const dim3 block = g_block(m->ng);
  const dim3 grid((unsigned)((n + block.y - 1) / block.y), (m->ng + block.x - 1) / block.x);
  ckd_planck_kernel<<<grid, block, 0, ctx->stream>>>(m->nlut, m->ng, m->planck_t0, m->planck_dt, m->d_planck_lut, n,
                                                     d_temperature, d_planck);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_ckd_calc_rayleigh_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, int ncol, int nlay,
                                                        const double* d_pressure_hl, double* d_od) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && d_pressure_hl && d_od && ncol > 0 && nlay > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_ckd_calc_rayleigh_optical_depth: bad argument");
  ECCKD_REQUIRE(m->is_sw, ECCKD_PARAMETER_ERROR, "ecckd_ckd_calc_rayleigh_optical_depth: longwave model");
  const dim3 block = g_block(m->ng);
  const int64_t ncl = (int64_t)ncol * nlay;
  const dim3 grid((unsigned)((ncl + block.y - 1) / block.y), (m->ng + block.x - 1) / block.x);
  rayleigh_od_kernel<<<grid, block, 0, ctx->stream>>>(ncol, nlay, m->ng, d_pressure_hl, m->d_rayleigh, d_od);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// d_bb_dn / d_bb_up (both or neither): the sums over the n spectral points, produced by the sweep
// itself when it runs as the staged kernel with n == 32; *bb_done tells the caller whether it did.
static int rt_lw_impl(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* d_planck, const double* d_od,
                      const double* d_surf_emissivity, const double* d_surf_planck, double* d_flux_dn, double* d_flux_up,
                      double* d_bb_dn, double* d_bb_up, bool* bb_done) {
  if (bb_done) *bb_done = false;
  ECCKD_REQUIRE(ctx && d_planck && d_od && d_surf_emissivity && d_surf_planck && d_flux_dn && d_flux_up,
                ECCKD_PARAMETER_ERROR, "ecckd_radiative_transfer_lw: NULL argument");
  const int bs = n >= 128 ? 128 : (n > 32 ? 64 : 32);
  const int64_t nblk = (int64_t)ncol * ((n + bs - 1) / bs);
  ECCKD_REQUIRE(ncol > 0 && nlay > 0 && n > 0 && nblk < 2147483647LL, ECCKD_PARAMETER_ERROR,
                "ecckd_radiative_transfer_lw: bad dimensions");
  const size_t staged_smem = (size_t)(2 * nlay + 1) * 256 + 16;
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_planck) | reinterpret_cast<uintptr_t>(d_od)) & 15u) == 0;
  if ((n & 31) == 0 && aligned && staged_smem <= ctx->smem_optin / 4 && !getenv("ECCKD_RT_PLAIN")) {
    RtLwStagedParams sp;
    sp.ncol = ncol; sp.nlay = nlay; sp.n = n; sp.planck = d_planck; sp.od = d_od; sp.surf_emissivity = d_surf_emissivity;
    sp.surf_planck = d_surf_planck; sp.flux_dn = d_flux_dn; sp.flux_up = d_flux_up;
    const bool fuse_bb = n == 32 && d_bb_dn && d_bb_up;
    sp.bb_dn = fuse_bb ? d_bb_dn : nullptr;
    sp.bb_up = fuse_bb ? d_bb_up : nullptr;
    if (bb_done) *bb_done = fuse_bb;
    trans_const_init(&sp.tc, ECCKD_LW_DIFFUSIVITY);
    const int per_sm = (int)std::min<size_t>(32, (ctx->smem_optin + 1024) / (staged_smem + 1024));
    const int64_t items = (int64_t)ncol * (n >> 5);
    const unsigned grid = (unsigned)std::min<int64_t>(items, (int64_t)ctx->num_sms * per_sm);
    if (n == 32) {
      ECCKD_CUDA_OK(cudaFuncSetAttribute(rt_lw_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)staged_smem));
      rt_lw_staged_kernel<true><<<grid, 32, staged_smem, ctx->stream>>>(sp);
    } else {
      ECCKD_CUDA_OK(cudaFuncSetAttribute(rt_lw_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)staged_smem));
      rt_lw_staged_kernel<false><<<grid, 32, staged_smem, ctx->stream>>>(sp);
    }
  } else {
    rt_lw_kernel<<<(unsigned)nblk, bs, 0, ctx->stream>>>(ncol, nlay, n, d_planck, d_od, d_surf_emissivity, d_surf_planck,
                                                         d_flux_dn, d_flux_up);
  }
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_radiative_transfer_lw_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* d_planck,
                                              const double* d_od, const double* d_surf_emissivity,
                                              const double* d_surf_planck, double* d_flux_dn, double* d_flux_up) {
  ECCKD_DEVICE(ctx);
  return rt_lw_impl(ctx, ncol, nlay, n, d_planck, d_od, d_surf_emissivity, d_surf_planck, d_flux_dn, d_flux_up, nullptr,
                    nullptr, nullptr);
}

ECCKD_API int ecckd_radiative_transfer_sw_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* h_cos_sza,
                                              const double* d_ssi, const double* d_od, const double* d_albedo,
                                              double* d_flux_dn, double* d_flux_up) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && h_cos_sza && d_ssi && d_od && d_flux_dn, ECCKD_PARAMETER_ERROR,
                "ecckd_radiative_transfer_sw: NULL argument");
  ECCKD_REQUIRE((d_albedo == nullptr) == (d_flux_up == nullptr), ECCKD_PARAMETER_ERROR,
                "ecckd_radiative_transfer_sw: albedo and flux_up must both be given or both be NULL");
  ECCKD_REQUIRE(ncol > 0 && nlay > 0 && n > 0 && (int64_t)ncol * ((n + 31) / 32) < 2147483647LL, ECCKD_PARAMETER_ERROR,
                "ecckd_radiative_transfer_sw: bad dimensions");
  void* d_mu = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)ncol * sizeof(double), &d_mu));
  ECCKD_TRY(upload_small(ctx, d_mu, h_cos_sza, (size_t)ncol * sizeof(double)));
  const size_t staged_smem = (size_t)nlay * 256 + 16;
  if ((n & 31) == 0 && (reinterpret_cast<uintptr_t>(d_od) & 15u) == 0 && staged_smem <= ctx->smem_optin / 4 &&
      !getenv("ECCKD_RT_PLAIN")) {
    RtSwStagedParams sp;
    sp.ncol = ncol; sp.nlay = nlay; sp.n = n; sp.cos_sza = static_cast<const double*>(d_mu); sp.ssi = d_ssi; sp.od = d_od;
    sp.albedo = d_albedo; sp.flux_dn = d_flux_dn; sp.flux_up = d_flux_up;
    const int per_sm = (int)std::min<size_t>(32, (ctx->smem_optin + 1024) / (staged_smem + 1024));
    const int64_t items = (int64_t)ncol * (n >> 5);
    const unsigned grid = (unsigned)std::min<int64_t>(items, (int64_t)ctx->num_sms * per_sm);
    ECCKD_CUDA_OK(cudaFuncSetAttribute(rt_sw_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
    rt_sw_staged_kernel<<<grid, 32, staged_smem, ctx->stream>>>(sp);
  } else {
    const int bs = n >= 128 ? 128 : (n > 32 ? 64 : 32);
    rt_sw_kernel<<<(unsigned)((int64_t)ncol * ((n + bs - 1) / bs)), bs, 0, ctx->stream>>>(
        ncol, nlay, n, static_cast<const double*>(d_mu), d_ssi, d_od, d_albedo, d_flux_dn, d_flux_up);
  }
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_heating_rate_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* d_pressure_hl,
                                     const double* d_flux_dn, const double* d_flux_up, double* d_hr) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_pressure_hl && d_flux_dn && d_hr, ECCKD_PARAMETER_ERROR, "ecckd_heating_rate: NULL argument");
  const int bs = n >= 128 ? 128 : (n > 32 ? 64 : 32);
  const int64_t nblk = (int64_t)ncol * nlay * ((n + bs - 1) / bs);
  ECCKD_REQUIRE(ncol > 0 && nlay > 0 && n > 0 && nblk < 2147483647LL, ECCKD_PARAMETER_ERROR,
                "ecckd_heating_rate: bad dimensions");
  heating_rate_kernel<<<(unsigned)nblk, bs, 0, ctx->stream>>>(
      ncol, nlay, n, d_pressure_hl, d_flux_dn, d_flux_up, d_hr);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_run_ckd_dev(ecckd_ctx* ctx, ecckd_ckd_model* m, const double* const* h_d_vmr_fl, double tsi,
                                int ncol, int nlay, const double* d_pressure_hl, const double* d_temperature_hl,
                                double* d_od, double* d_planck_hl, double* d_planck_surf, double* d_flux_dn,
                                double* d_flux_up, double* d_bb_flux_dn, double* d_bb_flux_up) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && d_pressure_hl && d_temperature_hl, ECCKD_PARAMETER_ERROR, "ecckd_run_ckd: NULL argument");
  ECCKD_REQUIRE(ncol > 0 && nlay > 0, ECCKD_PARAMETER_ERROR, "ecckd_run_ckd: bad dimensions");
  // The longwave sweep produces both fluxes together (radiative_transfer_lw.cpp:25-60 has no
  // one-sided form), and a broadband sum is the sum over g of a spectral flux that was asked for.
  ECCKD_REQUIRE(m->is_sw || ((d_flux_dn == nullptr) == (d_flux_up == nullptr)), ECCKD_PARAMETER_ERROR,
                "ecckd_run_ckd: longwave d_flux_dn and d_flux_up must both be given or both be NULL");
  ECCKD_REQUIRE((!d_bb_flux_dn || d_flux_dn) && (m->is_sw || !d_bb_flux_up || d_flux_up), ECCKD_PARAMETER_ERROR,
                "ecckd_run_ckd: a broadband flux needs its spectral flux array");
  std::vector<GasDev> gd(m->ngas);
  for (int i = 0; i < m->ngas; ++i) {
    const double* v = h_d_vmr_fl ? h_d_vmr_fl[i] : nullptr;
    ECCKD_TRY(check_vmr(m, i, v));
    gd[i] = make_gasdev(m, i, v);
  }
  ECCKD_TRY(upload_small(ctx, m->d_gasdesc, gd.data(), gd.size() * sizeof(GasDev)));
  RunCkdParams p;
  p.gases = static_cast<const GasDev*>(m->d_gasdesc);
  p.ngas = m->ngas;
  p.nt = m->nt;
  p.np = m->np;
  p.ng = m->ng;
  p.nlut = m->nlut;
  p.is_sw = m->is_sw ? 1 : 0;
  p.log_p_0 = m->log_p_0;
  p.d_log_p = m->d_log_p;
  p.d_t = m->d_t;
  p.planck_t0 = m->planck_t0;
  p.planck_dt = m->planck_dt;
  p.tsi_scaling = m->is_sw ? tsi / m->solar_sum : 0.0;  // run_ckd.cpp:98
  p.temperature = m->d_temperature;
  p.planck_lut = m->d_planck_lut;
  p.solar_irradiance = m->d_solar_irradiance;
  p.rayleigh = m->d_rayleigh;
  p.ncol = ncol;
  p.nlay = nlay;
  p.pressure_hl = d_pressure_hl;
  p.temperature_hl = d_temperature_hl;
  p.od = d_od;
  p.planck_hl = d_planck_hl;
  p.planck_surf = d_planck_surf;
  p.flux_dn = d_flux_dn;
  p.flux_up = m->is_sw ? nullptr : d_flux_up;
  p.bb_flux_dn = nullptr;  // broadband sums are done by sum_g_kernel below (fixed order)
  p.bb_flux_up = nullptr;
  p.scratch_od = nullptr;
  if (!d_od) {
    void* s = nullptr;
    ECCKD_TRY(ws_reserve(ctx, 1, (size_t)ncol * nlay * m->ng * sizeof(double), &s));
    p.scratch_od = static_cast<double*>(s);
  }
  const int threads = m->ng <= 32 ? 32 : (m->ng <= 64 ? 64 : 128);
  const size_t col_smem = run_ckd_col_smem(nlay, m->ngas);
  ECCKD_REQUIRE(m->ngas <= RUN_CKD_MAX_GAS && m->ngas <= threads, ECCKD_PARAMETER_ERROR,
                "ecckd_run_ckd: more than %d gases", std::min(RUN_CKD_MAX_GAS, threads));
  ECCKD_REQUIRE(col_smem <= ctx->smem_optin / 2, ECCKD_PARAMETER_ERROR,
                "ecckd_run_ckd: nlay*ngas too large for shared memory");
  // columns per block: as many resident warps per SM as shared memory allows, larger blocks on ties
  int cols = 1, best = 0;
  for (int c = std::min(256 / threads, ncol); c >= 1; --c) {
    if (c * col_smem > ctx->smem_optin / 2) continue;
    const int blocks = std::min<int>(32, (int)((ctx->smem_optin + 1024) / (c * col_smem + 1024)));
    const int warps = std::min(64, blocks * c * (threads / 32));
    if (warps > best) { best = warps; cols = c; }
  }
  ECCKD_CUDA_OK(cudaFuncSetAttribute(run_ckd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(col_smem * cols)));
  // Longwave fluxes: the kernel produces optical depth and Planck function (scratch arrays when the
  // caller does not want them) and the sweep is the staged flux solver, which reads both once.
  const bool lw_staged = !m->is_sw && d_flux_dn && d_flux_up;
  if (lw_staged) {
    const size_t n_hl = (size_t)ncol * (nlay + 1) * m->ng, n_s = (size_t)ncol * m->ng;
    if (!d_planck_hl || !d_planck_surf) {
      void* s = nullptr;
      ECCKD_TRY(ws_reserve(ctx, 0, (n_hl + n_s) * sizeof(double), &s));
      if (!d_planck_hl) p.planck_hl = static_cast<double*>(s);
      if (!d_planck_surf) p.planck_surf = static_cast<double*>(s) + n_hl;
    }
    p.flux_dn = nullptr;
    p.flux_up = nullptr;
  }
  p.vec2 = (m->ng % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_od ? d_od : p.scratch_od) & 15u) == 0);
  // warps per column (each takes every lz-th layer): fill the SM's warp slots
  // measured on the B200 (65536 x 54 x 32, whole run_ckd core): one g point per thread 7.1 / 6.4 / 4.6 / 4.2 ms for
  // 1 / 2 / 3 / 4 warps per column; two g points per thread 4.5 / 3.55 / 3.8 / 3.7 ms
  int lz = p.vec2 ? 2 : 4;
  while (lz > 1 && (threads * cols * lz > 1024 || lz > nlay)) --lz;
#ifdef ECCKD_TIMING_KNOBS  // tools/run_ckd_lz.sh
  if (const char* e = getenv("ECCKD_RUN_CKD_LZ")) lz = std::max(1, std::min(atoi(e), 1024 / (threads * cols)));
#endif
  const bool legacy_kernel = getenv("ECCKD_RUN_CKD_LEGACY") && atoi(getenv("ECCKD_RUN_CKD_LEGACY")) != 0;
  const bool layer_major = !m->is_sw && m->ngas <= TOD_MAXGAS && !legacy_kernel && (int64_t)nlay * m->ng * TOD_TILE < (1LL << 28);
  if (lw_staged && !m->d_ones) {  // surface emissivity 1 (run_ckd.cpp:344-345)
    std::vector<double> ones(m->ng, 1.0);
    ECCKD_CUDA_OK(cudaMalloc(&m->d_ones, m->ng * sizeof(double)));
    ECCKD_CUDA_OK(cudaMemcpyAsync(m->d_ones, ones.data(), m->ng * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  bool bb_done = false;
  double* const od_arr = d_od ? d_od : p.scratch_od;
  if (layer_major) {
    // longwave: optical depth and Planck function by the layer-major kernel (table rows in registers),
    // then the staged sweep.  (Running the two kernels over slices of columns sized for L2, so that the
    // sweep would read optical depth and Planck function there, was measured and rejected: 65536 x 54 x 32,
    // slices of 16 / 32 / 64 / 128 MB: 7.3 / 4.9 / 5.0 / 4.4 ms against 4.1 ms unsliced - the short
    // launches leave SMs idle and the sweep already runs at 89 % of the HBM peak.)
    TotalOdParams q = {};
    int order[TOD_MAXGAS], n = 0;
    for (int i = 0; i < m->ngas; ++i)
      if (gd[i].conc_dependence != ECCKD_CONC_LUT) order[n++] = i;
    q.nlin = n;
    for (int i = 0; i < m->ngas; ++i)
      if (gd[i].conc_dependence == ECCKD_CONC_LUT) order[n++] = i;
    for (int k = 0; k < m->ngas; ++k) {
      const int i = order[k];
      q.packed[k] = m->d_packed[i];
      q.vmr[k] = gd[i].vmr;
      q.reference_vmr[k] = gd[i].reference_vmr;
      q.log_vmr0[k] = gd[i].log_vmr0;
      q.d_log_c[k] = gd[i].d_log_c;
      q.conc_dependence[k] = gd[i].conc_dependence;
      q.nconc[k] = gd[i].nconc;
    }
    q.ngas = m->ngas;
    q.nt = m->nt;
    q.np = m->np;
    q.ng = m->ng;
    q.nlut = m->nlut;
    q.log_p_0 = m->log_p_0;
    q.d_log_p = m->d_log_p;
    q.d_t = m->d_t;
    q.planck_t0 = m->planck_t0;
    q.planck_dt = m->planck_dt;
    q.temperature = m->d_temperature;
    q.planck_lut = m->d_planck_lut;
    q.ncol = ncol;
    q.nlay = nlay;
    q.pressure_hl = d_pressure_hl;
    q.temperature_hl = d_temperature_hl;
    q.od = od_arr;
    q.planck_hl = p.planck_hl;
    q.planck_surf = p.planck_surf;
    // 4 warps x 3 blocks per SM: 168 registers per thread, nothing spilled.  (8 warps x 2 blocks at 128 registers
    // spills the prefetched table rows, which turns the prefetch into a blocking load: 1.68 against 1.44 ms for
    // 65536 x 54 x 32.)
    const int nwarp = 4, minb = 3;
    const int64_t ntile = (ncol + TOD_TILE - 1) / TOD_TILE;
    const bool want_planck = q.planck_hl || q.planck_surf;
    const bool has_lut = q.nlin < q.ngas;
    // Kernel choice (65536 x 54 x 32 on the B200, optical depth + Planck function): total_od_planck3_kernel 1.28 ms
    // (default); ECCKD_TOD_VARIANT=1 total_od_planck_kernel 1.45 ms (synchronous loads of the per-unit inputs,
    // separate Planck units); ECCKD_TOD_VARIANT=2 total_od_planck2_kernel 1.70-1.76 ms (everything through shared
    // memory: its long-scoreboard stalls are gone, 3.2 -> 1.3 per issue, but the per-column records and rows now
    // cost short-scoreboard stalls, 1.0 -> 3.2 per issue, with the same three warps per scheduler).
    const int tod_variant = getenv("ECCKD_TOD_VARIANT") ? atoi(getenv("ECCKD_TOD_VARIANT")) : 3;
    const bool v2 = tod_variant == 2;
    const bool v3 = tod_variant != 1 && tod_variant != 2;
    const int64_t units = (v2 || v3) ? ntile * nlay : ntile * nlay + (want_planck ? ntile * (nlay + 1) : 0);
    const unsigned grid = (unsigned)std::min<int64_t>((units + nwarp - 1) / nwarp, (int64_t)minb * ctx->num_sms);
    const size_t smem = v2 ? nwarp * sizeof(Tod2Smem) : (v3 ? nwarp * sizeof(Tod3Smem) : nwarp * TOD_TILE * sizeof(TodCol));
#define ECCKD_TOD_LAUNCH(N)                                                                                  \
  case N:                                                                                                    \
    if (v2) {                                                                                                \
      if (has_lut) {                                                                                         \
        ECCKD_CUDA_OK(cudaFuncSetAttribute(total_od_planck2_kernel<N, true, 4, 3>,                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        total_od_planck2_kernel<N, true, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);                  \
      } else {                                                                                               \
        ECCKD_CUDA_OK(cudaFuncSetAttribute(total_od_planck2_kernel<N, false, 4, 3>,                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        total_od_planck2_kernel<N, false, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);                 \
      }                                                                                                      \
    } else if (v3) {                                                                                         \
      if (has_lut) total_od_planck3_kernel<N, true, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);       \
      else total_od_planck3_kernel<N, false, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);              \
    } else if (has_lut) total_od_planck_kernel<N, true, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);   \
    else total_od_planck_kernel<N, false, 4, 3><<<grid, nwarp * 32, smem, ctx->stream>>>(q);                 \
    break;
    switch (q.nlin) {
      ECCKD_TOD_LAUNCH(0)
      ECCKD_TOD_LAUNCH(1)
      ECCKD_TOD_LAUNCH(2)
      ECCKD_TOD_LAUNCH(3)
      ECCKD_TOD_LAUNCH(4)
      ECCKD_TOD_LAUNCH(5)
      ECCKD_TOD_LAUNCH(6)
      ECCKD_TOD_LAUNCH(7)
      ECCKD_TOD_LAUNCH(8)
    }
#undef ECCKD_TOD_LAUNCH
    ECCKD_LAUNCH_CHECK(ctx);
    if (lw_staged) {
      ECCKD_TRY(rt_lw_impl(ctx, ncol, nlay, m->ng, p.planck_hl, od_arr, m->d_ones, p.planck_surf, d_flux_dn, d_flux_up,
                           d_bb_flux_dn, d_bb_flux_up, &bb_done));
    }
  } else {
    run_ckd_kernel<<<(ncol + cols - 1) / cols, dim3(threads, cols, lz), col_smem * cols, ctx->stream>>>(p);
    ECCKD_LAUNCH_CHECK(ctx);
    if (lw_staged) {
      ECCKD_TRY(rt_lw_impl(ctx, ncol, nlay, m->ng, p.planck_hl, od_arr, m->d_ones, p.planck_surf, d_flux_dn, d_flux_up,
                           d_bb_flux_dn, d_bb_flux_up, &bb_done));
    }
  }
  if (bb_done) return ECCKD_OK;
  const int64_t nrows = (int64_t)ncol * (nlay + 1);
  if (d_flux_dn && d_bb_flux_dn) {
    launch_sum_g(ctx, nrows, m->ng, d_flux_dn, d_bb_flux_dn);
    ECCKD_LAUNCH_CHECK(ctx);
  }
  if (!m->is_sw && d_flux_up && d_bb_flux_up) {
    launch_sum_g(ctx, nrows, m->ng, d_flux_up, d_bb_flux_up);
    ECCKD_LAUNCH_CHECK(ctx);
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_run_ckd(ecckd_ctx* ctx, ecckd_ckd_model* m, const double* const* vmr_fl, double tsi, int ncol,
                            int nlay, const double* pressure_hl, const double* temperature_hl, double* od,
                            double* planck_hl, double* planck_surf, double* flux_dn, double* flux_up,
                            double* bb_flux_dn, double* bb_flux_up) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && m && pressure_hl && temperature_hl, ECCKD_PARAMETER_ERROR, "ecckd_run_ckd: NULL argument");
  ECCKD_REQUIRE(ncol > 0 && nlay > 0, ECCKD_PARAMETER_ERROR, "ecckd_run_ckd: bad dimensions");
  const int ng = m->ng, ngas = m->ngas;
  const size_t n_hl = (size_t)ncol * (nlay + 1), n_fl = (size_t)ncol * nlay;
  // one device arena: inputs then outputs
  size_t n_in = 2 * n_hl + (size_t)ngas * n_fl;
  size_t n_out = n_fl * ng + n_hl * ng * 3 + (size_t)ncol * ng + 2 * n_hl;
  void* arena_v = nullptr;  // the context's host-API staging slot: kept between calls
  ECCKD_TRY(ws_reserve(ctx, 4, (n_in + n_out) * sizeof(double), &arena_v));
  double* arena = static_cast<double*>(arena_v);
  cudaStream_t st = ctx->stream;
  double* d_p = arena;
  double* d_t = d_p + n_hl;
  double* d_v = d_t + n_hl;
  double* d_od = d_v + (size_t)ngas * n_fl;
  double* d_plh = d_od + n_fl * ng;
  double* d_dn = d_plh + n_hl * ng;
  double* d_up = d_dn + n_hl * ng;
  double* d_pls = d_up + n_hl * ng;
  double* d_bdn = d_pls + (size_t)ncol * ng;
  double* d_bup = d_bdn + n_hl;
  int rc = ECCKD_OK;
  cudaError_t e = cudaMemcpyAsync(d_p, pressure_hl, n_hl * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, temperature_hl, n_hl * sizeof(double), cudaMemcpyHostToDevice, st);
  std::vector<const double*> dv(ngas, nullptr);
  for (int i = 0; i < ngas && e == cudaSuccess; ++i) {
    if (vmr_fl && vmr_fl[i]) {
      dv[i] = d_v + (size_t)i * n_fl;
      e = cudaMemcpyAsync(d_v + (size_t)i * n_fl, vmr_fl[i], n_fl * sizeof(double), cudaMemcpyHostToDevice, st);
    }
  }
  const bool want_flux = flux_dn != nullptr && (m->is_sw || flux_up != nullptr);
  if (e == cudaSuccess) {
    rc = ecckd_run_ckd_dev(ctx, m, dv.data(), tsi, ncol, nlay, d_p, d_t, d_od, m->is_sw ? nullptr : d_plh,
                           m->is_sw ? nullptr : d_pls, want_flux ? d_dn : nullptr,
                           (want_flux && !m->is_sw) ? d_up : nullptr, (want_flux && bb_flux_dn) ? d_bdn : nullptr,
                           (want_flux && bb_flux_up && !m->is_sw) ? d_bup : nullptr);
  }
  auto back = [&](double* h, const double* d, size_t n) {
    if (h && e == cudaSuccess && rc == ECCKD_OK) e = cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, st);
  };
  back(od, d_od, n_fl * ng);
  if (!m->is_sw) {
    back(planck_hl, d_plh, n_hl * ng);
    back(planck_surf, d_pls, (size_t)ncol * ng);
  }
  if (want_flux) {
    back(flux_dn, d_dn, n_hl * ng);
    if (!m->is_sw) back(flux_up, d_up, n_hl * ng);
    back(bb_flux_dn, d_bdn, n_hl);
    if (!m->is_sw) back(bb_flux_up, d_bup, n_hl);
  }
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc != ECCKD_OK) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) {
    set_error("ecckd_run_ckd: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  return ECCKD_OK;
}
