// ranking.cu - g-point ranking and assignment (reference: reorder_spectrum.cpp
// :111-300, single_gas_data.h:56-62, single_gas_data.cpp:23-124,
// find_g_points.cpp:1455-1481).
//
//  * sort keys: one thread per wavenumber walks the layers (coalesced across
//    wavenumbers), upward fluxes parked in shared memory [level][thread].
//  * rank: std::stable_sort(key[i] < key[j]) per band == a stable LSD radix
//    sort of an order-preserving 64-bit image of the key (with -0.0 folded onto
//    +0.0, which the comparator treats as equal).  8-bit digits, per-tile
//    histograms, one exclusive scan, stable in-tile multi-split with
//    __match_any_sync.  Passes whose digit is constant over the range are skipped.
//  * g-point assignment: integer per-wavenumber loops, tables in shared memory.
#include <algorithm>

#include "common.cuh"
#include "trans_core.cuh"

namespace ecckd {

__device__ __forceinline__ double planck_elem_r(double t, double nu, double dnu) {
  // planck_function.cpp:29-33,47-53
  const double h = 6.62606896e-34;
  const double c = 2.99792458e8;
  const double k = 1.3806504e-23;
  const double inv_cm_2_Hz = 100.0 * c;
  const double pi = 3.14159265358979323846;
  const double freq = nu * inv_cm_2_Hz;
  const double prefactor = (dnu * 2.0 * h * inv_cm_2_Hz * pi / (c * c)) * (freq * freq * freq);
  return prefactor / (exp((h / k) * (freq / t)) - 1.0);
}

__device__ __forceinline__ void lw_layer_r(double tau, double& e, double& f) {
  e = 1.0 - exp(-ECCKD_LW_DIFFUSIVITY * tau);
  f = (e > 1.0e-5) ? (1.0 - e * (1.0 / ECCKD_LW_DIFFUSIVITY) / tau) : (0.5 * e);
}

// Fast element forms for the sort-key kernel (the same schemes as the flux solver and the Planck
// weights of the spectral reduction): table exponential of trans_core.cuh, reciprocals refined by
// two Newton steps.  Results differ from the libm forms above by a few ulp.
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // ~20 bits
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double planck_elem_tab(double hk_over_t, double freq, double prefactor, const TransLane& tl) {
  // planck_function.cpp:47-53 with (h/k)/T formed on the host; exp argument capped where the result is 0 anyway
  const double e = exp_table(fmin(freq * hk_over_t, 700.0), tl) - 1.0;
  return prefactor * rcp_newton(e);
}
__device__ __forceinline__ void lw_layer_tab(double tau, const TransConst& K, const TransLane& tl, double& e, double& f) {
  e = trans_one_minus_exp(tau, K, tl);
  f = (e > 1.0e-5) ? (1.0 - (e * (1.0 / ECCKD_LW_DIFFUSIVITY)) * rcp_newton(tau)) : (0.5 * e);
}

// per-layer constants uploaded by the host: [0..nlay] (h/k)/T_hl, then conv[nlay],
// dh[nlay], dhh[nlay]  (reorder_spectrum.cpp:121-124, 178-183; heating_rate.h:38)
//
// One thread per wavenumber, three passes over the levels with two [level][thread] arrays in shared
// memory: (1) the Planck function of every level (independent: the exponentials pipeline), (2) the
// upward sweep, which parks the upward fluxes and replaces the Planck function of level l+1 by the
// layer's downward source B(l)(e-f) + B(l+1) f, (3) the downward sweep, which only needs the
// transmittance again (one table exponential), the heating rates and the two weighted sums.  Every
// exponential and reciprocal is evaluated once except the layer transmittance (twice); the optical
// depths are read twice, the second time out of L2.  All 32 lanes of a warp stay in the loops (the
// table exponential shuffles): threads past the end work on the last wavenumber and do not write.
template <int SK_G>
__global__ void sort_key_lw_kernel(int nlay, int64_t nwav, const double* __restrict__ lay, const double* __restrict__ wn,
                                   const double* __restrict__ dwn, const double* __restrict__ od, int64_t od_stride,
                                   double thr, TransConst K, double* __restrict__ key, double* __restrict__ column_od) {
  extern __shared__ double s_lw[];
  const int bs = blockDim.x, tx = threadIdx.x;
  double* s_up = s_lw;                            // [nlay+1][bs] upward flux
  double* s_b = s_lw + (size_t)(nlay + 1) * bs;   // [nlay+1][bs] Planck function, then downward sources
  const int64_t w_raw = (int64_t)blockIdx.x * bs + tx;
  const bool valid = w_raw < nwav;
  const int64_t w = valid ? w_raw : nwav - 1;
  const TransLane tl = trans_lane_init();
  const double* hkt = lay;
  const double* conv = lay + (nlay + 1);
  const double* dh = conv + nlay;
  const double* dhh = dh + nlay;
  const double nu = wn[w], dnu = dwn[w];
  // The optical depths of SK_G layers are loaded together, one group ahead of their use (with one
  // load in flight per warp the kernel was bound by memory latency: 10 ms for 54 x 2^23), and the
  // layer terms of a group are evaluated before its serial flux recurrence (independent chains).
  double nxt[SK_G];
  auto load_up = [&](int l0) {  // layers l0, l0-1, ...
#pragma unroll
    for (int j = 0; j < SK_G; ++j) nxt[j] = (l0 - j >= 0) ? od[(int64_t)(l0 - j) * od_stride + w] : 0.0;
  };
  auto load_dn = [&](int l0) {  // layers l0, l0+1, ...
#pragma unroll
    for (int j = 0; j < SK_G; ++j) nxt[j] = (l0 + j < nlay) ? od[(int64_t)(l0 + j) * od_stride + w] : 0.0;
  };
  load_up(nlay - 1);
  {
    const double h = 6.62606896e-34, c = 2.99792458e8, pi = 3.14159265358979323846;
    const double inv_cm_2_Hz = 100.0 * c;
    const double freq = nu * inv_cm_2_Hz;
    const double prefactor = (dnu * 2.0 * h * inv_cm_2_Hz * pi / (c * c)) * (freq * freq * freq);
#pragma unroll 4
    for (int l = 0; l <= nlay; ++l) s_b[l * bs + tx] = planck_elem_tab(hkt[l], freq, prefactor, tl);
  }
  // upward sweep: with surface emissivity 1 (reorder_spectrum.cpp:138-139)
  // flux_up(nlay) = surf_planck*1 + (1-1)*flux_dn(nlay), independent of flux_dn
  double f_up = s_b[nlay * bs + tx];
  s_up[nlay * bs + tx] = f_up;
  for (int l0 = nlay - 1; l0 >= 0; l0 -= SK_G) {
    double tau[SK_G], e[SK_G], f[SK_G];
#pragma unroll
    for (int j = 0; j < SK_G; ++j) tau[j] = nxt[j];
    if (l0 - SK_G >= 0) load_up(l0 - SK_G); else load_dn(0);
#pragma unroll
    for (int j = 0; j < SK_G; ++j) lw_layer_tab(tau[j], K, tl, e[j], f[j]);
#pragma unroll
    for (int j = 0; j < SK_G; ++j) {
      const int l = l0 - j;
      if (l >= 0) {  // warp-uniform
        const double b_bot = s_b[(l + 1) * bs + tx], b_t = s_b[l * bs + tx];
        f_up = f_up * (1.0 - e[j]) + b_bot * (e[j] - f[j]) + b_t * f[j];
        s_up[l * bs + tx] = f_up;
        s_b[(l + 1) * bs + tx] = b_t * (e[j] - f[j]) + b_bot * f[j];  // downward source of layer l (level l+1 is done with)
      }
    }
  }
  // downward sweep, heating rate and the two weighted sums in ascending layer order
  double f_dn = 0.0, col = 0.0, num = 0.0, den = 0.0;
  for (int l0 = 0; l0 < nlay; l0 += SK_G) {
    double tau[SK_G], e[SK_G];
#pragma unroll
    for (int j = 0; j < SK_G; ++j) tau[j] = nxt[j];
    if (l0 + SK_G < nlay) load_dn(l0 + SK_G);
#pragma unroll
    for (int j = 0; j < SK_G; ++j) e[j] = trans_one_minus_exp(tau[j], K, tl);
#pragma unroll
    for (int j = 0; j < SK_G; ++j) {
      const int l = l0 + j;
      if (l < nlay) {  // warp-uniform
        col += tau[j];
        const double f_next = f_dn * (1.0 - e[j]) + s_b[(l + 1) * bs + tx];
        double hr = conv[l] * (f_next - f_dn - s_up[(l + 1) * bs + tx] + s_up[l * bs + tx]);
        if (hr > 0.0) hr = 0.0;  // reorder_spectrum.cpp:175
        num += hr * dhh[l];
        den += hr * dh[l];
        f_dn = f_next;
      }
    }
  }
  double k = num / den;
  if (thr > 0.0 && col < thr) k = -thr + col;  // reorder_spectrum.cpp:187-190
  if (valid) {
    key[w] = k;
    if (column_od) column_od[w] = col;
  }
}

// reorder_spectrum.cpp:197-228; lay = pseudo_height_hl[nlay+1]
__global__ void sort_key_sw_kernel(int nlay, int64_t nwav, const double* __restrict__ h_hl, const double* __restrict__ od,
                                   int64_t od_stride, double thr, double* __restrict__ key,
                                   double* __restrict__ column_od, int* __restrict__ err) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  double col = 0.0;
  for (int l = 0; l < nlay; ++l) col += od[(int64_t)l * od_stride + w];
  if (column_od) column_od[w] = col;
  double k;
  if (col <= thr) {
    k = col - thr;
  } else {
    double cum = 0.0;
    k = 0.0;
    for (int l = 0; l < nlay; ++l) {
      const double tau = od[(int64_t)l * od_stride + w];
      const double next = cum + tau;
      if (next >= thr) {
        k = ((thr - cum) * h_hl[l + 1] + (next - thr) * h_hl[l]) / fmax(1.0e-12, tau);
        if (k > 30.0) *err = 1;  // the reference throws here (:214-216)
        break;
      }
      cum = next;
    }
  }
  key[w] = k;
}

// ------------------------------------------------------------------ bands
__global__ void band_kernel(int64_t nwav, const double* __restrict__ wn, int nband, const double* __restrict__ b1,
                            const double* __restrict__ b2, int32_t* __restrict__ iband, int* __restrict__ first,
                            int* __restrict__ last, int32_t* __restrict__ g_index) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const double nu = wn[w];
  int ib = -1;
  for (int j = 0; j < nband; ++j) {
    const bool in = (j < nband - 1) ? (nu >= b1[j] && nu < b2[j]) : (nu >= b1[j] && nu <= b2[j]);
    if (in) {
      ib = j;  // later bands overwrite (iband(index) = jband, reorder_spectrum.cpp:290)
      atomicMin(&first[j], (int)w);
      atomicMax(&last[j], (int)w);
    }
  }
  if (iband) iband[w] = ib;
  g_index[w] = (int32_t)w;
}

// order-preserving image of a double under `a < b`, with -0.0 == +0.0
__device__ __forceinline__ unsigned long long key_image(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  if ((b << 1) == 0ull) b = 0ull;  // fold -0.0 onto +0.0
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void gather_keys_kernel(int64_t n, const int32_t* __restrict__ g_index, const double* __restrict__ key,
                                   unsigned long long* __restrict__ k_out, uint32_t* __restrict__ v_out,
                                   unsigned long long* __restrict__ or_bits, unsigned long long* __restrict__ and_bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long img = 0ull, img_and = ~0ull;
  if (i < n) {
    const int32_t idx = g_index[i];
    img = key_image(key[idx]);
    img_and = img;
    k_out[i] = img;
    v_out[i] = (uint32_t)idx;
  }
  // which bits vary over the range: OR and AND reductions
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    img |= __shfl_xor_sync(0xffffffffu, img, o);
    img_and &= __shfl_xor_sync(0xffffffffu, img_and, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(or_bits, img);
    atomicAnd(and_bits, img_and);
  }
}

// All bands of a call at once (their index ranges [first, last] disjoint, sorted by position):
// element i gets the image of its key if it lies in a band and 0 otherwise, the value i, and the
// segment seg[i] = 2k+1 inside the k-th band, 2k in the gap before it.  After the stable sort by key
// image and a last stable pass by segment every band is sorted within its own range and the
// elements of the gaps have not moved.
__global__ void gather_all_kernel(int64_t n, const double* __restrict__ key, int nseg_bands,
                                  const int* __restrict__ first, const int* __restrict__ last,
                                  unsigned long long* __restrict__ k_out, uint32_t* __restrict__ v_out,
                                  uint8_t* __restrict__ seg, unsigned long long* __restrict__ or_bits,
                                  unsigned long long* __restrict__ and_bits) {
  __shared__ int s_first[128], s_last[128];
  for (int j = threadIdx.x; j < nseg_bands; j += blockDim.x) {
    s_first[j] = first[j];
    s_last[j] = last[j];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long img = 0ull, img_and = ~0ull;
  if (i < n) {
    int lo = 0, hi = nseg_bands;  // number of bands with first <= i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_first[mid] <= (int)i) lo = mid + 1; else hi = mid;
    }
    const bool inside = lo > 0 && (int)i <= s_last[lo - 1];
    seg[i] = (uint8_t)(inside ? 2 * (lo - 1) + 1 : 2 * lo);
    if (inside) {
      img = key_image(key[i]);
      img_and = img;
    }
    k_out[i] = img;
    v_out[i] = (uint32_t)i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    img |= __shfl_xor_sync(0xffffffffu, img, o);
    img_and &= __shfl_xor_sync(0xffffffffu, img_and, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (img) atomicOr(or_bits, img);
    if (~img_and) atomicAnd(and_bits, img_and);
  }
}

// g_index[i] = sorted value, rank[sorted value] = i (reorder_spectrum.cpp:296-300)
__global__ void order_and_rank_kernel(int64_t n, const uint32_t* __restrict__ vals, int32_t* __restrict__ g_index,
                                      int32_t* __restrict__ rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = vals[i];
  if (g_index) g_index[i] = (int32_t)v;
  if (rank) rank[v] = (int32_t)i;
}

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;  // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WSPAN = 32 * RS_ITEMS;  // a warp owns this many CONSECUTIVE elements of the tile
// pass digit: bits [shift, shift+8) of the key image, or (shift == RS_SEG_PASS) the segment of the
// element's original index (the bands of one ranking call sorted together, see ecckd_rank_bands_dev)
constexpr int RS_SEG_PASS = 64;

__device__ __forceinline__ int rs_digit(unsigned long long k, uint32_t v, int shift, const uint8_t* __restrict__ seg) {
  return (shift == RS_SEG_PASS) ? (int)seg[v] : (int)((k >> shift) & 0xFFull);
}

// per-tile digit histogram -> hist[digit][tile].  Per-warp histograms in shared memory (atomics only
// contend inside a warp); a warp whose 32 digits agree (smooth keys, constant high digits) adds once.
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(int64_t n, const unsigned long long* __restrict__ keys,
                                                                const uint32_t* __restrict__ vals,
                                                                const uint8_t* __restrict__ seg, int shift, int ntiles,
                                                                uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[RS_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) s_h[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  uint32_t* h = s_h[warp];
  auto count = [&](int digit) {  // digit < 0: out of range
    const int d0 = __shfl_sync(0xffffffffu, digit, 0);
    if (__all_sync(0xffffffffu, digit == d0)) {
      if (lane == 0 && d0 >= 0) h[d0] += 32u;
      __syncwarp();
    } else {
      if (digit >= 0) atomicAdd(&h[digit], 1u);
      __syncwarp();
    }
  };
  if (shift != RS_SEG_PASS && base + RS_TILE <= n) {  // full tile: two keys per load
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys + base);
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS / 2; ++r) {
      const ulonglong2 kk = k2[r * RS_THREADS + threadIdx.x];
      count((int)((kk.x >> shift) & 0xFFull));
      count((int)((kk.y >> shift) & 0xFFull));
    }
  } else {
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = base + r * RS_THREADS + threadIdx.x;
      count((i < n) ? rs_digit(keys[i], shift == RS_SEG_PASS ? vals[i] : 0u, shift, seg) : -1);
    }
  }
  __syncthreads();
  uint32_t t = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) t += s_h[w][threadIdx.x];
  hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = t;
}

// block d: exclusive scan of hist[d][0..ntiles) in place, the digit's total to totals[d]
__global__ void __launch_bounds__(1024) radix_scan_kernel(int ntiles, uint32_t* __restrict__ hist,
                                                          uint32_t* __restrict__ totals) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  uint32_t* h = hist + (size_t)blockIdx.x * ntiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < ntiles) ? h[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t t = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_warp[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t warp_off = (warp == 0) ? 0u : s_warp[warp - 1];
    if (i < ntiles) h[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

// exclusive scan of 256 values, one per thread of a 256-thread block (s_w: 8 words of scratch)
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  uint32_t off = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) off += (w < warp) ? s_w[w] : 0u;
  __syncthreads();
  return off + x - v;
}

// Stable scatter of one tile.  A warp owns RS_WSPAN consecutive elements and ranks them round by round
// (__match_any_sync + a per-warp running count per digit, warp-synchronous); the tile is then put in
// digit order in shared memory and written out in runs of consecutive addresses.
struct RsSmem {
  unsigned long long key[RS_TILE];
  uint32_t val[RS_TILE];
  uint32_t wcnt[RS_WARPS][257];  // per warp: running count per digit, then exclusive prefix over warps
  uint32_t dstart[256];          // first position of the digit in the sorted tile
  uint32_t gofs[256];            // global position of sorted-tile position p with digit d: gofs[d] + p
  uint32_t scratch[RS_WARPS];
  uint8_t dig[RS_TILE];
};

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(int64_t n, const unsigned long long* __restrict__ keys,
                                                                   const uint32_t* __restrict__ vals,
                                                                   const uint8_t* __restrict__ seg, int shift,
                                                                   int ntiles, const uint32_t* __restrict__ offs,
                                                                   const uint32_t* __restrict__ totals,
                                                                   unsigned long long* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  RsSmem& S = *reinterpret_cast<RsSmem*>(rs_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < RS_WARPS * 257; j += RS_THREADS) (&S.wcnt[0][0])[j] = 0;
  // global base of every digit: exclusive scan of the digit totals + this tile's offset within the digit
  const uint32_t dbase = block_exclusive_scan_256(totals[threadIdx.x], S.scratch) +
                         offs[(size_t)threadIdx.x * ntiles + blockIdx.x];
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WSPAN;
  unsigned long long k[RS_ITEMS];
  uint32_t v[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = base + r * 32 + lane;
    k[r] = 0ull;
    v[r] = 0u;
    if (i < n) {
      k[r] = keys[i];
      v[r] = vals[i];
    }
  }
  uint32_t dr[RS_ITEMS];  // digit | rank within (warp, digit) << 16
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = base + r * 32 + lane;
    const int digit = (i < n) ? rs_digit(k[r], v[r], shift, seg) : 256;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t before = (uint32_t)__popc(peers & ((1u << lane) - 1u));
    const uint32_t old = S.wcnt[warp][digit];
    __syncwarp();
    if (before == 0u) S.wcnt[warp][digit] = old + (uint32_t)__popc(peers);
    __syncwarp();
    dr[r] = (uint32_t)digit | ((old + before) << 16);
  }
  __syncthreads();
  uint32_t tile_count;
  {  // digit = threadIdx.x: exclusive prefix over warps; then the digit's start in the sorted tile
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = S.wcnt[w][threadIdx.x];
      S.wcnt[w][threadIdx.x] = run;
      run += c;
    }
    tile_count = run;
  }
  const uint32_t dstart = block_exclusive_scan_256(tile_count, S.scratch);
  S.dstart[threadIdx.x] = dstart;
  S.gofs[threadIdx.x] = dbase - dstart;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const uint32_t digit = dr[r] & 0xFFFFu;
    if (digit < 256u) {
      const uint32_t p = S.dstart[digit] + S.wcnt[warp][digit] + (dr[r] >> 16);
      S.key[p] = k[r];
      S.val[p] = v[r];
      S.dig[p] = (uint8_t)digit;
    }
  }
  __syncthreads();
  const int64_t tile_n = min((int64_t)RS_TILE, n - (int64_t)blockIdx.x * RS_TILE);
#pragma unroll 4
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int p = r * RS_THREADS + (int)threadIdx.x;
    if (p < tile_n) {
      const uint32_t pos = S.gofs[S.dig[p]] + (uint32_t)p;
      keys_out[pos] = S.key[p];
      vals_out[pos] = S.val[p];
    }
  }
}

__global__ void write_back_kernel(int64_t n, const uint32_t* __restrict__ vals, int32_t* __restrict__ g_index) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g_index[i] = (int32_t)vals[i];
}

__global__ void rank_from_order_kernel(int64_t nwav, const int32_t* __restrict__ g_index, int32_t* __restrict__ rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nwav) rank[g_index[i]] = (int32_t)i;  // reorder_spectrum.cpp:299-300
}

// ------------------------------------------------------------------ g points
__global__ void store_g_points_kernel(int64_t nwav, const int32_t* __restrict__ rank, int ng,
                                      const int32_t* __restrict__ r12, int32_t* __restrict__ g_point) {
  extern __shared__ int32_t s_r[];  // rank1[ng], rank2[ng]
  for (int j = threadIdx.x; j < 2 * ng; j += blockDim.x) s_r[j] = r12[j];
  __syncthreads();
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const int32_t r = rank[w];
  int g = -1;
  for (int ig = 0; ig < ng; ++ig) {
    if (r >= s_r[ig] && r <= s_r[ng + ig]) g = ig;  // later ig overwrites (single_gas_data.h:59-61)
  }
  g_point[w] = g;
}

__global__ void assign_g_points_kernel(int ngas, int64_t nwav, const int32_t* __restrict__ gas_g, int ng, int cap,
                                       const int32_t* __restrict__ gmm, int32_t* __restrict__ g_point,
                                       unsigned long long* __restrict__ n_unassigned) {
  extern __shared__ int32_t s_t[];  // g_min[ngas][ng], g_max[ngas][ng]
  for (int j = threadIdx.x; j < 2 * ngas * ng; j += blockDim.x) {
    const int which = j / (ngas * ng), rem = j - which * ngas * ng;
    const int igas = rem / ng, ig = rem - igas * ng;
    s_t[j] = gmm[(size_t)which * ngas * cap + (size_t)igas * cap + ig];
  }
  __syncthreads();
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  int g = -1;
  for (int ig = 0; ig < ng; ++ig) {
    bool found = true;
    for (int igas = 0; igas < ngas; ++igas) {
      const int32_t v = gas_g[(size_t)igas * nwav + w];
      if (v < s_t[igas * ng + ig] || v > s_t[ngas * ng + igas * ng + ig]) found = false;
    }
    if (found) g = ig;  // ascending ig, later overwrites (find_g_points.cpp:1463-1470)
  }
  g_point[w] = g;
  if (g < 0) atomicAdd(n_unassigned, 1ull);
}

// Stable LSD radix sort of n (key image, value) pairs in place in (k0, v0), using (k1, v1) as the
// ping-pong buffers; `vary` = key bits that differ somewhere (passes over constant digits are
// skipped); with `seg` a last pass orders by seg[value] (the segments of a multi-band call).
// hist: 256*ntiles + 256 words.
static int radix_sort_pairs(ecckd_ctx* ctx, int64_t n, unsigned long long* k0, uint32_t* v0, unsigned long long* k1,
                            uint32_t* v1, uint32_t* hist, unsigned long long vary, const uint8_t* seg,
                            uint32_t** v_result) {
  const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  uint32_t* totals = hist + (size_t)256 * ntiles;
  unsigned long long* kin = k0;
  unsigned long long* kout = k1;
  uint32_t* vin = v0;
  uint32_t* vout = v1;
  ECCKD_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(RsSmem)));
  for (int pass = 0; pass < 9; ++pass) {
    const int shift = pass * 8;  // pass 8 = RS_SEG_PASS
    if (pass < 8 && ((vary >> shift) & 0xFFull) == 0ull) continue;  // digit constant over the range: identity pass
    if (pass == 8 && !seg) continue;
    radix_hist_kernel<<<ntiles, RS_THREADS, 0, ctx->stream>>>(n, kin, vin, seg, shift, ntiles, hist);
    ECCKD_LAUNCH_CHECK(ctx);
    radix_scan_kernel<<<256, 1024, 0, ctx->stream>>>(ntiles, hist, totals);
    ECCKD_LAUNCH_CHECK(ctx);
    radix_scatter_kernel<<<ntiles, RS_THREADS, sizeof(RsSmem), ctx->stream>>>(n, kin, vin, seg, shift, ntiles, hist,
                                                                              totals, kout, vout);
    ECCKD_LAUNCH_CHECK(ctx);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  *v_result = vin;
  return ECCKD_OK;
}

}  // namespace ecckd

using namespace ecckd;

ECCKD_API int ecckd_sort_key_lw_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, const double* h_pressure_hl,
                                    const double* d_wn, const double* d_dwn, const double* d_od, int64_t od_stride,
                                    double thr, double* d_key, double* d_column_od) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && h_pressure_hl && d_wn && d_dwn && d_od && d_key, ECCKD_PARAMETER_ERROR,
                "ecckd_sort_key_lw: NULL argument");
  ECCKD_REQUIRE(nlay > 0 && nlay <= 400 && nwav > 0 && od_stride >= nwav, ECCKD_PARAMETER_ERROR,
                "ecckd_sort_key_lw: bad dimensions");
  // per-layer constants (reorder_spectrum.cpp:119-124,178-183; heating_rate.h:38)
  std::vector<double> lay((size_t)(nlay + 1) + 3 * (size_t)nlay);
  const double x0 = log(1.0), x1 = log(100000.0), y0 = 273.15 - 100.0, y1 = 273.15 + 15.0;
  for (int i = 0; i <= nlay; ++i) {
    const double x = log(h_pressure_hl[i]);
    const double t = (x <= x0) ? y0 : (x >= x1) ? y1 : y0 + (y1 - y0) * (x - x0) / (x1 - x0);
    lay[i] = (6.62606896e-34 / 1.3806504e-23) / t;  // (h/k)/T_hl, planck_function.cpp:29-33
  }
  double* conv = lay.data() + (nlay + 1);
  double* dh = conv + nlay;
  double* dhh = dh + nlay;
  for (int i = 0; i < nlay; ++i) {
    conv[i] = -(ECCKD_ACCEL_GRAVITY / ECCKD_SPECIFIC_HEAT_AIR) / (h_pressure_hl[i + 1] - h_pressure_hl[i]);
    const double ph = log(h_pressure_hl[nlay]) - 0.5 * (log(h_pressure_hl[i]) + log(h_pressure_hl[i + 1]));
    dh[i] = log(h_pressure_hl[i + 1]) - log(h_pressure_hl[i]);
    dhh[i] = dh[i] * ph;
  }
  void* d_lay = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, lay.size() * sizeof(double), &d_lay));
  ECCKD_TRY(upload_small(ctx, d_lay, lay.data(), lay.size() * sizeof(double)));
  // two [nlay+1][bs] arrays in shared memory: two blocks of 128 threads per SM at 54 layers
  int bs = 128;
  while (bs > 32 && 2 * (size_t)(nlay + 1) * bs * sizeof(double) > 110 * 1024) bs >>= 1;
  const size_t smem = 2 * (size_t)(nlay + 1) * bs * sizeof(double);
  ECCKD_REQUIRE(smem <= ctx->smem_optin, ECCKD_PARAMETER_ERROR, "ecckd_sort_key_lw: too many layers for shared memory");
  TransConst K;
  trans_const_init(&K, ECCKD_LW_DIFFUSIVITY);
  // layers per group (loads in flight and independent evaluation chains per thread; 54 x 2^23: 6.27 ms with 6,
  // 5.82 with 9, 5.69 with 18): the deepest group that wastes at most a tenth of the evaluations on padding
  int skg = 6;
  for (const int g : {18, 9}) {
    if (((nlay + g - 1) / g) * g - nlay <= nlay / 10) {
      skg = g;
      break;
    }
  }
  if (const char* env = getenv("ECCKD_SK_G")) skg = atoi(env);
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)((nwav + bs - 1) / bs), bs, smem, ctx->stream>>>(
        nlay, nwav, static_cast<const double*>(d_lay), d_wn, d_dwn, d_od, od_stride, thr, K, d_key, d_column_od);
    return cudaSuccess;
  };
  if (skg == 9) ECCKD_CUDA_OK(launch(sort_key_lw_kernel<9>));
  else if (skg == 18) ECCKD_CUDA_OK(launch(sort_key_lw_kernel<18>));
  else if (skg == 3) ECCKD_CUDA_OK(launch(sort_key_lw_kernel<3>));
  else ECCKD_CUDA_OK(launch(sort_key_lw_kernel<6>));
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_sort_key_sw_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, const double* h_pressure_hl,
                                    const double* d_od, int64_t od_stride, double thr, double* d_key,
                                    double* d_column_od) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && h_pressure_hl && d_od && d_key, ECCKD_PARAMETER_ERROR, "ecckd_sort_key_sw: NULL argument");
  ECCKD_REQUIRE(nlay > 0 && nwav > 0 && od_stride >= nwav, ECCKD_PARAMETER_ERROR, "ecckd_sort_key_sw: bad dimensions");
  std::vector<double> h_hl(nlay + 1);
  for (int i = 0; i <= nlay; ++i) h_hl[i] = log(h_pressure_hl[nlay]) - log(h_pressure_hl[i]);
  void* d_lay = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)(nlay + 2) * sizeof(double), &d_lay));
  ECCKD_TRY(upload_small(ctx, d_lay, h_hl.data(), h_hl.size() * sizeof(double)));
  int* d_err = reinterpret_cast<int*>(static_cast<double*>(d_lay) + (nlay + 1));
  ECCKD_CUDA_OK(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  sort_key_sw_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(
      nlay, nwav, static_cast<const double*>(d_lay), d_od, od_stride, thr, d_key, d_column_od, d_err);
  ECCKD_LAUNCH_CHECK(ctx);
  int err = 0;
  ECCKD_CUDA_OK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (err) {
    set_error("sorting variable exceeds 30 (reorder_spectrum.cpp:214-216)");
    return ECCKD_PROCESSING_ERROR;
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_rank_bands_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wn, const double* d_key, int nband,
                                   const double* h_b1, const double* h_b2, int32_t* d_iband, int32_t* d_rank,
                                   int32_t* d_ordered_index) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_wn && d_key && h_b1 && h_b2, ECCKD_PARAMETER_ERROR, "ecckd_rank_bands: NULL argument");
  ECCKD_REQUIRE(nwav > 0 && nwav < (int64_t)1 << 31, ECCKD_PARAMETER_ERROR, "ecckd_rank_bands: nwav out of range");
  // reorder_spectrum.cpp:245-248
  ECCKD_REQUIRE(nband > 0, ECCKD_PARAMETER_ERROR,
                "Failure to interpret wavenumber1 and wavenumber2 as a list of band boundaries");
  // scratch layout (slot 1): k0 | k1 | v0 | v1 | g_index | seg | hist (+ digit totals) | band meta
  const int ntiles_max = (int)((nwav + RS_TILE - 1) / RS_TILE);
  const size_t kbytes = ((size_t)nwav * sizeof(unsigned long long) + 255) & ~size_t(255);  // 16-byte loads of k1
  const size_t vbytes = ((size_t)nwav * sizeof(uint32_t) + 255) & ~size_t(255);
  const size_t sbytes = ((size_t)nwav + 255) & ~size_t(255);
  const size_t hbytes = ((size_t)256 * (ntiles_max + 1) * sizeof(uint32_t) + 255) & ~size_t(255);
  const size_t mbytes = ((size_t)nband * (2 * sizeof(double) + 4 * sizeof(int)) + 2 * sizeof(unsigned long long) + 255) &
                        ~size_t(255);
  void* base = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, 2 * kbytes + 3 * vbytes + sbytes + hbytes + mbytes + 1024, &base));
  char* pch = static_cast<char*>(base);
  unsigned long long* k0 = reinterpret_cast<unsigned long long*>(pch);
  unsigned long long* k1 = reinterpret_cast<unsigned long long*>(pch + kbytes);
  pch += 2 * kbytes;
  uint32_t* v0 = reinterpret_cast<uint32_t*>(pch);
  uint32_t* v1 = reinterpret_cast<uint32_t*>(pch + vbytes);
  int32_t* g_index = d_ordered_index ? d_ordered_index : reinterpret_cast<int32_t*>(pch + 2 * vbytes);
  pch += 3 * vbytes;
  uint8_t* d_seg = reinterpret_cast<uint8_t*>(pch);
  pch += sbytes;
  uint32_t* hist = reinterpret_cast<uint32_t*>(pch);
  pch += hbytes;
  double* d_b1 = reinterpret_cast<double*>(pch);
  double* d_b2 = d_b1 + nband;
  unsigned long long* d_bits = reinterpret_cast<unsigned long long*>(d_b2 + nband);  // [or, and]
  int* d_first = reinterpret_cast<int*>(d_bits + 2);
  int* d_last = d_first + nband;
  int* d_sfirst = d_last + nband;  // the bands' ranges sorted by position (one-shot path)
  int* d_slast = d_sfirst + nband;

  ECCKD_TRY(upload_small(ctx, d_b1, h_b1, (size_t)nband * sizeof(double)));
  ECCKD_TRY(upload_small(ctx, d_b2, h_b2, (size_t)nband * sizeof(double)));
  std::vector<int> init(2 * nband);
  for (int j = 0; j < nband; ++j) {
    init[j] = 0x7fffffff;
    init[nband + j] = -1;
  }
  ECCKD_TRY(upload_small(ctx, d_first, init.data(), init.size() * sizeof(int)));
  const unsigned nb = (unsigned)((nwav + 255) / 256);
  band_kernel<<<nb, 256, 0, ctx->stream>>>(nwav, d_wn, nband, d_b1, d_b2, d_iband, d_first, d_last, g_index);
  ECCKD_LAUNCH_CHECK(ctx);
  std::vector<int> fl(2 * nband);
  ECCKD_CUDA_OK(cudaMemcpyAsync(fl.data(), d_first, fl.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < nband; ++j) {
    const int i1 = fl[j], i2 = fl[nband + j];
    // the reference indexes an empty vector here (reorder_spectrum.cpp:291)
    ECCKD_REQUIRE(i2 >= 0 && i1 <= i2, ECCKD_PARAMETER_ERROR, "band %d (%g-%g cm-1) contains no wavenumbers", j, h_b1[j],
                  h_b2[j]);
  }
  const unsigned long long bits_init[2] = {0ull, ~0ull};
  // Bands with pairwise disjoint index ranges (the usual case: ascending wavenumbers, adjoining
  // bands) are sorted in ONE radix sort of the whole spectrum with the band as the last digit.
  std::vector<std::pair<int, int>> ranges(nband);
  for (int j = 0; j < nband; ++j) ranges[j] = std::make_pair(fl[j], fl[nband + j]);
  std::sort(ranges.begin(), ranges.end());
  bool disjoint = nband <= 120;
  for (int j = 1; j < nband && disjoint; ++j) disjoint = ranges[j].first > ranges[j - 1].second;
  if (disjoint) {
    std::vector<int> sr(2 * nband);
    for (int j = 0; j < nband; ++j) {
      sr[j] = ranges[j].first;
      sr[nband + j] = ranges[j].second;
    }
    ECCKD_TRY(upload_small(ctx, d_sfirst, sr.data(), sr.size() * sizeof(int)));
    ECCKD_TRY(upload_small(ctx, d_bits, bits_init, sizeof(bits_init)));
    gather_all_kernel<<<nb, 256, 0, ctx->stream>>>(nwav, d_key, nband, d_sfirst, d_slast, k0, v0, d_seg, d_bits,
                                                   d_bits + 1);
    ECCKD_LAUNCH_CHECK(ctx);
    unsigned long long bits[2];
    ECCKD_CUDA_OK(cudaMemcpyAsync(bits, d_bits, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    const unsigned long long vary = bits[0] & ~bits[1];
    uint32_t* v_sorted = nullptr;
    // one band spanning everything needs no segment pass
    const bool one_segment = nband == 1 && ranges[0].first == 0 && ranges[0].second == (int)(nwav - 1);
    ECCKD_TRY(radix_sort_pairs(ctx, nwav, k0, v0, k1, v1, hist, vary, one_segment ? nullptr : d_seg, &v_sorted));
    order_and_rank_kernel<<<nb, 256, 0, ctx->stream>>>(nwav, v_sorted, g_index, d_rank);
    ECCKD_LAUNCH_CHECK(ctx);
    return ECCKD_OK;
  }
  // overlapping bands: one after the other on g_index, as the reference's loop does
  for (int j = 0; j < nband; ++j) {
    const int i1 = fl[j], i2 = fl[nband + j];
    const int64_t n = (int64_t)i2 - i1 + 1;
    if (n < 2) continue;
    ECCKD_TRY(upload_small(ctx, d_bits, bits_init, sizeof(bits_init)));
    gather_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, g_index + i1, d_key, k0, v0, d_bits,
                                                                              d_bits + 1);
    ECCKD_LAUNCH_CHECK(ctx);
    unsigned long long bits[2];
    ECCKD_CUDA_OK(cudaMemcpyAsync(bits, d_bits, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    const unsigned long long vary = bits[0] & ~bits[1];
    uint32_t* v_sorted = nullptr;
    ECCKD_TRY(radix_sort_pairs(ctx, n, k0, v0, k1, v1, hist, vary, nullptr, &v_sorted));
    write_back_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, v_sorted, g_index + i1);
    ECCKD_LAUNCH_CHECK(ctx);
  }
  if (d_rank) {
    rank_from_order_kernel<<<nb, 256, 0, ctx->stream>>>(nwav, g_index, d_rank);
    ECCKD_LAUNCH_CHECK(ctx);
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_store_g_points_dev(ecckd_ctx* ctx, int64_t nwav, const int32_t* d_rank, int ng,
                                       const int32_t* h_rank1, const int32_t* h_rank2, int32_t* d_g_point) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_rank && h_rank1 && h_rank2 && d_g_point, ECCKD_PARAMETER_ERROR,
                "ecckd_store_g_points: NULL argument");
  ECCKD_REQUIRE(nwav > 0 && ng > 0 && ng <= 4096, ECCKD_PARAMETER_ERROR, "ecckd_store_g_points: bad dimensions");
  std::vector<int32_t> r12(2 * (size_t)ng);
  memcpy(r12.data(), h_rank1, (size_t)ng * sizeof(int32_t));
  memcpy(r12.data() + ng, h_rank2, (size_t)ng * sizeof(int32_t));
  void* d_r = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, r12.size() * sizeof(int32_t), &d_r));
  ECCKD_TRY(upload_small(ctx, d_r, r12.data(), r12.size() * sizeof(int32_t)));
  store_g_points_kernel<<<(unsigned)((nwav + 255) / 256), 256, r12.size() * sizeof(int32_t), ctx->stream>>>(
      nwav, d_rank, ng, static_cast<const int32_t*>(d_r), d_g_point);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// Hypercube partition of Hogan (2010): single_gas_data.cpp:23-124.
ECCKD_API int ecckd_overlap_g_points(int ngas, int nband, const int32_t* n_g, const double* sort_var,
                                     const int32_t* gas_offset, int cap, int32_t* g_min, int32_t* g_max,
                                     int32_t* band_number, int* ng_out) {
  ECCKD_REQUIRE(n_g && sort_var && gas_offset && g_min && g_max && band_number && ng_out, ECCKD_PARAMETER_ERROR,
                "ecckd_overlap_g_points: NULL argument");
  ECCKD_REQUIRE(ngas > 0 && nband > 0, ECCKD_PARAMETER_ERROR, "ecckd_overlap_g_points: bad dimensions");
  // merged g points per band: 1 - ngas + sum_gas n_g (Eq. 7 of Hogan 2010)
  int total = 0;
  std::vector<int> per_band(nband);
  for (int b = 0; b < nband; ++b) {
    int s = 1 - ngas;
    for (int a = 0; a < ngas; ++a) s += n_g[a * nband + b];
    per_band[b] = s;
    total += s;
  }
  ECCKD_REQUIRE(total <= cap, ECCKD_PARAMETER_ERROR, "ecckd_overlap_g_points: capacity %d < %d merged g points", cap,
                total);
  std::vector<int> cur(ngas, 0);  // current g point of each gas
  int m = 0;                      // current merged g point
  for (int b = 0; b < nband; ++b) {
    const std::vector<int> start = cur;
    // first merged point of the band: every gas at its weakest interval
    for (int a = 0; a < ngas; ++a) g_min[a * cap + m] = g_max[a * cap + m] = start[a];
    band_number[m] = b;
    for (int step = 1; step < per_band[b]; ++step) {
      // advance the gas whose next interval has the smallest sorting variable
      int pick = -1;
      double best = 1.0e30;
      for (int a = 0; a < ngas; ++a) {
        const bool can_advance = cur[a] < start[a] + n_g[a * nband + b] - 1;
        const double v = can_advance ? sort_var[gas_offset[a] + cur[a] + 1] : 1.0e30;
        if (v < best) {
          best = v;
          pick = a;
        }
      }
      if (pick < 0) {
        set_error("Could not locate next gas to advance");  // single_gas_data.cpp:92-95
        return ECCKD_PROCESSING_ERROR;
      }
      ++cur[pick];
      ++m;
      band_number[m] = b;
      for (int a = 0; a < ngas; ++a) {
        g_min[a * cap + m] = (a == pick) ? cur[a] : start[a];
        g_max[a * cap + m] = cur[a];
      }
    }
    ++m;
    for (int a = 0; a < ngas; ++a) ++cur[a];
  }
  *ng_out = total;
  return ECCKD_OK;
}

ECCKD_API int ecckd_assign_g_points_dev(ecckd_ctx* ctx, int ngas, int64_t nwav, const int32_t* d_gas_g_point, int ng,
                                        int cap, const int32_t* h_g_min, const int32_t* h_g_max, int32_t* d_g_point,
                                        int64_t* h_n_unassigned) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_gas_g_point && h_g_min && h_g_max && d_g_point, ECCKD_PARAMETER_ERROR,
                "ecckd_assign_g_points: NULL argument");
  ECCKD_REQUIRE(ngas > 0 && nwav > 0 && ng > 0 && cap >= ng, ECCKD_PARAMETER_ERROR, "ecckd_assign_g_points: bad dimensions");
  const size_t tab = (size_t)ngas * cap;
  const size_t smem = (size_t)2 * ngas * ng * sizeof(int32_t);
  ECCKD_REQUIRE(smem <= 48 * 1024, ECCKD_PARAMETER_ERROR, "ecckd_assign_g_points: ngas*ng too large");
  std::vector<int32_t> gmm(2 * tab);
  memcpy(gmm.data(), h_g_min, tab * sizeof(int32_t));
  memcpy(gmm.data() + tab, h_g_max, tab * sizeof(int32_t));
  void* d_t = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, gmm.size() * sizeof(int32_t) + 64, &d_t));
  unsigned long long* d_cnt = static_cast<unsigned long long*>(d_t);
  int32_t* d_gmm = reinterpret_cast<int32_t*>(d_cnt + 1);
  ECCKD_CUDA_OK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), ctx->stream));
  ECCKD_TRY(upload_small(ctx, d_gmm, gmm.data(), gmm.size() * sizeof(int32_t)));
  assign_g_points_kernel<<<(unsigned)((nwav + 255) / 256), 256, smem, ctx->stream>>>(ngas, nwav, d_gas_g_point, ng, cap,
                                                                                     d_gmm, d_g_point, d_cnt);
  ECCKD_LAUNCH_CHECK(ctx);
  if (h_n_unassigned) {
    unsigned long long c = 0;
    ECCKD_CUDA_OK(cudaMemcpyAsync(&c, d_cnt, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *h_n_unassigned = (int64_t)c;
  }
  return ECCKD_OK;
}
