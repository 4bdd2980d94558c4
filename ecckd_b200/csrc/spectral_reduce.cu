// spectral_reduce.cu - stage 1 of the hot path: reduce line-by-line spectra to
// per-g-point averages (average_optical_depth.cpp:22-197 of the reference),
// plus the small stage-1 companions (Planck function, merge, gpoint_fraction,
// Planck LUT).
//
// Kernel design (B200, sm_100a) -- see DESIGN.md section 3:
//   * A "row" is one (spectrum, layer) vector of nwav FP64 optical depths.  One
//     warp owns 32 rows (row = lane) and sweeps a span of wavenumber chunks.
//   * Chunks of CHUNK wavenumbers are staged into shared memory with
//     asynchronous copies (coalesced 512 B per row per request, double
//     buffered), so every HBM byte is read once, with full sectors.
//   * The g-point map is identical for every row and every call of a LUT build,
//     so a "plan" holds, per chunk, the local indices in g-sorted order.  The
//     warp visits a chunk in that order: the g index is warp-uniform, each lane
//     accumulates its row's running sums in registers, and only at a g change
//     (warp-uniform branch) are they folded into lane-private shared-memory
//     accumulators [g][4][lane] -- no atomics, no bank conflicts, deterministic
//     summation order.
//   * Spans of the wavenumber axis go to different warps/CTAs; their partial
//     accumulators are combined in fixed span order by the epilogue kernel,
//     which also applies the inverse transform, the min/max clamps and the
//     conversion to molar absorption.
#include "spectral_reduce.cuh"

#include <algorithm>

namespace ecckd {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// planck_function.cpp:29-33,47-53, one element
__device__ __forceinline__ double planck_elem(double t, double nu, double dnu) {
  const double h = 6.62606896e-34;
  const double c = 2.99792458e8;
  const double k = 1.3806504e-23;
  const double inv_cm_2_Hz = 100.0 * c;
  const double pi = 3.14159265358979323846;
  const double freq = nu * inv_cm_2_Hz;
  const double prefactor = (dnu * 2.0 * h * inv_cm_2_Hz * pi / (c * c)) * (freq * freq * freq);
  return prefactor / (exp((h / k) * (freq / t)) - 1.0);
}

// ------------------------------------------------------------------ plan
// One block of CHUNK threads per chunk: stable counting sort of the chunk's
// local indices by g point; slot = (g << 6) | local index, SLOT_END padding.
__global__ void __launch_bounds__(CHUNK) plan_kernel(const int32_t* __restrict__ g_point, int64_t nwav, int ng,
                                                      uint16_t* __restrict__ slots,
                                                      unsigned long long* __restrict__ count) {
  __shared__ int sg[CHUNK];
  const int64_t c = blockIdx.x;
  const int i = threadIdx.x;
  const int64_t w = c * CHUNK + i;
  int g = -1;
  if (w < nwav) {
    g = g_point[w];
    if (g < 0 || g >= ng) g = -1;
  }
  sg[i] = g;
  slots[c * CHUNK + i] = (uint16_t)SLOT_END;
  __syncthreads();
  if (g >= 0) {
    int pos = 0, same_before = 0, same = 0;
#pragma unroll 8
    for (int j = 0; j < CHUNK; ++j) {
      const int gj = sg[j];
      if (gj >= 0) {
        pos += (gj < g) || (gj == g && j < i);
        same += (gj == g);
        same_before += (gj == g && j < i);
      }
    }
    slots[c * CHUNK + pos] = (uint16_t)((g << 6) | i);
    if (same_before == 0) atomicAdd(&count[g], (unsigned long long)same);
  }
}

// ------------------------------------------------------------------ reduce
template <int FKIND>
__device__ __forceinline__ double f_transform(double tau, double kd, bool is_log, bool& valid) {
  valid = true;
  if (FKIND == F_LINEAR) return tau;
  if (FKIND == F_TRANS) return 1.0 - exp(-tau * kd);
  if (FKIND == F_SQRT) return sqrt(tau);
  if (FKIND == F_LOG) {
    valid = tau > 0.0;
    return valid ? log(tau) : 0.0;
  }
  if (FKIND == F_HYBRID) {
    if (is_log) {
      valid = tau > 0.0;
      return valid ? log(tau) : 0.0;
    }
    return 1.0 - exp(-tau * kd);
  }
  return 1.0;  // F_NONE: only the weight sum is wanted
}

struct WarpStage {
  double* data;    // [32][PITCH]
  double* wt;      // TILE: [maxw][PITCH]; VECTOR: [CHUNK]
  double* nu;      // PLANCK only: [CHUNK]
  double* dnu;     // PLANCK only: [CHUNK]
  uint16_t* plan;  // [CHUNK]
};

template <int VEC>
__device__ __forceinline__ void stage_row(double* dst, const double* src, int64_t w0, int64_t nwav, int lane) {
  // dst: CHUNK doubles in shared memory; src: row base in global memory
  if (VEC == 16) {
    const int col = lane * 2;
    if (w0 + col + 1 < nwav) cp_async16(dst + col, src + w0 + col);
    else if (w0 + col < nwav) cp_async8(dst + col, src + w0 + col);
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = lane + 32 * h;
      if (w0 + col < nwav) cp_async8(dst + col, src + w0 + col);
    }
  }
}

template <int FKIND, int WMODE, int VEC>
__global__ void __launch_bounds__(128) reduce_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * (blockDim.x >> 5) + warp;
  if (task >= p.ntasks) return;  // no block-wide barriers below
  const int group = task / p.nspans;
  const int span = task - group * p.nspans;
  const int64_t c0 = (p.nchunks * span) / p.nspans;
  const int64_t c1 = (p.nchunks * (span + 1)) / p.nspans;
  const int ng = p.ng;

  unsigned char* my = smem_raw + (size_t)warp * p.smem_per_warp;
  double* acc = reinterpret_cast<double*>(my);  // [ng][4][32]
  size_t off = (size_t)ng * 4 * 32 * sizeof(double);
  WarpStage st[NSTAGE];
  const int wt_doubles = (WMODE == W_VECTOR) ? CHUNK : p.maxw * PITCH;
#pragma unroll
  for (int s = 0; s < NSTAGE; ++s) {
    st[s].data = reinterpret_cast<double*>(my + off);
    off += (size_t)32 * PITCH * sizeof(double);
    st[s].wt = reinterpret_cast<double*>(my + off);
    off += (size_t)wt_doubles * sizeof(double);
    st[s].nu = reinterpret_cast<double*>(my + off);
    st[s].dnu = st[s].nu + CHUNK;
    if (WMODE == W_TILE_PLANCK) off += (size_t)2 * CHUNK * sizeof(double);
    st[s].plan = reinterpret_cast<uint16_t*>(my + off);
    off += CHUNK * sizeof(uint16_t);
  }

  const GroupDesc* __restrict__ gd = p.groups + group;
  const double* my_od = gd->od[lane];
  const bool active = my_od != nullptr;
  const int my_wslot = gd->wslot[lane];
  const bool my_is_log = (FKIND == F_HYBRID) ? (gd->is_log[lane] != 0) : (FKIND == F_LOG);
  const int n_w = gd->n_wrows;
  constexpr bool LOGKIND = (FKIND == F_LOG || FKIND == F_HYBRID);

  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  for (int g = 0; g < ng; ++g) {
    double* a = acc + (size_t)g * 128 + lane;
    a[0] = 0.0;
    a[32] = 0.0;
    a[64] = inf;
    a[96] = -inf;
  }

  auto issue = [&](int s, int64_t c) {
    const int64_t w0 = c * CHUNK;
    for (int r = 0; r < 32; ++r) {
      const double* src = gd->od[r];  // warp-uniform
      if (src) stage_row<VEC>(st[s].data + r * PITCH, src, w0, p.nwav, lane);
    }
    if (WMODE == W_TILE_MATRIX) {
      for (int t = 0; t < n_w; ++t) {
        stage_row<VEC>(st[s].wt + t * PITCH, p.weight + (int64_t)gd->wrows[t] * p.w_stride, w0, p.nwav, lane);
      }
    } else if (WMODE == W_VECTOR) {
      stage_row<VEC>(st[s].wt, p.weight, w0, p.nwav, lane);
    } else {
      stage_row<VEC>(st[s].nu, p.wn, w0, p.nwav, lane);
      stage_row<VEC>(st[s].dnu, p.dwn, w0, p.nwav, lane);
    }
    if (lane < 8) cp_async16(st[s].plan + lane * 8, p.plan + c * CHUNK + lane * 8);
  };

  // running sums of the current g run (registers)
  int gcur = -1;
  double s0 = 0.0, s1 = 0.0, mn = inf, mx = -inf;
  int zc = 0;
  const int slot_row = group * 32 + lane;

  auto flush = [&]() {
    if (gcur >= 0) {
      double* a = acc + (size_t)gcur * 128 + lane;
      a[0] += s0;
      a[32] += s1;
      a[64] = fmin(a[64], mn);
      a[96] = fmax(a[96], mx);
      if (LOGKIND && zc != 0 && active) atomicAdd(&p.zcount[(size_t)slot_row * ng + gcur], zc);
    }
    s0 = 0.0;
    s1 = 0.0;
    mn = inf;
    mx = -inf;
    zc = 0;
  };

  if (c0 < c1) issue(0, c0);
  cp_async_commit();
  for (int64_t c = c0; c < c1; ++c) {
    const int s = (int)((c - c0) & 1);
    if (c + 1 < c1) issue(s ^ 1, c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();

    const int64_t w0 = c * CHUNK;
    if (WMODE == W_TILE_PLANCK) {
      // weights of this chunk for the group's distinct temperatures
      for (int e = lane; e < n_w * CHUNK; e += 32) {
        const int t = e >> 6, i = e & (CHUNK - 1);
        if (w0 + i < p.nwav) {
          st[s].wt[t * PITCH + i] = planck_elem(p.temperature[gd->wrows[t]], st[s].nu[i], st[s].dnu[i]);
        }
      }
      __syncwarp();
    }

    const double* myrow = st[s].data + (active ? lane : 0) * PITCH;
    const double* myw = (WMODE == W_VECTOR) ? st[s].wt : (st[s].wt + my_wslot * PITCH);
    const uint4* plan4 = reinterpret_cast<const uint4*>(st[s].plan);
    bool done = false;
#pragma unroll 1
    for (int j0 = 0; j0 < CHUNK && !done; j0 += 8) {
      const uint4 pk = plan4[j0 >> 3];
      unsigned sl[8];
      sl[0] = pk.x & 0xFFFFu; sl[1] = pk.x >> 16;
      sl[2] = pk.y & 0xFFFFu; sl[3] = pk.y >> 16;
      sl[4] = pk.z & 0xFFFFu; sl[5] = pk.z >> 16;
      sl[6] = pk.w & 0xFFFFu; sl[7] = pk.w >> 16;
      double tau[8], ww[8], f[8];
      bool valid[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = sl[u] & (CHUNK - 1);
        tau[u] = myrow[idx];
        ww[u] = myw[idx];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = f_transform<FKIND>(tau[u], p.kd, my_is_log, valid[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (sl[u] == SLOT_END) {  // warp-uniform: only trailing slots are invalid
          done = true;
          break;
        }
        const int g = (int)(sl[u] >> 6);
        if (g != gcur) {  // warp-uniform
          flush();
          gcur = g;
        }
        if (LOGKIND) {
          const double we = valid[u] ? ww[u] : 0.0;
          s0 += f[u] * we;
          s1 += we;
          zc += valid[u] ? 0 : 1;
        } else {
          s0 += f[u] * ww[u];
          s1 += ww[u];
        }
        mn = fmin(mn, tau[u]);
        mx = fmax(mx, tau[u]);
      }
    }
    __syncwarp();  // everyone has finished reading stage s before it is refilled
  }
  flush();
  __syncwarp();
  double* out = p.partial + (size_t)task * ng * 128;
  for (int i = lane; i < ng * 128; i += 32) out[i] = acc[i];
}

// ------------------------------------------------------------------ combine + finalize
// Combine the span partials of every (row, g) in fixed span order (deterministic).
__global__ void combine_kernel(const CombineParams p) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (int64_t)p.nslots * p.ng) return;
  const int slot = (int)(tid / p.ng);
  const int g = (int)(tid - (int64_t)slot * p.ng);
  const RowDesc rd = p.rows[slot];
  if (rd.od == nullptr) return;
  const int group = slot >> 5, lane = slot & 31;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double s0 = 0.0, s1 = 0.0, mn = inf, mx = -inf;
  for (int s = 0; s < p.nspans; ++s) {
    const double* a = p.partial + ((size_t)(group * p.nspans + s) * p.ng + g) * 128 + lane;
    s0 += a[0];
    s1 += a[32];
    mn = fmin(mn, a[64]);
    mx = fmax(mx, a[96]);
  }
  double* so = p.sums + ((size_t)rd.row * p.ng + g) * NSUM;
  so[0] = s0;
  so[1] = s1;
  so[2] = (rd.method == M_LOG) ? (double)p.zcount[(size_t)slot * p.ng + g] : 0.0;
  double* mo = p.minmax + ((size_t)rd.row * p.ng + g) * NMINMAX;
  mo[0] = mn;
  mo[1] = -mx;
}

// average_optical_depth.cpp:43-193 after the sums: inverse transform, empty-g
// handling, min/max clamp and fixes, conversion to molar absorption.
__global__ void finalize_kernel(const FinalizeParams p) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (int64_t)p.nrows * p.ng) return;
  const int row = (int)(tid / p.ng);
  const int g = (int)(tid - (int64_t)row * p.ng);
  const OutDesc od = p.outs[row];
  const double* so = p.sums + ((size_t)row * p.ng + g) * NSUM;
  const double* mo = p.minmax + ((size_t)row * p.ng + g) * NMINMAX;
  const double s0 = so[0], s1 = so[1];
  double mn = mo[0], mx = -mo[1];
  if (od.method == M_SUMW) {  // planck look-up table: sum of weights only
    od.out[g] = s1;
    return;
  }
  const long long n = p.count[g];
  double fit;
  if (od.method == M_LINEAR) {
    fit = s0 / s1;
  } else if (od.method == M_TRANS) {
    const double x = fmin(0.9999999999999999, s0 / s1);
    fit = fabs(-log(1.0 - x) / p.kd);
  } else if (od.method == M_SQRT) {
    const double r = s0 / s1;
    fit = r * r;
  } else {  // M_LOG (average_optical_depth.cpp:79-96)
    const long long nnz = n - (long long)so[2];
    if (nnz == n) fit = exp(s0 / s1);
    else if (nnz == 0) fit = 0.0;
    else fit = exp(s0 / s1) * ((double)nnz / (double)n);
  }
  int wflag = 0;
  if (n == 0) {
    wflag |= ECCKD_WARN_EMPTY_G;
    fit = 0.0;
    mn = 0.0;
    mx = 0.0;
  } else {
    fit = fmax(mn, fmin(fit, mx));
    if (mn > fit) {
      wflag |= ECCKD_WARN_MIN_GT_FIT;
      mn = fit;
    }
    if (mn > 0.0 && mn >= mx) {
      wflag |= ECCKD_WARN_MIN_GE_MAX;
      mn *= 0.99;
      mx *= 1.01;
    }
  }
  if (od.factor > 0.0) {
    od.out[g] = od.factor * fit / od.dp;
    if (od.out_min) od.out_min[g] = od.factor * mn / od.dp;
    if (od.out_max) od.out_max[g] = od.factor * mx / od.dp;
  } else {
    od.out[g] = fit;
    if (od.out_min) od.out_min[g] = mn;
    if (od.out_max) od.out_max[g] = mx;
  }
  if (od.warn && wflag) atomicOr(&od.warn[g], wflag);
}

// ------------------------------------------------------------------ small kernels
__global__ void planck_kernel(int nt, const double* __restrict__ temperature, int64_t nwav,
                              const double* __restrict__ wn, const double* __restrict__ dwn,
                              double* __restrict__ ans, int64_t stride) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const double nu = wn[w], dnu = dwn[w];
  for (int t = blockIdx.y; t < nt; t += gridDim.y) ans[(int64_t)t * stride + w] = planck_elem(temperature[t], nu, dnu);
}

__global__ void merge_axpy_kernel(int nlay, int64_t nwav, double* __restrict__ dst, int64_t stride,
                                  const double* __restrict__ src, int64_t src_stride, double scaling,
                                  const double* __restrict__ profile, int mode) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  for (int l = blockIdx.y; l < nlay; l += gridDim.y) {
    const double v = src[(int64_t)l * src_stride + w];
    double* d = dst + (int64_t)l * stride + w;
    // read_merged_spectrum.cpp:156-167: three distinct expressions
    if (mode == 2) *d += v * profile[l];
    else if (mode == 1) *d += v * scaling;
    else *d += v;
  }
}

// gpoint_fraction numerators and per-g totals (create_look_up_table.cpp:537-548)
__global__ void gfrac_hist_kernel(int ng, int64_t nwav, const int32_t* __restrict__ g_point,
                                  const double* __restrict__ wn, const double* __restrict__ dwn, int nbin,
                                  double bin0, double dwav, double* __restrict__ sums,
                                  double* __restrict__ totals) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const int g = g_point[w];
  if (g < 0 || g >= ng) return;
  const double nu = wn[w], d = dwn[w];
  atomicAdd(&totals[g], d);
  // bin i covers (bin0 + i*dwav, bin0 + (i+1)*dwav]
  int i = (int)ceil((nu - bin0) / dwav) - 1;
  // repair rounding of the division against the exact edge tests of :544-545
  if (i >= 0 && i < nbin && !(nu > bin0 + dwav * i)) --i;
  if (i + 1 >= 0 && i + 1 < nbin && nu > bin0 + dwav * (i + 1)) ++i;
  if (i >= 0 && i < nbin && nu > bin0 + dwav * i && nu <= bin0 + dwav * (i + 1)) {
    atomicAdd(&sums[(size_t)g * nbin + i], d);
  }
}

__global__ void gfrac_norm_kernel(int ng, int nbin, const double* __restrict__ sums,
                                  const double* __restrict__ totals, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng * nbin) return;
  out[i] = sums[i] / totals[i / nbin];
}

__global__ void ssi_hist_kernel(int64_t nwav, const double* __restrict__ wn, const double* __restrict__ ssi,
                                int nbin, double bin0, double dwav, double* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const double nu = wn[w];
  int i = (int)ceil((nu - bin0) / dwav) - 1;
  if (i >= 0 && i < nbin && !(nu > bin0 + dwav * i)) --i;
  if (i + 1 >= 0 && i + 1 < nbin && nu > bin0 + dwav * (i + 1)) ++i;
  if (i >= 0 && i < nbin && nu > bin0 + dwav * i && nu <= bin0 + dwav * (i + 1)) atomicAdd(&out[i], ssi[w]);
}

__global__ void compact_used_kernel(int ng, int64_t nwav, const int32_t* __restrict__ g_point,
                                    int32_t* __restrict__ used) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const int g = g_point[w];
  if (g >= 0 && g < ng && used[g] == 0) used[g] = 1;  // benign race: all writers store 1
}

__global__ void compact_remap_kernel(int ng, int64_t nwav, int32_t* __restrict__ g_point,
                                     const int32_t* __restrict__ new_of_old, int32_t* __restrict__ bad) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  const int g = g_point[w];
  const int v = (g >= 0 && g < ng) ? new_of_old[g] : -1;
  if (v < 0) *bad = 1;
  g_point[w] = v;
}

// ------------------------------------------------------------------ host side
template <int FKIND, int WMODE>
static void launch_reduce_vec(int vec, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const ReduceParams& p,
                              cudaError_t* attr_err) {
  if (vec == 16) {
    *attr_err = cudaFuncSetAttribute(reduce_kernel<FKIND, WMODE, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
    if (*attr_err == cudaSuccess) reduce_kernel<FKIND, WMODE, 16><<<grid, block, smem, st>>>(p);
  } else {
    *attr_err = cudaFuncSetAttribute(reduce_kernel<FKIND, WMODE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
    if (*attr_err == cudaSuccess) reduce_kernel<FKIND, WMODE, 8><<<grid, block, smem, st>>>(p);
  }
}

template <int FKIND>
static void launch_reduce_w(int wmode, int vec, dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            const ReduceParams& p, cudaError_t* e) {
  if (wmode == W_TILE_MATRIX) launch_reduce_vec<FKIND, W_TILE_MATRIX>(vec, grid, block, smem, st, p, e);
  else if (wmode == W_VECTOR) launch_reduce_vec<FKIND, W_VECTOR>(vec, grid, block, smem, st, p, e);
  else launch_reduce_vec<FKIND, W_TILE_PLANCK>(vec, grid, block, smem, st, p, e);
}

static void launch_reduce(int fkind, int wmode, int vec, dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          const ReduceParams& p, cudaError_t* e) {
  switch (fkind) {
    case F_LINEAR: launch_reduce_w<F_LINEAR>(wmode, vec, grid, block, smem, st, p, e); break;
    case F_TRANS: launch_reduce_w<F_TRANS>(wmode, vec, grid, block, smem, st, p, e); break;
    case F_SQRT: launch_reduce_w<F_SQRT>(wmode, vec, grid, block, smem, st, p, e); break;
    case F_LOG: launch_reduce_w<F_LOG>(wmode, vec, grid, block, smem, st, p, e); break;
    case F_HYBRID: launch_reduce_w<F_HYBRID>(wmode, vec, grid, block, smem, st, p, e); break;
    default: launch_reduce_w<F_NONE>(wmode, vec, grid, block, smem, st, p, e); break;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Streaming stage: reduce `rows` (any mix of spectra/layers sharing one plan and
// one weight definition) to per-(row, g) partial results in d_sums / d_minmax
// (layouts in spectral_reduce.cuh), rows[i].row giving each row's index there.
static int reduce_partial(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, const std::vector<RowDesc>& rows, int fkind,
                          double kd, int wmode, const double* d_weight, int64_t w_stride,
                          const double* h_temperature, int n_temperature, const double* d_wn, const double* d_dwn,
                          double* d_sums, double* d_minmax) {
  const int nrows = (int)rows.size();
  if (nrows == 0) return ECCKD_OK;
  const int ng = plan->ng;
  const int64_t nwav = plan->nwav;
  const int ngroups = (nrows + 31) / 32;
  const int nslots = ngroups * 32;

  // vector width: 16 B copies need 16 B aligned rows and an even nwav
  int vec = (nwav % 2 == 0) ? 16 : 8;
  for (const RowDesc& r : rows) {
    if (!aligned16(r.od)) vec = 8;
  }
  if (wmode == W_TILE_MATRIX) {
    if (!aligned16(d_weight) || (w_stride % 2) != 0) vec = 8;
  } else if (wmode == W_VECTOR) {
    if (!aligned16(d_weight)) vec = 8;
  } else {
    if (!aligned16(d_wn) || !aligned16(d_dwn)) vec = 8;
  }

  // group descriptors: 32 consecutive rows per warp
  std::vector<GroupDesc> groups(ngroups);
  std::vector<RowDesc> slots(nslots);
  int maxw = 1;
  for (int gi = 0; gi < ngroups; ++gi) {
    GroupDesc& gd = groups[gi];
    memset(&gd, 0, sizeof(gd));
    for (int l = 0; l < 32; ++l) {
      const int r = gi * 32 + l;
      RowDesc& sd = slots[r];
      if (r < nrows) {
        sd = rows[r];
        gd.od[l] = sd.od;
        gd.is_log[l] = (sd.method == M_LOG) ? 1 : 0;
        int ws = -1;
        for (int t = 0; t < gd.n_wrows; ++t) {
          if (gd.wrows[t] == sd.wrow) ws = t;
        }
        if (ws < 0) {
          ws = gd.n_wrows++;
          gd.wrows[ws] = sd.wrow;
        }
        gd.wslot[l] = (uint8_t)ws;
      } else {
        memset(&sd, 0, sizeof(sd));
      }
    }
    if (gd.n_wrows == 0) {
      gd.n_wrows = 1;
      gd.wrows[0] = 0;
    }
    maxw = std::max(maxw, gd.n_wrows);
  }

  // shared memory per warp
  const size_t acc_bytes = (size_t)ng * 4 * 32 * sizeof(double);
  const size_t wt_bytes = (wmode == W_VECTOR ? (size_t)CHUNK : (size_t)maxw * PITCH) * sizeof(double);
  const size_t stage_bytes = (size_t)32 * PITCH * sizeof(double) + wt_bytes +
                             (wmode == W_TILE_PLANCK ? (size_t)2 * CHUNK * sizeof(double) : 0) +
                             CHUNK * sizeof(uint16_t);
  size_t per_warp = acc_bytes + NSTAGE * stage_bytes;
  per_warp = (per_warp + 127) & ~size_t(127);
  const size_t smem_max = ctx->smem_optin ? ctx->smem_optin : (size_t)227 * 1024;
  ECCKD_REQUIRE(per_warp <= smem_max, ECCKD_PARAMETER_ERROR,
                "average_optical_depth: ng = %d needs %zu B of shared memory per warp (limit %zu)", ng, per_warp,
                smem_max);
  const int warps = (int)std::min<size_t>(4, smem_max / per_warp);
  const size_t smem = per_warp * warps;

  // spans of the wavenumber axis: one wave of warps fills the machine
  const int64_t total_warps = (int64_t)ctx->num_sms * warps;
  int nspans = (int)std::max<int64_t>(1, total_warps / ngroups);
  nspans = (int)std::min<int64_t>(nspans, std::max<int64_t>(1, plan->nchunks));
  const int ntasks = ngroups * nspans;
  const int nblocks = (ntasks + warps - 1) / warps;

  // device scratch: [groups | slots] in ws0, span partials in ws1, zcount in ws2, temperatures in ws3
  void* d_desc = nullptr;
  const size_t groups_bytes = ((size_t)ngroups * sizeof(GroupDesc) + 255) & ~size_t(255);
  const size_t slots_bytes = (size_t)nslots * sizeof(RowDesc);
  ECCKD_TRY(ws_reserve(ctx, 0, groups_bytes + slots_bytes, &d_desc));
  GroupDesc* d_groups = static_cast<GroupDesc*>(d_desc);
  RowDesc* d_slots = reinterpret_cast<RowDesc*>(static_cast<char*>(d_desc) + groups_bytes);
  ECCKD_TRY(upload_small(ctx, d_groups, groups.data(), (size_t)ngroups * sizeof(GroupDesc)));
  ECCKD_TRY(upload_small(ctx, d_slots, slots.data(), slots_bytes));

  void* d_partial = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, (size_t)ntasks * ng * 128 * sizeof(double), &d_partial));
  void* d_z = nullptr;
  const bool logkind = (fkind == F_LOG || fkind == F_HYBRID);
  const size_t z_bytes = (size_t)nslots * ng * sizeof(int32_t);
  ECCKD_TRY(ws_reserve(ctx, 2, z_bytes, &d_z));
  if (logkind) ECCKD_CUDA_OK(cudaMemsetAsync(d_z, 0, z_bytes, ctx->stream));
  void* d_temp = nullptr;
  if (wmode == W_TILE_PLANCK) {
    ECCKD_TRY(ws_reserve(ctx, 3, (size_t)n_temperature * sizeof(double), &d_temp));
    ECCKD_TRY(upload_small(ctx, d_temp, h_temperature, (size_t)n_temperature * sizeof(double)));
  }

  ReduceParams rp;
  rp.groups = d_groups;
  rp.ngroups = ngroups;
  rp.nspans = nspans;
  rp.ntasks = ntasks;
  rp.plan = plan->d_slots;
  rp.nwav = nwav;
  rp.nchunks = plan->nchunks;
  rp.ng = ng;
  rp.weight = d_weight;
  rp.w_stride = w_stride;
  rp.temperature = static_cast<const double*>(d_temp);
  rp.wn = d_wn;
  rp.dwn = d_dwn;
  rp.kd = kd;
  rp.partial = static_cast<double*>(d_partial);
  rp.zcount = static_cast<int32_t*>(d_z);
  rp.maxw = maxw;
  rp.smem_per_warp = (int)per_warp;

  cudaError_t attr_err = cudaSuccess;
  launch_reduce(fkind, wmode, vec, dim3(nblocks), dim3(32 * warps), smem, ctx->stream, rp, &attr_err);
  ECCKD_CUDA_OK(attr_err);
  ECCKD_LAUNCH_CHECK(ctx);

  CombineParams cp;
  cp.rows = d_slots;
  cp.nslots = nslots;
  cp.nspans = nspans;
  cp.ng = ng;
  cp.partial = static_cast<const double*>(d_partial);
  cp.zcount = static_cast<const int32_t*>(d_z);
  cp.sums = d_sums;
  cp.minmax = d_minmax;
  const int64_t nthreads = (int64_t)nslots * ng;
  combine_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, ctx->stream>>>(cp);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// Finalize stage: partial results -> outputs (average_optical_depth.cpp:43-193).
// d_count[ng]: wavenumbers per g point over the WHOLE spectrum (all shards).
static int finalize_rows(ecckd_ctx* ctx, int ng, const long long* d_count, const std::vector<OutDesc>& outs, double kd,
                         const double* d_sums, const double* d_minmax) {
  const int nrows = (int)outs.size();
  if (nrows == 0) return ECCKD_OK;
  void* d_outs = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 0, (size_t)nrows * sizeof(OutDesc), &d_outs));
  ECCKD_TRY(upload_small(ctx, d_outs, outs.data(), (size_t)nrows * sizeof(OutDesc)));
  FinalizeParams fp;
  fp.outs = static_cast<const OutDesc*>(d_outs);
  fp.nrows = nrows;
  fp.ng = ng;
  fp.sums = d_sums;
  fp.minmax = d_minmax;
  fp.count = d_count;
  fp.kd = kd;
  const int64_t nthreads = (int64_t)nrows * ng;
  finalize_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, ctx->stream>>>(fp);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// Resolve the public averaging method into (kernel f kind, kd).
static int resolve_method(int averaging_method, int* fkind, double* kd) {
  const double OD_SCALING = 1.0;  // average_optical_depth.cpp:35
  switch (averaging_method) {
    case ECCKD_AVG_LINEAR: *fkind = F_LINEAR; *kd = 1.0; break;
    case ECCKD_AVG_TRANSMISSION: *fkind = F_TRANS; *kd = ECCKD_LW_DIFFUSIVITY * OD_SCALING; break;
    case ECCKD_AVG_TRANSMISSION_2: *fkind = F_TRANS; *kd = ECCKD_LW_DIFFUSIVITY * OD_SCALING * 2.0; break;
    case ECCKD_AVG_TRANSMISSION_3: *fkind = F_TRANS; *kd = ECCKD_LW_DIFFUSIVITY * OD_SCALING * 3.0; break;
    case ECCKD_AVG_TRANSMISSION_10: *fkind = F_TRANS; *kd = ECCKD_LW_DIFFUSIVITY * OD_SCALING * 10.0; break;
    case ECCKD_AVG_SQUARE_ROOT: *fkind = F_SQRT; *kd = 1.0; break;
    case ECCKD_AVG_LOGARITHMIC: *fkind = F_LOG; *kd = 1.0; break;
    case ECCKD_AVG_HYBRID_LOG_TRANS3: *fkind = F_HYBRID; *kd = ECCKD_LW_DIFFUSIVITY * OD_SCALING * 3.0; break;
    default:
      set_error("averaging_method %d not understood", averaging_method);
      return ECCKD_PARAMETER_ERROR;  // average_optical_depth.cpp:130-133
  }
  return ECCKD_OK;
}

static int row_method(int fkind, double pressure_fl) {
  switch (fkind) {
    case F_LINEAR: return M_LINEAR;
    case F_TRANS: return M_TRANS;
    case F_SQRT: return M_SQRT;
    case F_LOG: return M_LOG;
    default: return (pressure_fl > 100.0e2) ? M_LOG : M_TRANS;  // average_optical_depth.cpp:101
  }
}

static int resolve_weights(const ecckd_gpoint_plan* plan, int weight_mode, const double* d_weight, int64_t w_stride,
                           const double* h_temperature, const double* d_wn, const double* d_dwn, int* wmode) {
  if (weight_mode == ECCKD_WEIGHT_MATRIX) {
    ECCKD_REQUIRE(d_weight != nullptr, ECCKD_PARAMETER_ERROR, "weight matrix is NULL");
    *wmode = (w_stride == 0) ? W_VECTOR : W_TILE_MATRIX;  // stride 0 == spread<0>(ssi, nlay)
    ECCKD_REQUIRE(w_stride == 0 || w_stride >= plan->nwav, ECCKD_PARAMETER_ERROR, "weight stride smaller than nwav");
  } else if (weight_mode == ECCKD_WEIGHT_VECTOR) {
    ECCKD_REQUIRE(d_weight != nullptr, ECCKD_PARAMETER_ERROR, "weight vector is NULL");
    *wmode = W_VECTOR;
  } else if (weight_mode == ECCKD_WEIGHT_PLANCK) {
    ECCKD_REQUIRE(h_temperature && d_wn && d_dwn, ECCKD_PARAMETER_ERROR,
                  "Planck weights need temperature, wavenumber and d_wavenumber");
    *wmode = W_TILE_PLANCK;
  } else {
    set_error("weight_mode %d not understood", weight_mode);
    return ECCKD_PARAMETER_ERROR;
  }
  return ECCKD_OK;
}

}  // namespace ecckd

using namespace ecckd;

// ------------------------------------------------------------------ C ABI
ECCKD_API int ecckd_gpoint_plan_create(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng,
                                       ecckd_gpoint_plan** out) {
  ECCKD_REQUIRE(ctx && d_g_point && out, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_create: NULL argument");
  ECCKD_REQUIRE(nwav > 0 && ng > 0 && ng <= MAX_NG, ECCKD_PARAMETER_ERROR,
                "ecckd_gpoint_plan_create: need nwav > 0 and 0 < ng <= %d (got nwav=%lld ng=%d)", MAX_NG,
                (long long)nwav, ng);
  ecckd_gpoint_plan* p = new ecckd_gpoint_plan();
  p->ctx = ctx;
  p->nwav = nwav;
  p->ng = ng;
  p->nchunks = (nwav + CHUNK - 1) / CHUNK;
  *out = nullptr;
  cudaError_t e = cudaMalloc(&p->d_slots, (size_t)p->nchunks * CHUNK * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_count, (size_t)ng * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_count, 0, (size_t)ng * sizeof(unsigned long long), ctx->stream);
  if (e != cudaSuccess) {
    set_error("ecckd_gpoint_plan_create: %s", cudaGetErrorString(e));
    cudaFree(p->d_slots);
    cudaFree(p->d_count);
    delete p;
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  plan_kernel<<<(unsigned)p->nchunks, CHUNK, 0, ctx->stream>>>(d_g_point, nwav, ng, p->d_slots, p->d_count);
  ctx->launches++;
  std::vector<unsigned long long> cnt(ng);
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(cnt.data(), p->d_count, (size_t)ng * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    set_error("ecckd_gpoint_plan_create: %s", cudaGetErrorString(e));
    cudaFree(p->d_slots);
    cudaFree(p->d_count);
    delete p;
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  p->h_count.assign(cnt.begin(), cnt.end());
  *out = p;
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_destroy(ecckd_gpoint_plan* plan) {
  if (!plan) return ECCKD_OK;
  if (plan->ctx) cudaStreamSynchronize(plan->ctx->stream);
  cudaFree(plan->d_slots);
  cudaFree(plan->d_count);
  cudaFree(plan->d_scratch);
  delete plan;
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_counts(ecckd_gpoint_plan* plan, int64_t* h_count) {
  ECCKD_REQUIRE(plan && h_count, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_counts: NULL argument");
  for (int g = 0; g < plan->ng; ++g) h_count[g] = plan->h_count[g];
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_counts_dev(ecckd_gpoint_plan* plan, int64_t** d_count) {
  ECCKD_REQUIRE(plan && d_count, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_counts_dev: NULL argument");
  *d_count = reinterpret_cast<int64_t*>(plan->d_count);
  return ECCKD_OK;
}

ECCKD_API int ecckd_planck_function_dev(ecckd_ctx* ctx, int nt, const double* h_temperature, int64_t nwav,
                                        const double* d_wn, const double* d_dwn, double* d_ans,
                                        int64_t ans_stride) {
  ECCKD_REQUIRE(ctx && h_temperature && d_wn && d_dwn && d_ans, ECCKD_PARAMETER_ERROR,
                "ecckd_planck_function: NULL argument");
  // planck_function.cpp:40-44: output must be nt x nwav
  ECCKD_REQUIRE(nt > 0 && nwav > 0 && ans_stride >= nwav, ECCKD_PARAMETER_ERROR,
                "planck_function output array must be %d x %lld (row stride %lld)", nt, (long long)nwav,
                (long long)ans_stride);
  void* d_t = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 3, (size_t)nt * sizeof(double), &d_t));
  ECCKD_TRY(upload_small(ctx, d_t, h_temperature, (size_t)nt * sizeof(double)));
  dim3 grid((unsigned)((nwav + 255) / 256), (unsigned)std::min(nt, 64));
  planck_kernel<<<grid, 256, 0, ctx->stream>>>(nt, static_cast<const double*>(d_t), nwav, d_wn, d_dwn, d_ans,
                                               ans_stride);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_average_optical_depth_partial_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nspec, int nlay,
                                                      const double* h_pressure_fl, const double* const* h_d_od,
                                                      int64_t od_stride, int weight_mode, const double* d_weight,
                                                      int64_t w_stride, const double* h_temperature,
                                                      const double* d_wn, const double* d_dwn, int averaging_method,
                                                      double* d_sums, double* d_minmax) {
  ECCKD_REQUIRE(ctx && plan && h_pressure_fl && h_d_od && d_sums && d_minmax, ECCKD_PARAMETER_ERROR,
                "ecckd_average_optical_depth_partial: NULL argument");
  ECCKD_REQUIRE(nspec > 0 && nlay > 0, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: nspec and nlay must be > 0");
  ECCKD_REQUIRE(od_stride >= plan->nwav, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: od_stride < nwav");
  int fkind, wmode;
  double kd;
  ECCKD_TRY(resolve_method(averaging_method, &fkind, &kd));
  ECCKD_TRY(resolve_weights(plan, weight_mode, d_weight, w_stride, h_temperature, d_wn, d_dwn, &wmode));
  // rows ordered layer-major so that the 32 rows of a warp share few weight rows
  std::vector<RowDesc> rows;
  rows.reserve((size_t)nspec * nlay);
  for (int l = 0; l < nlay; ++l) {
    for (int s = 0; s < nspec; ++s) {
      ECCKD_REQUIRE(h_d_od[s] != nullptr, ECCKD_PARAMETER_ERROR, "spectrum %d: NULL device pointer", s);
      RowDesc r;
      r.od = h_d_od[s] + (int64_t)l * od_stride;
      r.method = row_method(fkind, h_pressure_fl[l]);
      r.wrow = l;
      r.row = l * nspec + s;
      r.pad_ = 0;
      rows.push_back(r);
    }
  }
  return reduce_partial(ctx, plan, rows, fkind, kd, wmode, d_weight, w_stride, h_temperature, nlay, d_wn, d_dwn, d_sums,
                        d_minmax);
}

ECCKD_API int ecckd_average_optical_depth_finalize_dev(ecckd_ctx* ctx, int ng, const int64_t* d_count, int nspec,
                                                       const double* h_reference_surface_vmr, int nlay,
                                                       const double* h_pressure_fl, const double* h_pressure_hl,
                                                       int averaging_method, const double* d_sums,
                                                       const double* d_minmax, double* const* h_d_molar_abs,
                                                       double* const* h_d_min, double* const* h_d_max,
                                                       int64_t out_stride, int32_t* const* h_d_warn) {
  ECCKD_REQUIRE(ctx && d_count && h_reference_surface_vmr && h_pressure_fl && h_pressure_hl && d_sums && d_minmax &&
                    h_d_molar_abs,
                ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth_finalize: NULL argument");
  ECCKD_REQUIRE(nspec > 0 && nlay > 0 && ng > 0 && out_stride >= ng, ECCKD_PARAMETER_ERROR,
                "ecckd_average_optical_depth_finalize: bad dimensions");
  int fkind;
  double kd;
  ECCKD_TRY(resolve_method(averaging_method, &fkind, &kd));
  std::vector<OutDesc> outs((size_t)nspec * nlay);
  for (int l = 0; l < nlay; ++l) {
    for (int s = 0; s < nspec; ++s) {
      ECCKD_REQUIRE(h_d_molar_abs[s] != nullptr, ECCKD_PARAMETER_ERROR, "spectrum %d: NULL output pointer", s);
      OutDesc& o = outs[(size_t)l * nspec + s];
      o.out = h_d_molar_abs[s] + (int64_t)l * out_stride;
      const bool mm = h_d_min && h_d_max && h_d_min[s] && h_d_max[s];
      o.out_min = mm ? h_d_min[s] + (int64_t)l * out_stride : nullptr;
      o.out_max = mm ? h_d_max[s] + (int64_t)l * out_stride : nullptr;
      o.warn = (h_d_warn && h_d_warn[s]) ? h_d_warn[s] : nullptr;
      const double vmr = h_reference_surface_vmr[s];
      o.factor = (vmr > 0.0) ? ((ECCKD_ACCEL_GRAVITY * 0.001 * ECCKD_MOLAR_MASS_DRY_AIR) / vmr) : 0.0;
      o.dp = h_pressure_hl[l + 1] - h_pressure_hl[l];
      o.method = row_method(fkind, h_pressure_fl[l]);
      o.pad_ = 0;
    }
  }
  return finalize_rows(ctx, ng, reinterpret_cast<const long long*>(d_count), outs, kd, d_sums, d_minmax);
}

ECCKD_API int ecckd_average_optical_depth_batch_dev(
    ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nspec, const double* h_reference_surface_vmr, int nlay,
    const double* h_pressure_fl, const double* h_pressure_hl, const double* const* h_d_od, int64_t od_stride,
    int weight_mode, const double* d_weight, int64_t w_stride, const double* h_temperature, const double* d_wn,
    const double* d_dwn, int averaging_method, double* const* h_d_molar_abs, double* const* h_d_min,
    double* const* h_d_max, int64_t out_stride, int32_t* const* h_d_warn) {
  ECCKD_REQUIRE(ctx && plan, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: NULL argument");
  ECCKD_REQUIRE(nspec > 0 && nlay > 0, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: nspec and nlay must be > 0");
  // partial sums live in the plan-independent scratch slot 3 tail; sized [nrows][ng][NSUM + NMINMAX]
  const size_t nrg = (size_t)nspec * nlay * plan->ng;
  if (plan->scratch_bytes < nrg * (NSUM + NMINMAX) * sizeof(double)) {
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(plan->d_scratch);
    plan->d_scratch = nullptr;
    plan->scratch_bytes = 0;
    ECCKD_CUDA_OK(cudaMalloc(&plan->d_scratch, nrg * (NSUM + NMINMAX) * sizeof(double)));
    plan->scratch_bytes = nrg * (NSUM + NMINMAX) * sizeof(double);
  }
  double* d_sums = plan->d_scratch;
  double* d_minmax = d_sums + nrg * NSUM;
  ECCKD_TRY(ecckd_average_optical_depth_partial_dev(ctx, plan, nspec, nlay, h_pressure_fl, h_d_od, od_stride,
                                                    weight_mode, d_weight, w_stride, h_temperature, d_wn, d_dwn,
                                                    averaging_method, d_sums, d_minmax));
  return ecckd_average_optical_depth_finalize_dev(ctx, plan->ng, reinterpret_cast<const int64_t*>(plan->d_count), nspec,
                                                  h_reference_surface_vmr, nlay, h_pressure_fl, h_pressure_hl,
                                                  averaging_method, d_sums, d_minmax, h_d_molar_abs, h_d_min, h_d_max,
                                                  out_stride, h_d_warn);
}

ECCKD_API int ecckd_average_optical_depth_to_g_point_dev(
    ecckd_ctx* ctx, ecckd_gpoint_plan* plan, double reference_surface_vmr, int nlay, const double* h_pressure_fl,
    const double* h_pressure_hl, const double* d_od, int64_t od_stride, int weight_mode, const double* d_weight,
    int64_t w_stride, const double* h_temperature, const double* d_wn, const double* d_dwn, int averaging_method,
    double* d_molar_abs, double* d_min, double* d_max, int64_t out_stride, int32_t* d_warn) {
  const double* ods[1] = {d_od};
  double* outs[1] = {d_molar_abs};
  double* mins[1] = {d_min};
  double* maxs[1] = {d_max};
  int32_t* warns[1] = {d_warn};
  return ecckd_average_optical_depth_batch_dev(ctx, plan, 1, &reference_surface_vmr, nlay, h_pressure_fl,
                                               h_pressure_hl, ods, od_stride, weight_mode, d_weight, w_stride,
                                               h_temperature, d_wn, d_dwn, averaging_method, outs, mins, maxs,
                                               out_stride, warns);
}

ECCKD_API int ecckd_average_optical_depth_to_g_point(ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay,
                                                     int64_t nwav, const double* pressure_fl,
                                                     const double* pressure_hl, const int32_t* g_point,
                                                     const double* optical_depth, int64_t od_stride,
                                                     const double* planck_fl, int64_t planck_stride,
                                                     const char* averaging_method, double* molar_abs,
                                                     double* min_molar_abs, double* max_molar_abs, int64_t out_stride,
                                                     int32_t* warn) {
  ECCKD_REQUIRE(ctx && pressure_fl && pressure_hl && g_point && optical_depth && planck_fl && molar_abs,
                ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth_to_g_point: NULL argument");
  const int method = ecckd_averaging_method_from_string(averaging_method);
  if (method < 0) {
    set_error("averaging_method \"%s\" not understood", averaging_method ? averaging_method : "(null)");
    return ECCKD_PARAMETER_ERROR;
  }
  ECCKD_REQUIRE(nlay > 0 && nwav > 0 && ng > 0 && od_stride >= nwav && out_stride >= ng, ECCKD_PARAMETER_ERROR,
                "ecckd_average_optical_depth_to_g_point: bad dimensions");
  cudaStream_t st = ctx->stream;
  const bool vecw = (planck_stride == 0);
  // device staging (even row strides so 16 B copies are legal)
  const int64_t dstride = (nwav + 1) & ~int64_t(1);
  double *d_od = nullptr, *d_w = nullptr, *d_out = nullptr;
  int32_t *d_g = nullptr, *d_warn = nullptr;
  const size_t od_bytes = (size_t)nlay * dstride * sizeof(double);
  const size_t w_bytes = (vecw ? (size_t)dstride : (size_t)nlay * dstride) * sizeof(double);
  const size_t out_elems = (size_t)nlay * ng;
  ECCKD_CUDA_OK(cudaMalloc(&d_od, od_bytes + w_bytes + 3 * out_elems * sizeof(double)));
  d_w = d_od + (size_t)nlay * dstride;
  d_out = d_w + w_bytes / sizeof(double);
  int rc = ECCKD_OK;
  ecckd_gpoint_plan* plan = nullptr;
  do {
    if (cudaMalloc(&d_g, (size_t)(nwav + ng) * sizeof(int32_t)) != cudaSuccess) { rc = ECCKD_UNEXPECTED_EXCEPTION; set_error("cudaMalloc failed"); break; }
    d_warn = d_g + nwav;
    cudaError_t e = cudaMemcpyAsync(d_g, g_point, (size_t)nwav * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_warn, 0, (size_t)ng * sizeof(int32_t), st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(d_od, dstride * sizeof(double), optical_depth, od_stride * sizeof(double),
                            nwav * sizeof(double), nlay, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      if (vecw) e = cudaMemcpyAsync(d_w, planck_fl, nwav * sizeof(double), cudaMemcpyHostToDevice, st);
      else
        e = cudaMemcpy2DAsync(d_w, dstride * sizeof(double), planck_fl, planck_stride * sizeof(double),
                              nwav * sizeof(double), nlay, cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) { rc = ECCKD_UNEXPECTED_EXCEPTION; set_error("H2D copy failed: %s", cudaGetErrorString(e)); break; }
    rc = ecckd_gpoint_plan_create(ctx, d_g, nwav, ng, &plan);
    if (rc) break;
    const bool mm = min_molar_abs && max_molar_abs;
    rc = ecckd_average_optical_depth_to_g_point_dev(
        ctx, plan, reference_surface_vmr, nlay, pressure_fl, pressure_hl, d_od, dstride,
        vecw ? ECCKD_WEIGHT_VECTOR : ECCKD_WEIGHT_MATRIX, d_w, vecw ? 0 : dstride, nullptr, nullptr, nullptr, method,
        d_out, mm ? d_out + out_elems : nullptr, mm ? d_out + 2 * out_elems : nullptr, ng, d_warn);
    if (rc) break;
    e = cudaMemcpy2DAsync(molar_abs, out_stride * sizeof(double), d_out, ng * sizeof(double), ng * sizeof(double), nlay,
                          cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && mm) {
      e = cudaMemcpy2DAsync(min_molar_abs, out_stride * sizeof(double), d_out + out_elems, ng * sizeof(double),
                            ng * sizeof(double), nlay, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(max_molar_abs, out_stride * sizeof(double), d_out + 2 * out_elems, ng * sizeof(double),
                              ng * sizeof(double), nlay, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess && warn) e = cudaMemcpyAsync(warn, d_warn, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = ECCKD_UNEXPECTED_EXCEPTION; set_error("D2H copy failed: %s", cudaGetErrorString(e)); break; }
  } while (0);
  cudaStreamSynchronize(st);
  if (plan) ecckd_gpoint_plan_destroy(plan);
  cudaFree(d_g);
  cudaFree(d_od);
  return rc;
}

ECCKD_API int ecckd_merge_axpy_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, double* d_optical_depth, int64_t stride,
                                   const double* d_od, int64_t od_stride, double scaling,
                                   const double* h_scaling_profile) {
  ECCKD_REQUIRE(ctx && d_optical_depth && d_od, ECCKD_PARAMETER_ERROR, "ecckd_merge_axpy: NULL argument");
  ECCKD_REQUIRE(nlay > 0 && nwav > 0, ECCKD_PARAMETER_ERROR, "ecckd_merge_axpy: bad dimensions");
  void* d_prof = nullptr;
  int mode = (scaling != 1.0) ? 1 : 0;
  if (h_scaling_profile) {
    ECCKD_TRY(ws_reserve(ctx, 3, (size_t)nlay * sizeof(double), &d_prof));
    ECCKD_TRY(upload_small(ctx, d_prof, h_scaling_profile, (size_t)nlay * sizeof(double)));
    mode = 2;
  }
  dim3 grid((unsigned)((nwav + 255) / 256), (unsigned)std::min(nlay, 64));
  merge_axpy_kernel<<<grid, 256, 0, ctx->stream>>>(nlay, nwav, d_optical_depth, stride, d_od, od_stride, scaling,
                                                   static_cast<const double*>(d_prof), mode);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

static int check_uniform_bins(int nbin, const double* w1, const double* w2, double* bin0, double* dwav) {
  ECCKD_REQUIRE(nbin > 0 && w1 && w2, ECCKD_PARAMETER_ERROR, "gpoint_fraction: no bins");
  *bin0 = w1[0];
  *dwav = w2[0] - w1[0];
  ECCKD_REQUIRE(*dwav > 0.0, ECCKD_PARAMETER_ERROR, "gpoint_fraction: non-positive bin width");
  for (int i = 0; i < nbin; ++i) {
    ECCKD_REQUIRE(w1[i] == *bin0 + *dwav * i && w2[i] == *bin0 + *dwav * (i + 1), ECCKD_PARAMETER_ERROR,
                  "gpoint_fraction: bins must form the uniform contiguous grid of create_look_up_table.cpp:526-535");
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_fraction_dev(ecckd_ctx* ctx, int ng, int64_t nwav, const int32_t* d_g_point,
                                        const double* d_wn, const double* d_dwn, int nbin, const double* h_w1,
                                        const double* h_w2, double* d_out) {
  ECCKD_REQUIRE(ctx && d_g_point && d_wn && d_dwn && d_out, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_fraction: NULL argument");
  double bin0, dwav;
  ECCKD_TRY(check_uniform_bins(nbin, h_w1, h_w2, &bin0, &dwav));
  void* d_tmp = nullptr;
  const size_t n = (size_t)ng * nbin + ng;
  ECCKD_TRY(ws_reserve(ctx, 1, n * sizeof(double), &d_tmp));
  ECCKD_CUDA_OK(cudaMemsetAsync(d_tmp, 0, n * sizeof(double), ctx->stream));
  double* sums = static_cast<double*>(d_tmp);
  double* totals = sums + (size_t)ng * nbin;
  gfrac_hist_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(ng, nwav, d_g_point, d_wn, d_dwn, nbin,
                                                                             bin0, dwav, sums, totals);
  ECCKD_LAUNCH_CHECK(ctx);
  gfrac_norm_kernel<<<(ng * nbin + 255) / 256, 256, 0, ctx->stream>>>(ng, nbin, sums, totals, d_out);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_ssi_intervals_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wn, const double* d_ssi, int nbin,
                                      const double* h_w1, const double* h_w2, double* d_out) {
  ECCKD_REQUIRE(ctx && d_wn && d_ssi && d_out, ECCKD_PARAMETER_ERROR, "ecckd_ssi_intervals: NULL argument");
  double bin0, dwav;
  ECCKD_TRY(check_uniform_bins(nbin, h_w1, h_w2, &bin0, &dwav));
  ECCKD_CUDA_OK(cudaMemsetAsync(d_out, 0, (size_t)nbin * sizeof(double), ctx->stream));
  ssi_hist_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(nwav, d_wn, d_ssi, nbin, bin0, dwav, d_out);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_planck_lut_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, const double* d_wn, const double* d_dwn,
                                   int nlut, const double* h_temperature_lut, double* d_out) {
  ECCKD_REQUIRE(ctx && plan && d_wn && d_dwn && h_temperature_lut && d_out && nlut > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_planck_lut: bad argument");
  // sum over each g point of planck(T, nu): the reduction with f = 1 and Planck weights
  std::vector<RowDesc> rows(nlut);
  std::vector<OutDesc> outs(nlut);
  for (int t = 0; t < nlut; ++t) {
    RowDesc& r = rows[t];
    r.od = d_wn;  // any valid row: F_NONE ignores the values
    r.method = M_SUMW;
    r.wrow = t;
    r.row = t;
    r.pad_ = 0;
    OutDesc& o = outs[t];
    memset(&o, 0, sizeof(o));
    o.out = d_out + (int64_t)t * plan->ng;
    o.method = M_SUMW;
  }
  void* scratch = nullptr;
  const size_t nrg = (size_t)nlut * plan->ng;
  ECCKD_CUDA_OK(cudaMalloc(&scratch, nrg * (NSUM + NMINMAX) * sizeof(double)));
  double* d_sums = static_cast<double*>(scratch);
  double* d_minmax = d_sums + nrg * NSUM;
  int rc = reduce_partial(ctx, plan, rows, F_NONE, 1.0, W_TILE_PLANCK, nullptr, 0, h_temperature_lut, nlut, d_wn, d_dwn,
                          d_sums, d_minmax);
  if (rc == ECCKD_OK)
    rc = finalize_rows(ctx, plan->ng, reinterpret_cast<const long long*>(plan->d_count), outs, 1.0, d_sums, d_minmax);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(scratch);
  return rc;
}

ECCKD_API int ecckd_compact_g_points_dev(ecckd_ctx* ctx, int ng, int64_t nwav, int32_t* d_g_point,
                                         int32_t* h_g_point_map, int* new_ng) {
  ECCKD_REQUIRE(ctx && d_g_point && h_g_point_map && new_ng && ng > 0 && nwav > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_compact_g_points: bad argument");
  void* d_tmp = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 2, (size_t)(2 * ng + 1) * sizeof(int32_t), &d_tmp));
  int32_t* d_used = static_cast<int32_t*>(d_tmp);
  int32_t* d_map = d_used + ng;
  int32_t* d_bad = d_map + ng;
  ECCKD_CUDA_OK(cudaMemsetAsync(d_tmp, 0, (size_t)(2 * ng + 1) * sizeof(int32_t), ctx->stream));
  const unsigned nb = (unsigned)((nwav + 255) / 256);
  compact_used_kernel<<<nb, 256, 0, ctx->stream>>>(ng, nwav, d_g_point, d_used);
  ECCKD_LAUNCH_CHECK(ctx);
  std::vector<int32_t> used(ng), new_of_old(ng, -1);
  ECCKD_CUDA_OK(cudaMemcpyAsync(used.data(), d_used, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  int n = 0;
  for (int g = 0; g < ng; ++g) {
    if (used[g]) {
      h_g_point_map[n] = g;
      new_of_old[g] = n++;
    }
  }
  *new_ng = n;
  ECCKD_TRY(upload_small(ctx, d_map, new_of_old.data(), (size_t)ng * sizeof(int32_t)));
  compact_remap_kernel<<<nb, 256, 0, ctx->stream>>>(ng, nwav, d_g_point, d_map, d_bad);
  ECCKD_LAUNCH_CHECK(ctx);
  int32_t bad = 0;
  ECCKD_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (bad) {
    set_error("Some unassigned spectral points after mapping");  // create_look_up_table.cpp:146-149
    return 1;
  }
  return ECCKD_OK;
}

// Host form of planck_function (planck_function.cpp:22-54)
ECCKD_API int ecckd_planck_function(ecckd_ctx* ctx, int nt, const double* temperature, int64_t nwav,
                                    const double* wavenumber_cm_1, const double* d_wavenumber_cm_1, double* ans,
                                    int64_t ans_stride) {
  ECCKD_REQUIRE(ctx && temperature && wavenumber_cm_1 && d_wavenumber_cm_1 && ans, ECCKD_PARAMETER_ERROR,
                "ecckd_planck_function: NULL argument");
  ECCKD_REQUIRE(nt > 0 && nwav > 0 && ans_stride >= nwav, ECCKD_PARAMETER_ERROR, "ecckd_planck_function: bad dimensions");
  double* d = nullptr;
  ECCKD_CUDA_OK(cudaMalloc(&d, ((size_t)2 * nwav + (size_t)nt * nwav) * sizeof(double)));
  int rc = ECCKD_OK;
  cudaError_t e = cudaMemcpyAsync(d, wavenumber_cm_1, nwav * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + nwav, d_wavenumber_cm_1, nwav * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) rc = ecckd_planck_function_dev(ctx, nt, temperature, nwav, d, d + nwav, d + 2 * nwav, nwav);
  if (e == cudaSuccess && rc == ECCKD_OK)
    e = cudaMemcpy2DAsync(ans, ans_stride * sizeof(double), d + 2 * nwav, nwav * sizeof(double), nwav * sizeof(double), nt,
                          cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) {
    set_error("ecckd_planck_function: %s", cudaGetErrorString(e));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  return rc;
}
