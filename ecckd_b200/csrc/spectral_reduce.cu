// spectral_reduce.cu - stage 1 of the hot path: reduce line-by-line spectra to
// per-g-point averages (average_optical_depth.cpp:22-197 of the reference),
// plus the small stage-1 companions (Planck function, merge, gpoint_fraction,
// Planck LUT).
//
// Kernel design (B200, sm_100a) -- see DESIGN.md section 3:
//   * A "row" is one (spectrum, layer) vector of nwav FP64 optical depths.  Rows
//     that share a weight row (all spectra of one layer) form a group of up to
//     20; a CTA gives every consumer warp two rows of the group and walks chunks
//     of `ch` (<= 480) wavenumbers.  A consumer warp stages its rows' chunks in
//     shared memory with bulk asynchronous copies (cp.async.bulk + mbarrier,
//     two chunk visits deep, L2 evict-first): every optical depth is read from
//     HBM exactly once, 3.8 KB at a time, ~75 KB in flight per SM.
//   * The weights of a chunk (Planck function, or a streamed row) are produced
//     once per CTA by two producer warps into a 4-slot ring together with the
//     chunk's plan record.
//   * The g-point map is identical for every row and every call of a LUT build,
//     so a "plan" assigns, per chunk, every wavenumber to a (lane, slot) such
//     that each lane only sees one g point per round (32 lanes x 16 slots), the
//     lanes of a g point are adjacent and consecutive wavenumbers sit in
//     consecutive lanes.  A lane keeps its 16 shared-memory addresses and 16
//     weights in registers for both of its rows, evaluates the slots with no
//     branches (trans_core.cuh: 12 FP64 instructions per 1 - exp), then a
//     segmented shuffle reduction whose lane predicates come from the plan
//     folds the lanes of each g point and the head lane adds the result to the
//     warp's shared-memory accumulators [row][ng][4].  No atomics anywhere;
//     the summation order is fixed by the plan and the launch geometry.
//   * Work items are ordered (wavenumber window, group, chunk) and cut into
//     equal contiguous pieces, one per CTA, so that the per-wavenumber data of
//     a window (wavenumbers, plan) are re-read from L2, not HBM, by the next
//     group.  A CTA flushes its accumulators when it crosses into the next
//     (window, group); a small combine kernel adds the pieces in fixed order.
#include "spectral_reduce.cuh"
#include "reduce_rows.cuh"
#include "async_copy.cuh"

#include <algorithm>
#include <ctime>
#include <type_traits>

// (defined with the C ABI at the end of this file)
static int plan_create_impl(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng, bool free_layout,
                            ecckd_gpoint_plan** out);

namespace ecckd {

// exp(x) without branches: Cody-Waite reduction by ln 2 and a degree-11 polynomial; maximum
// error 0.90 ulp on [-690, 700] (checked against a long-double reference on the host, see
// DESIGN.md).  Used by the Planck weights of the producer warps.
__constant__ double EXP_C[12] = {2.5052097064908941e-8, 2.7626262793835868e-7, 2.7557414788000726e-6,
                                 2.4801504602132958e-5, 1.9841269707468915e-4, 1.3888888932258898e-3,
                                 8.3333333333978320e-3, 4.1666666666573905e-2, 1.6666666666666563e-1,
                                 5.0000000000000056e-1, 1.0,                   1.0};
__constant__ double EXP_K[4] = {1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01,
                                -1.90821492927058770002e-10};

__device__ __forceinline__ double fast_exp(double x) {
  const double t = fma(x, EXP_K[0], EXP_K[1]);
  int k = __double2loint(t);
  const double kf = t - EXP_K[1];
  double r = fma(kf, EXP_K[2], x);
  r = fma(kf, EXP_K[3], r);
  double q = EXP_C[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) q = fma(q, r, EXP_C[i]);
  k = max(-1000, min(k, 1000));
  return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
}

// planck_function.cpp:29-33,47-53, one element, for the producer warps (all 32 lanes call it
// together: the exponential looks its 2^(j/32) factor up by warp shuffle, trans_core.cuh).  The
// reference evaluates prefactor / (exp((h/k)*(freq/T)) - 1) with prefactor =
// (dnu*2*h*inv_cm_2_Hz*pi/(c*c)) * freq^3; here the constants are folded (hk_over_t = (h/k)/T is
// formed on the host), and the division is a reciprocal refined by two Newton steps: the weight
// differs from the reference's by a few ulp, far inside the 1e-12 tolerance of the averages.
__device__ __forceinline__ double planck_elem_fast(double hk_over_t, double nu, double dnu, const TransLane& tl) {
  const double h = 6.62606896e-34;
  const double c = 2.99792458e8;
  const double inv_cm_2_Hz = 100.0 * c;
  const double pi = 3.14159265358979323846;
  const double c1 = 2.0 * h * inv_cm_2_Hz * pi / (c * c);
  const double freq = nu * inv_cm_2_Hz;
  const double prefactor = (dnu * c1) * (freq * freq * freq);
  const double e = exp_table(freq * hk_over_t, tl) - 1.0;
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(e));  // ~20 bits
  x = fma(x, fma(-e, x, 1.0), x);
  x = fma(x, fma(-e, x, 1.0), x);
  return prefactor * x;
}

// The form the producers of reduce2_kernel use (they are FP64-bound next to two consumer warps on
// their scheduler): the per-wavenumber factor dnu c1 (nu c)^3 comes precomputed (planck_prefactor_kernel,
// same operations in the same order as above), the exponential from the consumers' 2^(j/2048) table
// in shared memory with the minimax polynomial of (e^r - 1)/r (relative error < 1e-17) and a two-constant
// Cody-Waite reduction, and the reciprocal from one cubically convergent step on the 23-bit
// approximation (error eps^3 < 2^-69): 13 FP64 instructions per weight instead of ~23.
// hkc_over_t = (h/k) (100 c) / T, formed on the host.
#if ECCKD_TAB_LOGN == 11 && ECCKD_TAB_LOGR == 0
#define ECCKD_PLANCK_TAB 1
__device__ __forceinline__ double planck_weight_tab(double hkc_over_t, double nu, double pref,
                                                    const unsigned char* __restrict__ tab) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  const double x = nu * hkc_over_t;
  const double t = fma(x, 0x1.71547652b82fep+11, magic);  // 2048/ln2
  const int n = __double2loint(t);
  const double nf = t - magic;
  double r = fma(nf, -0x1.62e42fee00000p-12, x);  // -ln2/2048, high part (21 trailing zero bits)
  r = fma(nf, -0x1.a39ef35793c76p-44, r);         // low part
  double q = fma(EXPM1_Q2_2048_D[2], r, EXPM1_Q2_2048_D[1]);
  q = fma(q, r, EXPM1_Q2_2048_D[0]);
  const double P = fma(q, r, 1.0);
  const uint2 e = *reinterpret_cast<const uint2*>(tab + (((unsigned)n << 3) & 0x3ff8u));
  const double T = __hiloint2double((int)e.y + (n << 9), (int)e.x);
  const double em1 = fma(T, P, -1.0);
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(em1));  // relative error eps <= 2^-23
  const double eps = fma(-em1, r0, 1.0);
  const double r1 = fma(r0, fma(eps, eps, eps), r0);  // r0 (1 + eps + eps^2) = (1 - eps^3) / em1
  return pref * r1;
}
#endif
__global__ void planck_prefactor_kernel(int64_t n, const double* __restrict__ wn, const double* __restrict__ dwn,
                                        double* __restrict__ pref) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double h = 6.62606896e-34;
  const double c = 2.99792458e8;
  const double inv_cm_2_Hz = 100.0 * c;
  const double pi = 3.14159265358979323846;
  const double c1 = 2.0 * h * inv_cm_2_Hz * pi / (c * c);
  const double freq = wn[i] * inv_cm_2_Hz;
  pref[i] = (dwn[i] * c1) * (freq * freq * freq);
}

// exact evaluation order of planck_function.cpp:47-53 (used by the standalone planck kernel)
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
// One block of `ch` threads per chunk (ch is a multiple of 32).  Thread i owns local wavenumber i
// and finds its (round, lane, slot): g points in ascending order, each padded to a whole number of
// lanes; inside a g point, the k-th wavenumber goes to lane (k mod L), slot (k div L), so that
// consecutive wavenumbers sit in consecutive lanes.  With a window W > 0 a round is also closed
// before the g point that would make it span W or more consecutive g indices.
struct PlanPos {
  int g;        // -1: belongs to no g point
  int vlane;    // virtual lane over all rounds of the chunk
  int slot;
  int nvlanes;  // total virtual lanes of the chunk (with the window constraint, if any)
  int nvfree;   // total virtual lanes without the window constraint
  int kth;      // index of this wavenumber inside its g point (chunk-local)
  int count;    // wavenumbers of its g point in the chunk
  int L;        // lanes of its g point
};

__device__ __forceinline__ PlanPos plan_position(const int32_t* __restrict__ g_point, int64_t nwav, int ng, int ch,
                                                 int64_t c, int W, int* sg, int* sred) {
  const int i = threadIdx.x;
  const int64_t w = c * ch + i;
  int g = -1;
  if (w < nwav) {
    g = g_point[w];
    if (g < 0 || g >= ng) g = -1;
  }
  sg[i] = g;
  __syncthreads();
  int k = 0, n = 0;
  for (int j = 0; j < ch; ++j) {
    const int same = (sg[j] == g);
    n += same;
    k += same & (j < i);
  }
  const int L = (n + SLOTS - 1) / SLOTS;
  // Walk the distinct g points of the chunk in ascending order (block-wide minimum of the g points
  // not yet placed); every thread keeps the same running state.
  const int lane = i & 31, warp = i >> 5, nwarps = (int)(blockDim.x >> 5);
  int cur = -1, v = 0, v0 = 0, rgmin = -1, my_start = 0;
  while (true) {
    int cand = (g > cur) ? g : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if (lane == 0) sred[warp] = cand;
    __syncthreads();
    int gn = 0x7fffffff;
    for (int q = 0; q < nwarps; ++q) gn = min(gn, sred[q]);
    if (gn == 0x7fffffff) break;  // block-uniform
    if (g == gn && k == 0) sred[32] = L;
    __syncthreads();
    const int Lg = sred[32];
    if (W > 0 && rgmin >= 0 && gn - rgmin >= W) {  // the round would span W g points: close it
      v = (v + 31) & ~31;
      rgmin = -1;
    }
    if (rgmin < 0) rgmin = gn;
    const int start = v;
    v += Lg;
    v0 += Lg;
    if ((v & 31) == 0) rgmin = -1;                      // the next g point opens a round
    else if ((v >> 5) != (start >> 5)) rgmin = gn;      // this g point's tail opened the current round
    if (g == gn) my_start = start;
    cur = gn;
    __syncthreads();  // sred is rewritten by the next trip
  }
  PlanPos p;
  p.g = g;
  p.nvlanes = v;
  p.nvfree = v0;
  p.kth = k;
  p.count = n;
  p.L = L;
  p.vlane = (g >= 0) ? my_start + (k % L) : 0;
  p.slot = (g >= 0) ? (k / L) : 0;
  return p;
}

// rounds per chunk (at least one, so that every chunk visit has a ring entry) with the window
// constraint W (nrounds) and without it (nfree, optional), and, optionally, the wavenumbers per g point
__global__ void __launch_bounds__(CH_MAX) plan_count_kernel(const int32_t* __restrict__ g_point, int64_t nwav, int ng,
                                                             int W, uint32_t* __restrict__ nrounds,
                                                             unsigned long long* __restrict__ nfree_total,
                                                             unsigned long long* __restrict__ count) {
  __shared__ int sg[CH_MAX];
  __shared__ int sred[33];
  const int64_t c = blockIdx.x;
  const PlanPos p = plan_position(g_point, nwav, ng, (int)blockDim.x, c, W, sg, sred);
  if (threadIdx.x == 0) {
    nrounds[c] = (uint32_t)max(1, (p.nvlanes + 31) / 32);
    if (nfree_total != nullptr) atomicAdd(nfree_total, (unsigned long long)max(1, (p.nvfree + 31) / 32));
  }
  if (count != nullptr && p.g >= 0 && p.kth == 0) atomicAdd(&count[p.g], (unsigned long long)p.count);
}

__global__ void plan_init_kernel(int64_t nrounds, PlanRound* __restrict__ rounds, int32_t* __restrict__ rmin,
                                 int32_t* __restrict__ rmax) {
  // unused lanes: offsets 0 (a valid address), no valid slots, meta = NO_G
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrounds * 32) return;
  PlanRound& r = rounds[i >> 5];
  const int lane = (int)(i & 31);
  for (int k = 0; k < SLOTS; ++k) r.off[k][lane] = 0;
  r.meta[lane] = NO_G;
  if (lane == 0) {
    r.steps = 0;
    r.chunk = 0;
    r.flags = 0;
    r.win = 0;
    rmin[i >> 5] = 0x7fffffff;
    rmax[i >> 5] = -1;
  }
}

__global__ void __launch_bounds__(CH_MAX) plan_fill_kernel(const int32_t* __restrict__ g_point, int64_t nwav, int ng,
                                                            int W, const uint32_t* __restrict__ round_off,
                                                            PlanRound* __restrict__ rounds, int32_t* __restrict__ rmin,
                                                            int32_t* __restrict__ rmax) {
  __shared__ int sg[CH_MAX];
  __shared__ int sred[33];
  const int64_t c = blockIdx.x;
  const PlanPos p = plan_position(g_point, nwav, ng, (int)blockDim.x, c, W, sg, sred);
  const uint32_t r0 = round_off[c], r1 = round_off[c + 1];
  if (threadIdx.x == 0) {
    for (uint32_t r = r0; r < r1; ++r) {
      rounds[r].chunk = (uint32_t)c;
      rounds[r].flags = (r == r0 ? 1u : 0u) | (r + 1 == r1 ? 2u : 0u);
    }
  }
  if (p.g < 0) return;
  const size_t rix = (size_t)r0 + (size_t)(p.vlane >> 5);
  PlanRound& rd = rounds[rix];
  const int lane = p.vlane & 31;
  const uint16_t my = (uint16_t)(threadIdx.x * 8);
  auto slot_ref = [&](int slot) -> uint16_t& { return rd.off[slot][lane]; };
  slot_ref(p.slot) = my;
  if (p.slot == 0) {
    // this thread owns the lane: valid slots form a prefix of length nv; padding slots repeat slot 0
    const int li = p.kth;  // lane index inside the g point (kth < L here)
    const int nv = (p.count - 1 - li) / p.L + 1;
    for (int e = nv; e < SLOTS; ++e) slot_ref(e) = my;
    // lanes of the same g point that follow this one inside the round
    const int follow = min(p.L - 1 - li, 31 - lane);
    unsigned mask = 0;
    int steps = 0;
    for (int s = 0; s < 5; ++s) {
      if ((1 << s) <= follow) {
        mask |= 1u << s;
        steps = s + 1;
      }
    }
    const unsigned head = (lane == 0 || li == 0) ? 1u : 0u;
    rd.meta[lane] = (unsigned)p.g | ((unsigned)nv << 12) | (mask << 17) | (head << 22);
    if (steps > 0) atomicMax(&rd.steps, (uint32_t)steps);
    atomicMin(&rmin[rix], p.g);
    atomicMax(&rmax[rix], p.g);
  }
}

// ------------------------------------------------------------------ balanced plans
// The slot loop of the streaming kernels is bound by shared-memory bandwidth, and a third of its
// wavefronts were bank-conflict replays of the permuted 64-bit gathers (a warp's LDS.64 is served as
// lanes 0-15 then 16-31, each half costing as many wavefronts as the fullest of the 16 eight-byte bank
// pairs: tools/ubench/lds64_conflicts.cu).  reduce2/3_kernel add their lanes' sums through per-lane
// entries, so the lanes of a g point need not be adjacent and ANY placement of a g point's
// wavenumbers on its lanes' cells is allowed.  plan_fill_balanced_kernel uses that freedom:
//   * the lanes of every g point of a round alternate between the two half-warps (virtual lane v ->
//     physical lane v/2 + 16 (v & 1));
//   * the wavenumbers of a (g point, round) are dealt to the two half-warps residue class by residue
//     class (class = chunk index mod 16 = bank pair), by error diffusion, so that each half-warp holds
//     its share of every class: over a round no class then exceeds 16 wavenumbers per half-warp
//     (a chunk has at most 30 of each), except for rounding;
//   * inside a half-warp the wavenumbers are edges of a bipartite multigraph lanes x residue classes
//     of maximum degree 16: a proper edge colouring with 16 colours (Koenig) exists, and the colour is
//     the slot - every slot then reads 16 different bank pairs.  It is built edge by edge with the
//     alternating-path flip; an edge whose class already has 16 edges keeps a lane-proper colour
//     and is the only source of conflicts left;
//   * padding cells repeat an address another lane of the same g point reads in that slot
//     (broadcast), else an own wavenumber of an unused class.
// The valid slots of a lane are no longer a prefix: bits 12-27 of the lane's meta word are a slot mask.
constexpr int BAL_TASKS = 8;  // (round, half-warp) colouring tasks per batch

__global__ void __launch_bounds__(CH_MAX) plan_fill_balanced_kernel(const int32_t* __restrict__ g_point, int64_t nwav,
                                                                     int ng, int W,
                                                                     const uint32_t* __restrict__ round_off,
                                                                     PlanRound* __restrict__ rounds,
                                                                     int32_t* __restrict__ rmin,
                                                                     int32_t* __restrict__ rmax,
                                                                     unsigned* __restrict__ g_sinfo) {
  __shared__ int sg[CH_MAX];
  __shared__ int sred[33];
  // per wavenumber: bits 0-9 chunk-local round, 10 half-warp, 11-15 physical lane, 16-19 colour (slot),
  // 20 "half edge" (its class was full: not registered on the class side), 31 belongs to a g point
  __shared__ unsigned sinfo[CH_MAX];
  __shared__ short slead[CH_MAX];             // first wavenumber (thread) of the same (g point, round)
  __shared__ unsigned char scls[CH_MAX][16];  // [leader][class]: wavenumbers of the (g point, round) per class
  __shared__ unsigned char sa0[CH_MAX][16];   // [leader][class]: how many of them go to half-warp 0
  const int i = threadIdx.x;
  const int ch = (int)blockDim.x;
  const int64_t c = blockIdx.x;
  const PlanPos p = plan_position(g_point, nwav, ng, ch, c, W, sg, sred);
  const uint32_t r0 = round_off[c], r1 = round_off[c + 1];
  if (i == 0) {
    for (uint32_t r = r0; r < r1; ++r) {
      rounds[r].chunk = (uint32_t)c;
      rounds[r].flags = (r == r0 ? 1u : 0u) | (r + 1 == r1 ? 2u : 0u);
    }
  }
  const bool valid = p.g >= 0;
  // the (g point, round) of this wavenumber: the g point's capacity lanes, 16 wavenumbers each in
  // wavenumber order, map to virtual lanes gs .. gs+L-1; round = virtual lane / 32
  int gs = 0, rl = 0, l = 1, n = 0, a = 0;
  if (valid) {
    gs = p.vlane - (p.kth % p.L);
    rl = (gs + (p.kth >> 4)) >> 5;
    const int c0 = max(0, 32 * rl - gs);                          // capacity lanes of the g point before this round
    l = min(gs + p.L, 32 * rl + 32) - max(gs, 32 * rl);           // its lanes in this round
    n = min(p.count - 16 * c0, 16 * l);                           // its wavenumbers in this round
    a = max(gs, 32 * rl) - 32 * rl;                               // virtual position of its first lane
  }
  sinfo[i] = valid ? (0x80000000u | (unsigned)rl) : 0u;
  for (int q = 0; q < 16; ++q) scls[i][q] = 0;
  __syncthreads();
  int lead = i, t = 0, m = 0;
  if (valid) {
    for (int j = 0; j < ch; ++j) {
      if (sg[j] == p.g && (sinfo[j] & 0x3ffu) == (unsigned)rl) {
        lead = j;
        break;
      }
    }
    for (int j = (i & 15); j < ch; j += 16) {
      const int same = (sg[j] == p.g && (sinfo[j] & 0x3ffu) == (unsigned)rl) ? 1 : 0;
      m += same;
      t += same & (j < i);
    }
  }
  slead[i] = (short)lead;
  if (valid && t == 0) scls[lead][i & 15] = (unsigned char)m;
  __syncthreads();
  const int l0 = ((a + l + 1) >> 1) - ((a + 1) >> 1);  // even virtual lanes in [a, a+l): half-warp 0
  const int l1 = l - l0;
  if (valid && i == lead) {
    int n0 = (n * l0 + l / 2) / l;
    n0 = min(max(n0, n - 16 * l1), 16 * l0);
    int M = 0, A = 0;
    for (int q = 0; q < 16; ++q) {
      M += scls[i][q];
      const int tgt = (n0 * M + n / 2) / n;  // rounded share of half-warp 0 after this class
      sa0[i][q] = (unsigned char)(tgt - A);
      A = tgt;
    }
    PlanRound& rd = rounds[r0 + rl];
    for (int v = a; v < a + l; ++v) rd.meta[(v >> 1) + 16 * (v & 1)] = (unsigned)p.g;  // slot mask OR-ed in later
    atomicMin(&rmin[r0 + rl], p.g);
    atomicMax(&rmax[r0 + rl], p.g);
  }
  __syncthreads();
  if (valid) {
    const int q = i & 15;
    const int a0 = sa0[lead][q];
    const int h = (t < a0) ? 0 : 1;
    int u = h ? t - a0 : t;
    for (int qq = 0; qq < q; ++qq) u += h ? (int)scls[lead][qq] - (int)sa0[lead][qq] : (int)sa0[lead][qq];
    const int lh = h ? l1 : l0;  // >= 1 whenever the half-warp got a wavenumber
    const int v = a + ((a & 1) ^ h) + 2 * (u % lh);
    const int phys = (v >> 1) + 16 * h;
    sinfo[i] |= ((unsigned)h << 10) | ((unsigned)phys << 11);
  }
  __syncthreads();
  // the colouring of the (round, half-warp) cells is serial per cell set: plan_colour_kernel gives every set a warp
  g_sinfo[(size_t)c * ch + i] = sinfo[i];
}

// One (round, half-warp) colouring task per warp: lane 0 runs the serial edge colouring, then 16 lanes write the
// cells of one slot each.  (Inside
// plan_fill_balanced_kernel the two tasks of a chunk kept 14 warps of its block waiting at a barrier - 80 % of the
// kernel's stall samples - and only 4 tasks ran per SM; here up to 32 do.)
constexpr int COL_WARPS = 8;
__global__ void __launch_bounds__(COL_WARPS * 32) plan_colour_kernel(int64_t nrounds, int ch,
                                                                       const uint32_t* __restrict__ round_off,
                                                                       PlanRound* __restrict__ rounds,
                                                                       const unsigned* __restrict__ g_sinfo) {
  __shared__ unsigned s_info[COL_WARPS][CH_MAX];
  __shared__ short scolL[COL_WARPS][16][16];  // [warp][lane][colour] -> wavenumber or -1
  __shared__ short scolR[COL_WARPS][16][16];  // [warp][class][colour] -> wavenumber or -1
  __shared__ int slane_g[COL_WARPS][16];      // [warp][lane]: g point of the lane, NO_G when unused
  __shared__ unsigned sused[COL_WARPS][32];   // [warp][lane | 16 + class]: bit mask of the colours taken (the occupancy of
                                              // scolL / scolR as bits, so that "lowest free colour" is one find-first-set)
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = (int64_t)blockIdx.x * COL_WARPS + wib;
  if (task >= 2 * nrounds) return;
  const int64_t r = task >> 1;
  const int ht = (int)(task & 1);
  const int64_t c = rounds[r].chunk;
  const uint32_t r0 = round_off[c];
  const int rt = (int)(r - r0);
  unsigned* sinfo = s_info[wib];
  for (int x = lane; x < ch; x += 32) sinfo[x] = g_sinfo[(size_t)c * ch + x];
  for (int x = lane; x < 256; x += 32) {
    (&scolL[wib][0][0])[x] = -1;
    (&scolR[wib][0][0])[x] = -1;
  }
  __syncwarp();
  if (lane == 0) {
    short (*CL)[16] = scolL[wib];
    short (*CR)[16] = scolR[wib];
    unsigned* UL = sused[wib];
    unsigned* UR = sused[wib] + 16;
    for (int x = 0; x < 32; ++x) UL[x] = 0u;
    auto colour_of = [&](int e) { return (int)((sinfo[e] >> 16) & 15u); };
    auto is_half = [&](int e) { return ((sinfo[e] >> 20) & 1u) != 0; };
    auto lane_of = [&](int e) { return (int)((sinfo[e] >> 11) & 15u); };
    auto set_colour = [&](int e, int col) { sinfo[e] = (sinfo[e] & ~(15u << 16)) | ((unsigned)col << 16); };
    for (int e = 0; e < ch; ++e) {
      const unsigned inf = sinfo[e];
      if (!(inf >> 31) || (inf & 0x3ffu) != (unsigned)rt || ((inf >> 10) & 1u) != (unsigned)ht) continue;
      const int ln = (int)((inf >> 11) & 15u), q = e & 15;
      int ca = __ffs(~UL[ln] & 0xffffu) - 1, cb = __ffs(~UR[q] & 0xffffu) - 1;  // lowest free colour at the lane / class
      if (cb < 0) cb = 16;
      if (ca < 0) ca = 15;  // cannot happen: a lane holds at most 16 wavenumbers
      // a colour free at both ends needs no alternating-path flip (with the lowest colour free at the lane alone,
      // 53 % of the edges flipped a path and the kernel took three times as long)
      if (const unsigned both = ~UL[ln] & ~UR[q] & 0xffffu) ca = __ffs(both) - 1;
      if (cb == 16) {         // the class already has 16 wavenumbers in this half-warp: lane-proper only
        CL[ln][ca] = (short)e;
        UL[ln] |= 1u << ca;
        sinfo[e] |= 1u << 20;
        set_colour(e, ca);
        continue;
      }
      if (CR[q][ca] >= 0) {
        // colour ca is taken at the class by another lane: flip the ca/cb alternating path that
        // starts there (it cannot reach lane ln, which has no ca edge)
        short path[40];
        int np = 0;
        int x = CR[q][ca], cx = ca;
        while (x >= 0 && np < 40) {
          path[np++] = (short)x;
          if (cx == ca) {
            x = CL[lane_of(x)][cb];
            cx = cb;
          } else {
            if (is_half(x)) break;
            x = CR[x & 15][ca];
            cx = ca;
          }
        }
        for (int k = 0; k < np; ++k) {
          const int y = path[k], cy = colour_of(y);
          CL[lane_of(y)][cy] = -1;
          UL[lane_of(y)] &= ~(1u << cy);
          if (!is_half(y)) {
            CR[y & 15][cy] = -1;
            UR[y & 15] &= ~(1u << cy);
          }
        }
        for (int k = 0; k < np; ++k) {
          const int y = path[k], cy = colour_of(y) == ca ? cb : ca;
          CL[lane_of(y)][cy] = (short)y;
          UL[lane_of(y)] |= 1u << cy;
          if (!is_half(y)) {
            CR[y & 15][cy] = (short)y;
            UR[y & 15] |= 1u << cy;
          }
          set_colour(y, cy);
        }
      }
      if (CL[ln][ca] >= 0 || CR[q][ca] >= 0) {
        // (defensive: the flip could not free the colour - keep the plan valid with a lane-proper colour)
        int cf = __ffs(~UL[ln] & 0x7fffu) - 1;
        if (cf < 0) cf = 15;
        CL[ln][cf] = (short)e;
        UL[ln] |= 1u << cf;
        sinfo[e] |= 1u << 20;
        set_colour(e, cf);
        continue;
      }
      CL[ln][ca] = (short)e;
      CR[q][ca] = (short)e;
      UL[ln] |= 1u << ca;
      UR[q] |= 1u << ca;
      set_colour(e, ca);
    }
  }
  __syncwarp();
  // write the half-warp's 16 x 16 cells: the slots are independent of each other (a slot reads the colour tables of
  // all the lanes but writes only its own column of the class table), so that lane s takes slot s
  short (*CL)[16] = scolL[wib];
  short (*CR)[16] = scolR[wib];
  const unsigned* UL = sused[wib];
  int* lane_g = slane_g[wib];
  PlanRound& rd = rounds[r0 + rt];
  if (lane < 16)  // a lane of the round without a wavenumber in this half-warp is unused
    lane_g[lane] = UL[lane] ? (int)(rd.meta[lane + 16 * ht] & 0xFFFu) : (int)NO_G;
  __syncwarp();
  if (lane >= 16) return;
  {
    const int s = lane;
    int shared_e = -1;  // what the unused lanes of the half-warp read: the address of a used lane (broadcast)
    for (int ln = 0; ln < 16; ++ln) {
      if (lane_g[ln] == (int)NO_G) continue;
      int e = CL[ln][s];
      if (e < 0) {
        // padding: an address another lane of this g point reads in this slot, else an own
        // wavenumber of a class nobody reads in this slot, else any own wavenumber
        for (int l2 = 0; l2 < 16 && e < 0; ++l2)
          if (lane_g[l2] == lane_g[ln] && CL[l2][s] >= 0) e = CL[l2][s];
        int own = -1;
        for (int s2 = 0; s2 < 16 && e < 0; ++s2) {
          const int y = CL[ln][s2];
          if (y < 0) continue;
          if (own < 0) own = y;
          if (CR[y & 15][s] < 0) {
            e = y;
            CR[y & 15][s] = (short)y;  // the class is now read in this slot
          }
        }
        if (e < 0) e = own;
      }
      rd.off[s][ln + 16 * ht] = (uint16_t)(e * 8);
      if (shared_e < 0) shared_e = e;
    }
    // an unused lane still issues its loads (weight zero, sums discarded): offset 0 would add an
    // address of bank pair 0 to every slot, i.e. a conflict whenever a used lane reads that pair
    if (shared_e >= 0) {
      for (int ln = 0; ln < 16; ++ln)
        if (lane_g[ln] == (int)NO_G) rd.off[s][ln + 16 * ht] = (uint16_t)(shared_e * 8);
    }
  }
  // the mask of the slots that hold a wavenumber of the lane is the occupancy word of its colours
  rd.meta[lane + 16 * ht] =
      lane_g[lane] == (int)NO_G ? NO_G : ((unsigned)lane_g[lane] | ((UL[lane] & 0xffffu) << 12));
}

// Windowed plans: the window base of every round, in round order (one warp; a round keeps the
// previous window when its g points fit in it, else the window moves to the round's smallest g point).
__global__ void plan_window_kernel(int64_t nrounds, int W, const int32_t* __restrict__ rmin,
                                   const int32_t* __restrict__ rmax, PlanRound* __restrict__ rounds) {
  const int lane = threadIdx.x;
  int gbase = 0;
  // the chain over the rounds is serial: keep the memory latencies off it (the next 32 bounds are fetched while the
  // current ones are folded, and the flag is set by a reduction that nobody waits for)
  int mn_n = (lane < nrounds) ? rmin[lane] : 0x7fffffff;
  int mx_n = (lane < nrounds) ? rmax[lane] : -1;
  for (int64_t base = 0; base < nrounds; base += 32) {
    const int64_t r = base + lane;
    const int mn_l = mn_n, mx_l = mx_n;
    mn_n = (r + 32 < nrounds) ? rmin[r + 32] : 0x7fffffff;
    mx_n = (r + 32 < nrounds) ? rmax[r + 32] : -1;
    int mine = 0, same = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int mn = __shfl_sync(0xffffffffu, mn_l, j), mx = __shfl_sync(0xffffffffu, mx_l, j);
      const int before = gbase;
      if (mn <= mx && !(mn >= gbase && mx < gbase + W)) gbase = mn;
      if (lane == j) {
        mine = gbase;
        same = (base + j > 0 && gbase == before) ? 1 : 0;
      }
    }
    if (r < nrounds) {
      rounds[r].win = (uint32_t)mine;
      if (same) atomicOr(&rounds[r].flags, 4u);  // bit 2: same window as the round before
    }
  }
}

// exclusive scan of n uint32 by one block (n is a few 10^4 .. 10^5)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(int64_t n, const uint32_t* __restrict__ in,
                                                              uint32_t* __restrict__ out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = (i < n) ? in[i] : 0u;
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
      s_warp[lane] = t;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t warp_off = (warp == 0) ? 0u : s_warp[warp - 1];
    if (i < n) out[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_carry;
}

// ------------------------------------------------------------------ reduce
// Shared-memory layout of one CTA (byte offsets computed identically on the host)
struct SmemLayout {
  int bars;    // mbarriers: full_d[nw][NSETS*RPW], full_w[NSW], empty_w[NSW]
  int ring;    // ring slots [NSW]: {double tile[CH_MAX]; PlanRound plan}
  int data;    // data stages [nw][NSETS*RPW][CH_MAX] doubles
  int acc;     // accumulators [nw][RPW][ng][4] doubles: {sum f*w, sum w, min, max}
  int accz;    // zero counters [nw][RPW][ng] int32
  int tab;     // 2^(j/256) table of the consumer warps (trans_core.cuh)
  int total;
};
static_assert(RING_SLOT_BYTES % 16 == 0 && STAGE_BYTES % 16 == 0, "stage alignment");
__host__ __device__ inline SmemLayout smem_layout(int nw, int ng) {
  SmemLayout L;
  int off = 0;
  L.bars = off;
  off += (nw * NSETS * RPW + 2 * NSW) * 8;
  off = (off + 127) & ~127;
  L.ring = off;
  off += NSW * RING_SLOT_BYTES;
  L.data = off;
  off += nw * NSETS * RPW * STAGE_BYTES;
  L.acc = off;
  off += nw * RPW * ng * 4 * 8;
  L.accz = off;
  off += nw * RPW * ng * 4;
  off = (off + 15) & ~15;
  L.tab = off;
  off += TRANS_TAB_BYTES;
  L.total = (off + 15) & ~15;
  return L;
}

// position of a CTA in the (window, group, chunk) item order
struct Cursor {
  int v, l, c, c_begin, c_end;
};
__device__ __forceinline__ Cursor cursor_at(int64_t item, int ngroups, int cw, int nchunks) {
  Cursor q;
  const int64_t per_win = (int64_t)ngroups * cw;  // every window but the last holds cw chunks
  q.v = (int)(item / per_win);
  const int64_t rem = item - (int64_t)q.v * per_win;
  q.c_begin = q.v * cw;
  q.c_end = min(nchunks, q.c_begin + cw);
  const int n_v = q.c_end - q.c_begin;
  q.l = (int)(rem / n_v);
  q.c = q.c_begin + (int)(rem - (int64_t)q.l * n_v);
  return q;
}
// returns true when the step entered a new (window, group)
__device__ __forceinline__ bool cursor_next(Cursor& q, int ngroups, int cw, int nchunks) {
  if (++q.c < q.c_end) return false;
  if (++q.l == ngroups) {
    q.l = 0;
    ++q.v;
    q.c_begin = q.v * cw;
    q.c_end = min(nchunks, q.c_begin + cw);
  }
  q.c = q.c_begin;
  return true;
}

// The producer warps of the streaming kernels (reduce_kernel, reduce2_kernel): permuted weights +
// plan record of every round of the CTA's piece, through the ring {wperm, PlanRound} x NSW at
// shared-window address ring_s (generic pointer `ring`), full/empty mbarriers full_w/empty_w.
template <int WMODE, bool PLANCK_TAB = false>
__device__ __forceinline__ void produce_rounds(const ReduceParams& p, const Cursor first, const int niter, const int pw,
                                               const int lane, unsigned char* ring, const unsigned ring_s,
                                               const unsigned full_w, const unsigned empty_w, const int nsw = NSW,
                                               const unsigned char* tab = nullptr) {
  const int nchunks = (int)p.nchunks;
  const int ch = p.ch;
  // ===================== producers: permuted weights + plan record of every round =====================
  // Producer lane `lane` of warp pw fills wperm[k][lane] for the slots k = pw, pw + NPW, ...: the
  // weight of the wavenumber the plan puts there (zero for padding slots).  Consecutive
  // wavenumbers sit in consecutive lanes of a slot, so the gathers below are nearly coalesced.
  // Software pipeline over ring entries (= plan rounds): plan offsets are fetched two entries
  // ahead, wavenumbers / weights one entry ahead.
  constexpr int PER = SLOTS / NPW;  // slots per producer lane per round
  static_assert(SLOTS % NPW == 0, "slots must split evenly over the producer warps");
  const TransLane ptl = trans_lane_init();
  struct Entry {
    uint32_t r;       // plan round
    int64_t w0;       // first wavenumber of its chunk
    int wrow;         // weight row (temperature index / matrix row)
    bool valid;
  };
  Cursor q = first;
  int visit = 0;
  uint32_t r_cur = p.round_off[q.c], r_end = p.round_off[q.c + 1];
  auto next_entry = [&]() {
    Entry e;
    e.valid = visit < niter;
    e.r = r_cur;
    e.w0 = (int64_t)q.c * ch;
    e.wrow = e.valid ? p.groups[q.l].wrow : 0;
    if (e.valid && ++r_cur == r_end) {
      ++visit;
      if (visit < niter) {
        cursor_next(q, p.ngroups, p.cw, nchunks);
        r_cur = p.round_off[q.c];
        r_end = p.round_off[q.c + 1];
      }
    }
    return e;
  };
  auto load_plan = [&](const Entry& e, unsigned (&off)[PER], unsigned& meta) {
    if (!e.valid) return;
    const PlanRound* pr = p.rounds + e.r;
    meta = pr->meta[lane];
#pragma unroll
    for (int m = 0; m < PER; ++m) off[m] = pr->off[pw + NPW * m][lane];
  };
  auto load_src = [&](const Entry& e, const unsigned (&off)[PER], double (&x)[PER], double (&y)[PER]) {
    if (!e.valid) return;
#pragma unroll
    for (int m = 0; m < PER; ++m) {
      const int64_t w = e.w0 + (off[m] >> 3);
      if (WMODE == W_PLANCK) {
        x[m] = p.wn[w];
        y[m] = PLANCK_TAB ? p.pref[w] : p.dwn[w];
      } else if (WMODE == W_MATRIX) {
        x[m] = p.weight[(int64_t)e.wrow * p.w_stride + w];
      } else {
        x[m] = p.weight[w];
      }
    }
  };
  Entry e0 = next_entry(), e1 = next_entry();
  unsigned off1[PER], meta0 = NO_G, meta1 = NO_G, meta2 = NO_G;
  double x0[PER], y0[PER], x1[PER], y1[PER];
  load_plan(e0, off1, meta0);
  load_src(e0, off1, x0, y0);
  load_plan(e1, off1, meta1);
  int ws = 0, filled = 0;
  unsigned epar = 0;
  while (e0.valid) {
    const Entry e2 = next_entry();
    load_src(e1, off1, x1, y1);
    load_plan(e2, off1, meta2);  // off1 is free again: the gathers above have their addresses
    // valid slots: a prefix of nv slots, or (balanced plans) the slot mask in bits 12-27
    const unsigned vmask = p.balanced ? ((meta0 >> 12) & 0xFFFFu) : ((1u << ((meta0 >> 12) & 31u)) - 1u);
    double val[PER];
    if (WMODE == W_PLANCK) {
#ifdef ECCKD_PLANCK_TAB
      if (PLANCK_TAB) {
        const double hkct = p.temperature2[e0.wrow];  // (h/k)(100 c)/T, formed on the host
#pragma unroll
        for (int m = 0; m < PER; ++m) val[m] = planck_weight_tab(hkct, x0[m], y0[m], tab);
      } else
#endif
      {
        const double hkt = p.temperature[e0.wrow];  // (h/k)/T, formed on the host
#pragma unroll
        for (int m = 0; m < PER; ++m) val[m] = planck_elem_fast(hkt, x0[m], y0[m], ptl);
      }
    } else {
#pragma unroll
      for (int m = 0; m < PER; ++m) val[m] = x0[m];
    }
    unsigned char* slot = ring + ws * RING_SLOT_BYTES;
    double* wperm = reinterpret_cast<double*>(slot);
    if (filled >= nsw) mbar_wait(empty_w + ws * 8, epar);
#pragma unroll
    for (int m = 0; m < PER; ++m) {
      const int k = pw + NPW * m;
      wperm[k * 32 + lane] = ((vmask >> k) & 1u) ? val[m] : 0.0;
    }
    __syncwarp();
    if (lane == 0) {
      if (pw == 0) {
        // (no proxy fence: the consumers only READ the old record through the generic proxy, and
        // their release of the slot is ordered before this copy by the empty barrier)
        mbar_arrive_expect_tx(full_w + ws * 8, (unsigned)sizeof(PlanRound));
        bulk_g2s(ring_s + ws * RING_SLOT_BYTES + WPERM_BYTES, p.rounds + e0.r, (unsigned)sizeof(PlanRound),
                 full_w + ws * 8);
      } else {
        mbar_arrive(full_w + ws * 8);
      }
    }
    ++filled;
    if (++ws == nsw) {
      ws = 0;
      if (filled > nsw) epar ^= 1u;  // the first pass over the ring needs no wait; parity 0 comes first
    }
    e0 = e1;
    e1 = e2;
    meta0 = meta1;
    meta1 = meta2;
#pragma unroll
    for (int m = 0; m < PER; ++m) {
      x0[m] = x1[m];
      y0[m] = y1[m];
    }
  }
}

template <int FKIND, int WMODE>
__global__ void __launch_bounds__((NW_MAX + NPW) * 32, 1) reduce_kernel(const __grid_constant__ ReduceParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nw = p.nw, ng = p.ng;
  const SmemLayout L = smem_layout(nw, ng);
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned full_d = sbase + L.bars;                // mbarriers [nw][NSETS*RPW], 8 bytes each
  const unsigned full_w = full_d + nw * NSETS * RPW * 8; // [NSW]
  const unsigned empty_w = full_w + NSW * 8;             // [NSW]
  unsigned char* ring = smem_raw + L.ring;
  constexpr bool LOGKIND = (FKIND == F_LOG || FKIND == F_HYBRID);
  constexpr bool USES_TAU = (FKIND != F_NONE);

  if (threadIdx.x == 0) {
    for (int i = 0; i < nw * NSETS * RPW; ++i) mbar_init(full_d + i * 8, 1);
    for (int i = 0; i < NSW; ++i) {
      mbar_init(full_w + i * 8, NPW);
      mbar_init(empty_w + i * 8, nw);
    }
    fence_mbar_init();
  }
  if (FKIND == F_TRANS) trans_tab_init(reinterpret_cast<double*>(smem_raw + L.tab));
  __syncthreads();
  const unsigned char* tab = smem_raw + L.tab;

  // this CTA's contiguous piece of the (window, group, chunk) space
  const int nchunks = (int)p.nchunks;
  const int ch = p.ch;
  const int64_t T = (int64_t)p.ngroups * p.nchunks;
  const int64_t lo = (T * blockIdx.x) / p.ncta;
  const int64_t hi = (T * (blockIdx.x + 1)) / p.ncta;
  const int niter = (int)(hi - lo);
  if (niter == 0) return;
  const Cursor first = cursor_at(lo, p.ngroups, p.cw, nchunks);
  const int last_n = (int)(p.nwav - (int64_t)(nchunks - 1) * ch);  // length of the last chunk of a row

  if (warp >= nw) {
    produce_rounds<WMODE>(p, first, niter, warp - nw, lane, ring, sbase + L.ring, full_w, empty_w);
    return;
  }

  // ===================== consumers: warp = RPW rows of the group =====================
  const TransLane tl = trans_lane_init();
  unsigned char* data = smem_raw + L.data + (size_t)warp * NSETS * RPW * STAGE_BYTES;
  double* acc = reinterpret_cast<double*>(smem_raw + L.acc) + (size_t)warp * RPW * ng * 4;
  int32_t* accz = reinterpret_cast<int32_t*>(smem_raw + L.accz) + (size_t)warp * RPW * ng;
  const unsigned my_full = full_d + warp * NSETS * RPW * 8;
  const unsigned my_data = sbase + L.data + warp * NSETS * RPW * STAGE_BYTES;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const uint64_t pol = policy_evict_first();
  // the weight sums do not depend on the row: the last consumer warp (the one with the fewest rows)
  // accumulates them once per group into the s1 field of its row 0, except where they depend on
  // the row (logarithmic kinds)
  const bool s1_duty = !LOGKIND && warp == nw - 1;

  auto init_acc = [&]() {
    for (int i = lane; i < RPW * ng; i += 32) {
      acc[i * 4 + 0] = 0.0;
      acc[i * 4 + 1] = 0.0;
      acc[i * 4 + 2] = inf;
      acc[i * 4 + 3] = -inf;
      accz[i] = 0;
    }
    __syncwarp();
  };
  auto flush_partial = [&](int vg) {
    __syncwarp();
    const size_t slot = ((size_t)blockIdx.x * p.maxg + (size_t)vg) * (nw * RPW) + warp * RPW;
    double* out = p.partial + slot * ng * 4;
    for (int i = lane; i < RPW * ng * 4; i += 32) out[i] = acc[i];
    int32_t* outz = p.partial_z + slot * ng;
    for (int i = lane; i < RPW * ng; i += 32) outz[i] = accz[i];
    __syncwarp();
  };

  // issue cursor (runs NSETS chunk visits ahead of the compute cursor)
  Cursor qi = first;
  const double* od_i[RPW];
  bool bulk_i[RPW];
  auto load_issue_group = [&]() {
    const GroupDesc* gd = p.groups + qi.l;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = warp + r * nw;
      const bool act = USES_TAU && i < gd->nrows;
      od_i[r] = act ? gd->od[i] : nullptr;
      bulk_i[r] = act && gd->bulk_ok[i];
    }
  };
  auto issue = [&](int set) {
    const int n_c = (qi.c == nchunks - 1) ? last_n : ch;
    if (lane == 0 && ((n_c & 1) == 0)) {
      // the rows of a set complete on ONE mbarrier (one wait per chunk visit instead of one per row)
      const unsigned bar = my_full + (set * RPW) * 8;
      int nb = 0;
#pragma unroll
      for (int r = 0; r < RPW; ++r) nb += bulk_i[r] ? 1 : 0;
      if (nb > 0) {
        mbar_arrive_expect_tx(bar, (unsigned)(nb * n_c * 8));
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          if (bulk_i[r])
            bulk_g2s_hint(my_data + (set * RPW + r) * STAGE_BYTES, od_i[r] + (int64_t)qi.c * ch, (unsigned)(n_c * 8), bar, pol);
        }
      }
    }
    if (cursor_next(qi, p.ngroups, p.cw, nchunks)) load_issue_group();
  };
  load_issue_group();
  for (int j = 0; j < NSETS && j < niter; ++j) issue(j);

  Cursor q = first;
  int ws = 0, vg = 0;
  unsigned wpar = 0;
  unsigned phase_bits = 0;  // bit `set`: parity to wait for on that stage set's mbarrier
  const double* od_c[RPW];
  bool active[RPW], bulk_row[RPW], is_log[RPW];
  bool need_group = true;

  // one chunk visit on stage set `set`
  auto visit = [&](const int set, const int k) {
    if (need_group) {
      const GroupDesc* gd = p.groups + q.l;
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int i = warp + r * nw;
        active[r] = USES_TAU && i < gd->nrows;
        od_c[r] = active[r] ? gd->od[i] : nullptr;
        bulk_row[r] = active[r] && gd->bulk_ok[i];
        is_log[r] = (FKIND == F_LOG) || (FKIND == F_HYBRID && active[r] && gd->is_log[i] != 0);
      }
      init_acc();
      need_group = false;
    }
    const int n_c = (q.c == nchunks - 1) ? last_n : ch;
    bool any_bulk = false;
#pragma unroll
    for (int r = 0; r < RPW; ++r) any_bulk = any_bulk || bulk_row[r];
    if (any_bulk && ((n_c & 1) == 0)) {
      mbar_wait(my_full + (set * RPW) * 8, (phase_bits >> set) & 1u);
      phase_bits ^= 1u << set;
    }
    bool generic_fill = false;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int st = set * RPW + r;
      if (!(bulk_row[r] && ((n_c & 1) == 0)) && active[r]) {
        const double* src = od_c[r] + (int64_t)q.c * ch;
        double* stage = reinterpret_cast<double*>(data + st * STAGE_BYTES);
        for (int j = lane; j < n_c; j += 32) stage[j] = src[j];  // unaligned rows / odd tails
        generic_fill = true;
      }
    }
    __syncwarp();
    bool last_round = false;
    while (!last_round) {
      mbar_wait(full_w + ws * 8, wpar);
      const unsigned char* slot = ring + ws * RING_SLOT_BYTES;
      const PlanRound* pr = reinterpret_cast<const PlanRound*>(slot + WPERM_BYTES);
      const double* wp = reinterpret_cast<const double*>(slot) + lane;
      const uint16_t* offp = &pr->off[0][lane];
      const unsigned meta = pr->meta[lane];
      const int nsteps = (int)pr->steps;
      last_round = (pr->flags & 2u) != 0;
      const int nv = (int)((meta >> 12) & 31u);
      const bool lane_used = (meta & 0xFFFu) != NO_G;
      const unsigned gsafe = lane_used ? (meta & 0xFFFu) : 0u;
      if (s1_duty) {
        Partial a[1];
        partial_begin(a[0], acc, false);
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < SLOTS; ++e) s += wp[e * 32];
        a[0].s1 = s;
        fold_partials<1, false, true, false, false>(a, meta, nsteps, acc, accz, ng);
        __syncwarp();
      }
      if (USES_TAU && active[0] && !(p.debug & 1)) {
        const unsigned char* stage0 = data + (set * RPW) * STAGE_BYTES;
        if (active[RPW - 1]) {
          Partial a[RPW];
#pragma unroll
          for (int r = 0; r < RPW; ++r) partial_begin(a[r], acc + ((size_t)r * ng + gsafe) * 4, lane_used);
          if (LOGKIND) {
            rows_slots<FKIND, RPW>(a, stage0, offp, wp, nv, p.tc, tl, is_log, p.tc256, tab);
            fold_partials<RPW, true, true, true, true>(a, meta, nsteps, acc, accz, ng);
          } else {
            if (!(p.debug & 4)) rows_slots<FKIND, RPW>(a, stage0, offp, wp, nv, p.tc, tl, is_log, p.tc256, tab);
            fold_partials<RPW, true, false, true, false>(a, meta, nsteps, acc, accz, ng);
          }
        } else {
          Partial a[1];
          const bool il[1] = {is_log[0]};
          partial_begin(a[0], acc + (size_t)gsafe * 4, lane_used);
          if (LOGKIND) {
            rows_slots<FKIND, 1>(a, stage0, offp, wp, nv, p.tc, tl, il, p.tc256, tab);
            fold_partials<1, true, true, true, true>(a, meta, nsteps, acc, accz, ng);
          } else {
            if (!(p.debug & 4)) rows_slots<FKIND, 1>(a, stage0, offp, wp, nv, p.tc, tl, il, p.tc256, tab);
            fold_partials<1, true, false, true, false>(a, meta, nsteps, acc, accz, ng);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_w + ws * 8);
      if (++ws == NSW) {
        ws = 0;
        wpar ^= 1u;
      }
    }
    // re-arm the set for the visit after next.  The stage was only READ through the generic proxy
    // (those loads completed before the __syncwarp above: their values were used) unless this visit
    // filled it with generic stores (unaligned rows, odd tails): then a proxy fence orders those
    // stores before the asynchronous-proxy writes of the next bulk copy into the same stage.
    if (generic_fill) {
      fence_proxy_async();
      __syncwarp();
    }
    if (k + NSETS < niter) issue(set);
    if (cursor_next(q, p.ngroups, p.cw, nchunks)) {
      flush_partial(vg);
      ++vg;
      need_group = true;
    }
  };

#pragma unroll 1
  for (int k = 0; k < niter; ++k) visit(k & (NSETS - 1), k);
  if (!need_group) flush_partial(vg);
}

// ------------------------------------------------------------------ reduce, windowed plans
// reduce2_kernel: the same streaming structure as reduce_kernel (bulk-copied stages, producer ring,
// one (lane, slot) plan round at a time) for windowed plans and the kinds without per-row weight
// sums (linear, transmission-k, square-root, weight sums only).  What differs is everything that
// happens once per round, which took as long as the slot loop itself in reduce_kernel (ncu: 48 % of
// the consumer warps' time, profiles/README.md):
//   * no shuffle fold: a lane adds its round sums to ITS OWN shared-memory entry
//     ent[g - window][lane] (one 16-byte read-modify-write for both rows, conflict-free, no
//     synchronisation); the 32 lanes of an entry are only added up - butterfly, fixed order -
//     when the window of g points moves (a band change) or the CTA's piece ends;
//   * the running extremes of a (row, g point) AND the two words of their range check are cached in
//     shared memory (mmv, mmb): a round starts with two loads per row instead of recomputing them,
//     and they are only rewritten in the rare rounds in which some lane found a new extreme;
//   * the per-(window, group) results go straight to the CTA's partial slot in global memory
//     (no [row][ng][4] accumulators in shared memory).
// The summation order is fixed by the plan and the launch geometry: bitwise reproducible.
struct Smem2Layout {
  int tab;   // replicated 2^(j/64) table (trans_core.cuh), at offset 0
  int bars;  // mbarriers: full_d[nw][NSETS], full_w[NSW], empty_w[NSW]
  int ring;  // ring slots [NSW]: {double wperm[SLOTS*32]; PlanRound plan}
  int data;  // data stages [nw][NSETS*RPW][CH_MAX] doubles
  int ent;   // [nw][WIN][32] double2: per-lane sums of the window's g points, both rows
  int s1e;   // [WIN][32] double: per-lane weight sums (last consumer warp)
  int mmv;   // [nw][RPW][WIN] double2 {min, max}
  int mmb;   // [nw][RPW][WIN] uint2 {lowb, span} of the range check (reduce_rows.cuh)
  int total;
};
// nsw ring slots, stage_bytes per row chunk (chunk length of the plan x 8 + 16: a plan with shorter
// chunks leaves room for a deeper ring)
// The data stages are packed by row: warp w owns rows w and w + nw of a group, so only the first
// max_rows - nw warps need stages for two rows.
__host__ __device__ inline Smem2Layout smem2_layout(int nw, int nsw, int stage_bytes, int max_rows) {
  Smem2Layout L;
  int off = 0;
  L.tab = off;
#if defined(ECCKD_TAB_LOGN)
  off += 8 << (ECCKD_TAB_LOGN + ECCKD_TAB_LOGR);
#elif defined(ECCKD_V_TAB256)
  off += TRANS_TAB256R_BYTES;
#else
  off += TRANS_TAB64_BYTES;
#endif
  L.bars = off;
  off += (nw * NSETS + 2 * nsw) * 8;
  off = (off + 127) & ~127;
  L.ring = off;
  off += nsw * RING_SLOT_BYTES;
  L.data = off;
  off += max_rows * NSETS * stage_bytes;
  L.ent = off;
  off += nw * WIN * 32 * 16;
  L.s1e = off;
  off += WIN * 32 * 8;
  L.mmv = off;
  off += nw * RPW * WIN * 16;
  L.mmb = off;
  off += nw * RPW * WIN * 8;
  L.total = (off + 15) & ~15;
  return L;
}

// A rare round end: some lane found a new extreme.  Lane by lane (ballot order), lane 0 merges the
// lane's extremes into the (row, entry) cache and refreshes the range-check words.
template <int NR>
__device__ __forceinline__ void merge_extremes(const Partial (&a)[NR], const int e, const bool used, double2* mmv,
                                               uint2* mmb, const int lane, const unsigned hi_cap = 0x7ff00000u) {
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    unsigned m = __ballot_sync(0xffffffffu, used && a[r].upd != 0);
    while (m) {  // warp-uniform
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int es = __shfl_sync(0xffffffffu, e, src);
      const double mn = __shfl_sync(0xffffffffu, a[r].mn, src);
      const double mx = __shfl_sync(0xffffffffu, a[r].mx, src);
      if (lane == 0) {
        double2 v = mmv[r * WIN + es];
        v.x = (mn < v.x) ? mn : v.x;
        v.y = (mx > v.y) ? mx : v.y;
        mmv[r * WIN + es] = v;
        Partial t;
        t.mn = v.x;
        t.mx = v.y;
        minmax_bounds(t, hi_cap);
        mmb[r * WIN + es] = make_uint2(t.lowb, t.span);
      }
      __syncwarp();
    }
  }
}

template <int FKIND, int WMODE>
__global__ void __launch_bounds__((NW_MAX2 + NPW) * 32, 1) reduce2_kernel(const __grid_constant__ ReduceParams p) {
  static_assert(FKIND == F_LINEAR || FKIND == F_TRANS || FKIND == F_SQRT || FKIND == F_NONE,
                "reduce2_kernel: kinds with per-row weight sums run on reduce_kernel");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nw = p.nw, ng = p.ng;
  const int nsw = p.nsw, stage_bytes = p.stage_bytes;
  const Smem2Layout L = smem2_layout(nw, nsw, stage_bytes, p.max_rows);
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned full_d = sbase + L.bars;         // mbarriers [nw][NSETS], 8 bytes each
  const unsigned full_w = full_d + nw * NSETS * 8; // [NSW]
  const unsigned empty_w = full_w + nsw * 8;       // [nsw]
  unsigned char* ring = smem_raw + L.ring;
  constexpr bool USES_TAU = (FKIND != F_NONE);

  if (threadIdx.x == 0) {
    for (int i = 0; i < nw * NSETS; ++i) mbar_init(full_d + i * 8, 1);
    for (int i = 0; i < nsw; ++i) {
      mbar_init(full_w + i * 8, NPW);
      mbar_init(empty_w + i * 8, nw);
    }
    fence_mbar_init();
  }
#if defined(ECCKD_TAB_LOGN)
  if (FKIND == F_TRANS) trans_tabT_init<ECCKD_TAB_LOGN, ECCKD_TAB_LOGR>(reinterpret_cast<double*>(smem_raw + L.tab), p.exp2tab);
#elif defined(ECCKD_V_TAB256)
  if (FKIND == F_TRANS) trans_tab256r_init(reinterpret_cast<double*>(smem_raw + L.tab));
#else
  if (FKIND == F_TRANS) trans_tab64_init(reinterpret_cast<double*>(smem_raw + L.tab));
#endif
  __syncthreads();
  const unsigned char* tab = smem_raw + L.tab;
#if defined(ECCKD_V_NOCLAMP) && defined(ECCKD_TAB_LOGN)
  const unsigned hi_cap2 = FKIND == F_TRANS ? (unsigned)p.tct.hi_max : 0x7ff00000u;  // see rows_slots
#else
  const unsigned hi_cap2 = 0x7ff00000u;
#endif

  // this CTA's contiguous piece of the (window, group, chunk) space
  const int nchunks = (int)p.nchunks;
  const int ch = p.ch;
  const int64_t T = (int64_t)p.ngroups * p.nchunks;
  const int64_t lo = (T * blockIdx.x) / p.ncta;
  const int64_t hi = (T * (blockIdx.x + 1)) / p.ncta;
  const int niter = (int)(hi - lo);
  if (niter == 0) return;
  const Cursor first = cursor_at(lo, p.ngroups, p.cw, nchunks);
  const int last_n = (int)(p.nwav - (int64_t)(nchunks - 1) * ch);  // length of the last chunk of a row

  if (warp >= nw) {
#ifdef ECCKD_PLANCK_TAB
    produce_rounds<WMODE, (WMODE == W_PLANCK && FKIND == F_TRANS)>(p, first, niter, warp - nw, lane, ring, sbase + L.ring, full_w,
                                                                   empty_w, nsw, smem_raw + L.tab);
#else
    produce_rounds<WMODE>(p, first, niter, warp - nw, lane, ring, sbase + L.ring, full_w, empty_w, nsw);
#endif
    return;
  }

  // ===================== consumers: warp = RPW rows of the group =====================
  const TransLane tl = {0, 0};  // (only the logarithmic kinds of reduce_kernel use the lane table)
  const int n2 = p.max_rows - nw;               // warps 0 .. n2-1 have two rows, the others one
  const int rpw_w = (warp < n2) ? RPW : 1;      // rows this warp has stages for
  const int stage_first = (warp + min(warp, n2)) * NSETS;
  unsigned char* data = smem_raw + L.data + (size_t)stage_first * stage_bytes;
  const unsigned my_full = full_d + warp * NSETS * 8;
  const unsigned my_data = sbase + L.data + stage_first * stage_bytes;
  double2* ent = reinterpret_cast<double2*>(smem_raw + L.ent) + (size_t)warp * WIN * 32;
  double* s1e = reinterpret_cast<double*>(smem_raw + L.s1e);
  double2* mmv = reinterpret_cast<double2*>(smem_raw + L.mmv) + warp * RPW * WIN;
  uint2* mmb = reinterpret_cast<uint2*>(smem_raw + L.mmb) + warp * RPW * WIN;
#if defined(ECCKD_TAB_LOGN)
  const unsigned lane_off = sbase + L.tab + ((unsigned)(lane & ((1 << ECCKD_TAB_LOGR) - 1)) << 3);  // shared-window address
#elif defined(ECCKD_V_TAB256)
  const unsigned lane_off = (unsigned)(lane & 7) << 3;
#else
  const unsigned lane_off = (unsigned)(lane & 15) << 3;
#endif
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const uint64_t pol = policy_evict_first();
  // the weight sums do not depend on the row: the last consumer warp (the one with the fewest rows)
  // accumulates them once per group; they end up in the s1 field of its row 0
  // (with a floating row the last two warps sit on the producer-free schedulers, which carry 5.5 rows
  // each: the duty then goes to a warp on a scheduler with a producer and 4 rows)
  const bool s1_duty = warp == p.s1_warp;

  int gbase = -1;            // first g point of the current window (-1: none loaded)
  double* pout = nullptr;    // this warp's rows of the current (window, group) piece: [RPW][ng][4]
  int32_t* poutz = nullptr;

  // the warp's rows of a new (window, group) piece start as {0, 0, +inf, -inf}
  auto begin_piece = [&](int vg) {
    const size_t slot = ((size_t)blockIdx.x * p.maxg + (size_t)vg) * (nw * RPW) + warp * RPW;
    pout = p.partial + slot * ng * 4;
    poutz = p.partial_z + slot * ng;
    for (int i = lane; i < RPW * ng; i += 32) {
      double2* q = reinterpret_cast<double2*>(pout + (size_t)i * 4);
      q[0] = make_double2(0.0, 0.0);
      q[1] = make_double2(inf, -inf);
      poutz[i] = 0;
    }
    __syncwarp();
    gbase = -1;
  };
  // add the window's entries up over the lanes (fixed butterfly) and fold them, with the extremes,
  // into the piece: lane e writes g point gbase + e
  auto window_store = [&]() {
    if (gbase < 0) return;  // warp-uniform
    double t0 = 0.0, t1 = 0.0, ts = 0.0;
#pragma unroll
    for (int e = 0; e < WIN; ++e) {
      double2 v = ent[e * 32 + lane];
      double s = s1_duty ? s1e[e * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        if (s1_duty) s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      if (lane == e) {
        t0 = v.x;
        t1 = v.y;
        ts = s;
      }
    }
    const int g = gbase + lane;
    if (lane < WIN && g < ng) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        double* q = pout + ((size_t)r * ng + g) * 4;
        const double2 mm = mmv[r * WIN + lane];
        q[0] += (r == 0) ? t0 : t1;
        q[2] = (mm.x < q[2]) ? mm.x : q[2];
        q[3] = (mm.y > q[3]) ? mm.y : q[3];
      }
      if (s1_duty) pout[(size_t)g * 4 + 1] += ts;
    }
    __syncwarp();
  };
  // open the window [nb, nb + WIN): zero entries, extremes and range-check words from the piece
  auto window_load = [&](int nb) {
    gbase = nb;
#pragma unroll
    for (int e = 0; e < WIN; ++e) ent[e * 32 + lane] = make_double2(0.0, 0.0);
    if (s1_duty) {
#pragma unroll
      for (int e = 0; e < WIN; ++e) s1e[e * 32 + lane] = 0.0;
    }
    if (lane < WIN) {
      const int g = nb + lane;
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        Partial t;
        t.mn = inf;
        t.mx = -inf;
        if (g < ng) {
          const double* q = pout + ((size_t)r * ng + g) * 4;
          t.mn = q[2];
          t.mx = q[3];
        }
        minmax_bounds(t, hi_cap2);
        mmv[r * WIN + lane] = make_double2(t.mn, t.mx);
        mmb[r * WIN + lane] = make_uint2(t.lowb, t.span);
      }
    }
    __syncwarp();
  };

  // issue cursor (runs NSETS chunk visits ahead of the compute cursor).  Every row chunk is a bulk
  // copy: a row that is only 8-byte aligned is copied from the aligned address below it and read at
  // a byte shift of 8, an odd tail is copied up to the next even length (the bytes read in excess
  // stay inside the 16-byte granules of the caller's rows).
  Cursor qi = first;
  const unsigned char* src_i[RPW];  // next chunk of the warp's rows (nullptr: no such row in the group)
  // The floating row.  A group of 2 nw - 1 rows leaves one warp with a single row, and with it one
  // scheduler (warp w runs on scheduler w % 4) with a row less than its neighbour: 6 / 5 rows on the two
  // schedulers without producer warps.  The group's last row therefore changes hands every chunk visit:
  // warp nw-2 takes it on the visits of stage set 0, warp nw-1 on those of set 1 (visit k uses set k & 1,
  // so each of the two always finds its own copies in the set it issued them into; warp nw-1 borrows
  // warp nw-2's second-row stage of set 1, which that warp does not use).  Both schedulers then carry
  // 5.5 rows on average and the ring's slack absorbs the alternation.  The row's two partial results
  // (one per warp) are added up by combine_kernel.
  bool fl_i = false;  // the issue cursor's group has a floating row
  auto row1_of = [&](const GroupDesc* gd, bool& fl) -> int {  // index of this warp's second row, or -1
    fl = p.float_row && gd->nrows == 2 * nw - 1 && warp >= nw - 2;
    if (fl) return gd->nrows - 1;
    return (warp + nw < gd->nrows) ? warp + nw : -1;
  };
  auto row1_on = [&](const bool fl, const int set) { return !fl || set == (warp == nw - 2 ? 0 : 1); };
  auto load_issue_group = [&]() {
    const GroupDesc* gd = p.groups + qi.l;
    const int i1 = row1_of(gd, fl_i);
    src_i[0] = (USES_TAU && warp < gd->nrows) ? reinterpret_cast<const unsigned char*>(gd->od[warp] + (int64_t)qi.c * ch) : nullptr;
    src_i[1] = (USES_TAU && i1 >= 0) ? reinterpret_cast<const unsigned char*>(gd->od[i1] + (int64_t)qi.c * ch) : nullptr;
  };
  // byte offset of the stage of (set, row r of this warp) from the warp's first stage
  auto stage_off = [&](const int set, const int r) -> int {
    if (r == 1 && rpw_w == 1)  // borrowed: warp nw-2's (set, row 1) stage, NSETS*RPW stages before ours
      return (set * RPW + 1 - NSETS * RPW) * stage_bytes;
    return (set * rpw_w + r) * stage_bytes;
  };
  auto issue = [&](int set) {
    if (src_i[0] != nullptr) {
      const unsigned n_bytes = (unsigned)((qi.c == nchunks - 1) ? last_n : ch) * 8u;
      const bool on1 = src_i[1] != nullptr && row1_on(fl_i, set);
      unsigned shift[RPW], bytes[RPW], total = 0;
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const bool on = r == 0 || on1;
        shift[r] = (unsigned)(reinterpret_cast<uintptr_t>(src_i[r]) & 8u);
        bytes[r] = on ? ((shift[r] + n_bytes + 15u) & ~15u) : 0u;
        total += bytes[r];
      }
      if (lane == 0) {
        const unsigned bar = my_full + set * 8;  // the rows of a set complete on ONE mbarrier
        mbar_arrive_expect_tx(bar, total);
#pragma unroll
        for (int r = 0; r < RPW; ++r)
          if (bytes[r]) bulk_g2s_hint(my_data + stage_off(set, r), src_i[r] - shift[r], bytes[r], bar, pol);
      }
    }
    if (++qi.c < qi.c_end) {
#pragma unroll
      for (int r = 0; r < RPW; ++r)
        if (src_i[r]) src_i[r] += (size_t)ch * 8;
    } else {
      --qi.c;
      cursor_next(qi, p.ngroups, p.cw, nchunks);
      load_issue_group();
    }
  };
  load_issue_group();
  for (int j = 0; j < NSETS && j < niter; ++j) issue(j);

  Cursor q = first;
  int ws = 0, vg = 0;
  unsigned wpar = 0;
  unsigned phase_bits = 0;  // bit `set`: parity to wait for on that stage set's mbarrier
  const double* od_c[RPW];
  bool have[RPW];     // the warp's rows in the current group
  bool fl_c = false;  // ... the second of which is the floating row
  bool need_group = true;

  // the slots of one round for NR rows: range-check words and extremes from the cache, slot loop,
  // sums into the lane's entry, extremes back into the cache if one changed
  auto round_rows = [&](auto nr_tag, const unsigned char* stage0, const uint16_t* offp, const double* wp, const int e,
                        const bool used, const int row0, const int row_stride) {
    constexpr int NR = decltype(nr_tag)::value;
    Partial a[NR];
    bool il[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const double2 mm = mmv[(row0 + r) * WIN + e];
      const uint2 b = mmb[(row0 + r) * WIN + e];
      a[r].s0 = 0.0;
      a[r].s1 = 0.0;
      a[r].zc = 0;
      a[r].upd = 0;
      a[r].mn = mm.x;
      a[r].mx = mm.y;
      a[r].lowb = used ? b.x : 0u;          // unused lanes accept everything: no exact path, no update
      a[r].span = used ? b.y : 0xFFFFFFFFu;
      il[r] = false;
    }
    rows_slots<FKIND, NR, true>(a, stage0, offp, wp, SLOTS, p.tc, tl, il, p.tc256, tab, &p.tc64, lane_off, row_stride, &p.tct);
    if (used) {
      double2 v = ent[e * 32 + lane];
      if (NR > 1) {
        v.x += a[0].s0;
        v.y += a[NR - 1].s0;
      } else if (row0 == 0) {
        v.x += a[0].s0;
      } else {
        v.y += a[0].s0;
      }
      ent[e * 32 + lane] = v;
    }
    int u = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) u |= a[r].upd;
    if (__any_sync(0xffffffffu, used && u != 0)) merge_extremes<NR>(a, e, used, mmv + row0 * WIN, mmb + row0 * WIN, lane, hi_cap2);
  };

  // one chunk visit on stage set `set`
  auto visit = [&](const int set, const int k) {
    if (need_group) {
      const GroupDesc* gd = p.groups + q.l;
      const int i1 = row1_of(gd, fl_c);
      have[0] = USES_TAU && warp < gd->nrows;
      have[1] = USES_TAU && i1 >= 0;
      od_c[0] = have[0] ? gd->od[warp] : nullptr;
      od_c[1] = have[1] ? gd->od[i1] : nullptr;
      begin_piece(vg);
      need_group = false;
    }
    bool active[RPW];
    active[0] = have[0];
    active[1] = have[1] && row1_on(fl_c, set);
    if (active[0]) {
      mbar_wait(my_full + set * 8, (phase_bits >> set) & 1u);
      phase_bits ^= 1u << set;
    }
    // byte shift of each row's chunk inside its stage (chunk lengths are even: a row's parity is fixed)
    unsigned shift_c[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) shift_c[r] = (unsigned)(reinterpret_cast<uintptr_t>(od_c[r]) & 8u);
    const int row_stride = stage_off(set, 1) - stage_off(set, 0);
    bool last_round = false;
    while (!last_round) {
      mbar_wait(full_w + ws * 8, wpar);
      const unsigned char* slot = ring + ws * RING_SLOT_BYTES;
      const PlanRound* pr = reinterpret_cast<const PlanRound*>(slot + WPERM_BYTES);
      const double* wp = reinterpret_cast<const double*>(slot) + lane;
      const uint16_t* offp = &pr->off[0][lane];
      const unsigned meta = pr->meta[lane];
      const int wb = (int)pr->win;
      last_round = (pr->flags & 2u) != 0;
      if (wb != gbase) {  // warp-uniform: the window of g points moves (or the piece is new)
        window_store();
        window_load(wb);
      }
      const unsigned g = meta & 0xFFFu;
      const bool used = g != NO_G;
      const int e = used ? min(max((int)g - gbase, 0), WIN - 1) : 0;  // the plan keeps g inside the window
      if (s1_duty) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < SLOTS; ++kk) s += wp[kk * 32];
        if (used) s1e[e * 32 + lane] += s;
      }
      if (USES_TAU && active[0]) {
        const unsigned char* stage0 = data + stage_off(set, 0) + shift_c[0];
        // (both rows of a pair have the same alignment in every caller of this library: spectra are
        // [layer][nwav] arrays with one stride; a pair with different parities takes the single-row form twice)
        if (active[RPW - 1] && shift_c[RPW - 1] == shift_c[0]) {
          round_rows(std::integral_constant<int, RPW>{}, stage0, offp, wp, e, used, 0, row_stride);
        } else {
          round_rows(std::integral_constant<int, 1>{}, stage0, offp, wp, e, used, 0, 0);
          if (active[RPW - 1])
            round_rows(std::integral_constant<int, 1>{}, data + stage_off(set, 1) + shift_c[RPW - 1], offp, wp, e, used, 1, 0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_w + ws * 8);
      if (++ws == nsw) {
        ws = 0;
        wpar ^= 1u;
      }
    }
    // re-arm the set for the visit after next (the stage was only read through the generic proxy)
    if (k + NSETS < niter) issue(set);
    if (cursor_next(q, p.ngroups, p.cw, nchunks)) {
      window_store();
      ++vg;
      need_group = true;
    }
  };

#pragma unroll 1
  for (int k = 0; k < niter; ++k) visit(k & (NSETS - 1), k);
  if (!need_group) window_store();
}

// ------------------------------------------------------------------ combine + finalize
// Add the CTA pieces of every (row, g) in fixed (window, CTA) order (deterministic).
__global__ void combine_kernel(const CombineParams p) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.ngroups * p.nw * p.rpw * p.ng;
  if (tid >= total) return;
  const int g = (int)(tid % p.ng);
  const int i = (int)((tid / p.ng) % (p.nw * p.rpw));  // row of the group
  const int grp = (int)(tid / ((int64_t)p.ng * p.nw * p.rpw));
  const GroupDesc* gd = p.groups + grp;
  if (i >= gd->nrows) return;
  const int piece = (i % p.nw) * p.rpw + (i / p.nw);                 // (warp, row of the warp)
  const bool floating = p.float_row && gd->nrows == 2 * p.nw - 1 && i == gd->nrows - 1;
  const int piece_s1 = p.shared_s1 ? p.s1_warp * p.rpw : piece;      // weight sums: row 0 of the warp with that duty
  const int nrw = p.nw * p.rpw;
  const int nchunks = (int)p.nchunks;
  const int64_t T = (int64_t)p.ngroups * p.nchunks;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double s0 = 0.0, s1 = 0.0, mn = inf, mx = -inf;
  long long zc = 0;
  for (int v = 0; v < p.nwin; ++v) {
    const int c_begin = v * p.cw, c_end = min(nchunks, c_begin + p.cw);
    const int n_v = c_end - c_begin;
    if (n_v <= 0) break;
    const int64_t glo = (int64_t)p.ngroups * c_begin + (int64_t)grp * n_v, ghi = glo + n_v;
    const int vg = v * p.ngroups + grp;
    int b = (int)((glo * p.ncta) / T);
    if (b > 0) --b;
    for (; b < p.ncta; ++b) {
      const int64_t lo = (T * b) / p.ncta, hi = (T * (b + 1)) / p.ncta;
      if (lo >= ghi) break;
      if (hi <= glo || hi <= lo) continue;
      // (window, group) index of the CTA's first item
      const int64_t per_win = (int64_t)p.ngroups * p.cw;
      const int fv = (int)(lo / per_win);
      const int fn = min(nchunks, fv * p.cw + p.cw) - fv * p.cw;
      const int fvg = fv * p.ngroups + (int)((lo - (int64_t)fv * per_win) / fn);
      const size_t base = ((size_t)b * p.maxg + (size_t)(vg - fvg)) * nrw;
      const double* a = p.partial + ((base + piece) * p.ng + g) * 4;
      s0 += a[0];
      s1 += p.partial[((base + piece_s1) * p.ng + g) * 4 + 1];
      mn = fmin(mn, a[2]);
      mx = fmax(mx, a[3]);
      zc += p.partial_z[(base + piece) * p.ng + g];
      if (floating) {  // the floating row's other half: row 1 of the last warp (reduce2_kernel)
        const double* a2 = p.partial + ((base + (size_t)(p.nw - 1) * p.rpw + 1) * p.ng + g) * 4;
        s0 += a2[0];
        mn = fmin(mn, a2[2]);
        mx = fmax(mx, a2[3]);
      }
    }
  }
  const int row = gd->row[i];
  double* so = p.sums + ((size_t)row * p.ng + g) * NSUM;
  so[0] = s0;
  so[1] = s1;
  so[2] = (double)zc;
  double* mo = p.minmax + ((size_t)row * p.ng + g) * NMINMAX;
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
  } else if (s0 != s0) {
    // a NaN optical depth poisoned the sum (reduce_rows.cuh): keep it visible instead of clamping it away
    wflag |= ECCKD_WARN_NONFINITE;
    fit = s0;
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
    // (the product and the sum as one fused multiply-add, spelled out: merge_spectra_kernel must give the same bits)
    if (mode == 2) *d = fma(v, profile[l], *d);
    else if (mode == 1) *d = fma(v, scaling, *d);
    else *d += v;
  }
}

// gpoint_fraction numerators and per-g totals (create_look_up_table.cpp:537-548)
// gpoint_fraction / ssi_intervals (create_look_up_table.cpp:512-561): histograms of a per-wavenumber
// quantity over (g point, wavenumber bin) cells.  The sums are accumulated EXACTLY, as 128-bit
// fixed-point integers whose unit 2^lsb is derived from the largest magnitude in the data: integer
// addition is associative, so the result does not depend on the order in which threads, warps or
// blocks arrive (bitwise reproducible, no floating-point atomics), and it is rounded to FP64 once.
// Values more than 2^33 times smaller than the largest lose their bits below 2^lsb (an absolute
// error below max * 2^-85 each).
struct Fix128 {
  unsigned long long lo;
  long long hi;
};
__device__ __forceinline__ int hist_bin(double nu, int nbin, double bin0, double dwav) {
  // bin i covers (bin0 + i*dwav, bin0 + (i+1)*dwav]
  int i = (int)ceil((nu - bin0) / dwav) - 1;
  // repair rounding of the division against the exact edge tests of :544-545
  if (i >= 0 && i < nbin && !(nu > bin0 + dwav * i)) --i;
  if (i + 1 >= 0 && i + 1 < nbin && nu > bin0 + dwav * (i + 1)) ++i;
  return (i >= 0 && i < nbin && nu > bin0 + dwav * i && nu <= bin0 + dwav * (i + 1)) ? i : -1;
}
// largest magnitude (as the bit pattern of |v|, which orders like the value) of the values that count
__global__ void hist_max_kernel(int ng, int64_t nwav, const int32_t* __restrict__ g_point,
                                const double* __restrict__ val, unsigned long long* __restrict__ maxbits) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long b = 0ull;
  if (w < nwav && (!g_point || (g_point[w] >= 0 && g_point[w] < ng))) {
    const double v = fabs(val[w]);
    if (v == v && v <= 1.7976931348623157e308) b = (unsigned long long)__double_as_longlong(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t > b ? t : b;
  }
  if ((threadIdx.x & 31) == 0 && b) atomicMax(maxbits, b);
}
// v as a multiple of 2^lsb, two's complement
__device__ __forceinline__ Fix128 to_fix128(double v, int lsb) {
  Fix128 r = {0ull, 0ll};
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const int be = (int)((bits >> 52) & 0x7ffu);
  if (be == 0x7ff) return r;  // non-finite values are reported by the caller, not summed
  const unsigned long long m = (bits & 0xfffffffffffffull) | (be ? 0x10000000000000ull : 0ull);
  const int sh = (be ? be - 1075 : -1074) - lsb;  // v = m * 2^(lsb + sh)
  unsigned long long lo = 0ull, hi = 0ull;
  if (sh >= 64) {
    hi = sh < 128 ? m << (sh - 64) : 0ull;
  } else if (sh > 0) {
    lo = m << sh;
    hi = m >> (64 - sh);
  } else if (sh > -64) {
    lo = m >> (-sh);
  }
  if (bits >> 63) {  // negate
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  r.lo = lo;
  r.hi = (long long)hi;
  return r;
}
__device__ __forceinline__ void fix128_add(Fix128& a, const Fix128 b) {
  const unsigned long long lo = a.lo + b.lo;
  a.hi += b.hi + (lo < a.lo ? 1ll : 0ll);
  a.lo = lo;
}
// The lanes of a warp that hit the same cell add up first (lane order; any order gives the same
// integer), then one lane adds the group's sum to the cell: low word, then high word plus the carry
// out of the low word - the two atomics of different threads may interleave freely, the final
// 128-bit value is the exact sum either way.
__device__ __forceinline__ void fix128_warp_accumulate(Fix128* __restrict__ cells, const long long key, Fix128 x) {
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  unsigned rest = peers & (peers - 1);  // the group without its leader
  unsigned any = __ballot_sync(0xffffffffu, rest != 0u && lane == leader);
  while (any) {  // warp-uniform: some group still has members to fold
    const int src = rest ? __ffs(rest) - 1 : lane;
    Fix128 y;
    y.lo = __shfl_sync(0xffffffffu, x.lo, src);
    y.hi = __shfl_sync(0xffffffffu, x.hi, src);
    if (rest && lane == leader) fix128_add(x, y);
    rest &= rest - 1;
    any = __ballot_sync(0xffffffffu, rest != 0u && lane == leader);
  }
  if (lane == leader && key >= 0 && (x.lo | (unsigned long long)x.hi)) {
    unsigned long long* c = reinterpret_cast<unsigned long long*>(cells + key);
    const unsigned long long old = atomicAdd(c, x.lo);
    const unsigned long long carry = (old + x.lo < old) ? 1ull : 0ull;
    const unsigned long long h = (unsigned long long)x.hi + carry;
    if (h) atomicAdd(c + 1, h);
  }
}
__device__ __forceinline__ int hist_lsb(const unsigned long long maxbits) {
  // the largest value has its leading bit at 2^e; 2^40 of them still fit below bit 126
  const int be = (int)((maxbits >> 52) & 0x7ffu);
  const int e = be ? be - 1023 : -1022;
  return e + 41 - 126;
}
// cells: [ng*nbin] (g, bin) sums, then [ng] totals
__global__ void gfrac_hist_kernel(int ng, int64_t nwav, const int32_t* __restrict__ g_point,
                                  const double* __restrict__ wn, const double* __restrict__ dwn, int nbin,
                                  double bin0, double dwav, const unsigned long long* __restrict__ maxbits,
                                  Fix128* __restrict__ cells) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lsb = hist_lsb(*maxbits);
  long long kcell = -1, ktot = -1;
  Fix128 x = {0ull, 0ll};
  if (w < nwav) {
    const int g = g_point[w];
    if (g >= 0 && g < ng) {
      x = to_fix128(dwn[w], lsb);
      ktot = (long long)ng * nbin + g;
      const int i = hist_bin(wn[w], nbin, bin0, dwav);
      if (i >= 0) kcell = (long long)g * nbin + i;
    }
  }
  fix128_warp_accumulate(cells, ktot, x);
  fix128_warp_accumulate(cells, kcell, x);
}
__device__ __forceinline__ double fix128_to_double(Fix128 a, int lsb) {
  const bool neg = a.hi < 0;
  unsigned long long lo = a.lo, hi = (unsigned long long)a.hi;
  if (neg) {
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  // hi * 2^64 + lo, rounded once: fold the low word's top bits into a 64-bit significand with a sticky bit
  double r;
  if (hi == 0ull) {
    r = scalbn(__ull2double_rn(lo), lsb);
  } else {
    const int z = __clzll((long long)hi);
    unsigned long long top = z ? (hi << z) | (lo >> (64 - z)) : hi;
    const unsigned long long below = z ? (lo << z) : lo;
    if (below) top |= 1ull;  // sticky: the discarded bits only matter for ties
    r = scalbn(__ull2double_rn(top), lsb + 64 - z);
  }
  return neg ? -r : r;
}
__global__ void gfrac_norm_kernel(int ng, int nbin, const Fix128* __restrict__ cells,
                                  const unsigned long long* __restrict__ maxbits, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng * nbin) return;
  const int lsb = hist_lsb(*maxbits);
  out[i] = fix128_to_double(cells[i], lsb) / fix128_to_double(cells[(size_t)ng * nbin + i / nbin], lsb);
}

__global__ void ssi_hist_kernel(int64_t nwav, const double* __restrict__ wn, const double* __restrict__ ssi,
                                int nbin, double bin0, double dwav, const unsigned long long* __restrict__ maxbits,
                                Fix128* __restrict__ cells) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lsb = hist_lsb(*maxbits);
  long long key = -1;
  Fix128 x = {0ull, 0ll};
  if (w < nwav) {
    const int i = hist_bin(wn[w], nbin, bin0, dwav);
    if (i >= 0) {
      key = i;
      x = to_fix128(ssi[w], lsb);
    }
  }
  fix128_warp_accumulate(cells, key, x);
}
__global__ void fix128_out_kernel(int n, const Fix128* __restrict__ cells, const unsigned long long* __restrict__ maxbits,
                                  double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fix128_to_double(cells[i], hist_lsb(*maxbits));
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
template <int FKIND, int WMODE, int GEN>
static cudaError_t launch_reduce_t(ecckd_ctx* ctx, ReduceParams& p, int threads, size_t smem, int* ncta_out,
                                   bool dry_run) {
  void (*kernel)(const ReduceParams) = nullptr;
  if constexpr (GEN == 2) kernel = reduce2_kernel<FKIND, WMODE>;
  else kernel = reduce_kernel<FKIND, WMODE>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (dry_run) {
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    *ncta_out = std::max(1, nb) * ctx->num_sms;
    return cudaSuccess;
  }
  kernel<<<p.ncta, threads, smem, ctx->stream>>>(p);
  return cudaGetLastError();
}

template <int FKIND, int GEN>
static cudaError_t launch_reduce_w(int wmode, ecckd_ctx* ctx, ReduceParams& p, int threads, size_t smem, int* ncta,
                                   bool dry) {
  if (wmode == W_MATRIX) return launch_reduce_t<FKIND, W_MATRIX, GEN>(ctx, p, threads, smem, ncta, dry);
  if (wmode == W_VECTOR) return launch_reduce_t<FKIND, W_VECTOR, GEN>(ctx, p, threads, smem, ncta, dry);
  return launch_reduce_t<FKIND, W_PLANCK, GEN>(ctx, p, threads, smem, ncta, dry);
}

// v2: the windowed kernel (reduce2_kernel); only for the kinds reduce2_supports() names
static bool reduce2_supports(int fkind) {
  return fkind == F_LINEAR || fkind == F_TRANS || fkind == F_SQRT || fkind == F_NONE;
}
static cudaError_t launch_reduce(int fkind, int wmode, int gen, ecckd_ctx* ctx, ReduceParams& p, int threads,
                                 size_t smem, int* ncta, bool dry) {
  if (gen == 2) {
    switch (fkind) {
      case F_LINEAR: return launch_reduce_w<F_LINEAR, 2>(wmode, ctx, p, threads, smem, ncta, dry);
      case F_TRANS: return launch_reduce_w<F_TRANS, 2>(wmode, ctx, p, threads, smem, ncta, dry);
      case F_SQRT: return launch_reduce_w<F_SQRT, 2>(wmode, ctx, p, threads, smem, ncta, dry);
      default: return launch_reduce_w<F_NONE, 2>(wmode, ctx, p, threads, smem, ncta, dry);
    }
  }
  switch (fkind) {
    case F_LINEAR: return launch_reduce_w<F_LINEAR, 1>(wmode, ctx, p, threads, smem, ncta, dry);
    case F_TRANS: return launch_reduce_w<F_TRANS, 1>(wmode, ctx, p, threads, smem, ncta, dry);
    case F_SQRT: return launch_reduce_w<F_SQRT, 1>(wmode, ctx, p, threads, smem, ncta, dry);
    case F_LOG: return launch_reduce_w<F_LOG, 1>(wmode, ctx, p, threads, smem, ncta, dry);
    case F_HYBRID: return launch_reduce_w<F_HYBRID, 1>(wmode, ctx, p, threads, smem, ncta, dry);
    default: return launch_reduce_w<F_NONE, 1>(wmode, ctx, p, threads, smem, ncta, dry);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 2^(j/N), j = 0 .. N-1, evaluated in extended precision on the host and rounded to FP64 once; one
// device copy per context.
static int exp2tab_get(ecckd_ctx* ctx, int logn, const double** out) {
  if (!ctx->d_exp2tab || ctx->exp2tab_logn != logn) {
    const int n = 1 << logn;
    std::vector<double> h((size_t)n);
    for (int j = 0; j < n; ++j) h[j] = (double)exp2l((long double)j / (long double)n);
    if (ctx->d_exp2tab) {
      ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      ECCKD_CUDA_OK(cudaFree(ctx->d_exp2tab));
      ctx->d_exp2tab = nullptr;
    }
    ECCKD_CUDA_OK(cudaMalloc(&ctx->d_exp2tab, (size_t)n * sizeof(double)));
    ECCKD_CUDA_OK(cudaMemcpy(ctx->d_exp2tab, h.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    ctx->exp2tab_logn = logn;
  }
  *out = ctx->d_exp2tab;
  return ECCKD_OK;
}

// Streaming stage: reduce `rows` (any mix of spectra/layers sharing one plan and
// one weight definition) to per-(row, g) partial results in d_sums / d_minmax
// (layouts in spectral_reduce.cuh), rows[i].row giving each row's index there.
// Consecutive rows with the same weight row are grouped into one CTA-wide group.
static int reduce_partial(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, const std::vector<RowDesc>& rows, int fkind,
                          double kd, int wmode, const double* d_weight, int64_t w_stride,
                          const double* h_temperature, int n_temperature, const double* d_wn, const double* d_dwn,
                          double* d_sums, double* d_minmax) {
  const int nrows = (int)rows.size();
  if (nrows == 0) return ECCKD_OK;
  const int ng = plan->ng;
  const int64_t nwav = plan->nwav;

  // consumer warps per CTA: RPW rows each; as many as the longest run of rows sharing a weight row
  // needs, capped by NW_MAX and shared memory
  int longest = 1;
  for (int i = 0, run = 0; i < nrows; ++i) {
    run = (i > 0 && rows[i].wrow == rows[i - 1].wrow) ? run + 1 : 1;
    longest = std::max(longest, run);
  }
  const size_t smem_max = ctx->smem_optin ? ctx->smem_optin : (size_t)227 * 1024;
  // Kernel generation: windowed plans and the kinds without per-row weight sums run on
  // reduce2_kernel, everything else on reduce_kernel (shuffle fold, per-row weight sums).
  // ECCKD_REDUCE_GEN=1 forces reduce_kernel (comparisons).
  int gen = (plan->windowed && reduce2_supports(fkind)) ? 2 : 1;
  if (const char* env = getenv("ECCKD_REDUCE_GEN")) {
    const int want = atoi(env);
    if (want >= 1 && want < gen) gen = want;
  }
  if (gen == 1 && plan->balanced) {
    if (!plan->legacy) ECCKD_TRY(::plan_create_impl(ctx, plan->d_g_copy, plan->nwav, plan->ng, true, &plan->legacy));
    plan = plan->legacy;
  }
  const int rpw = RPW;
  const int stage2 = plan->ch * 8 + 16;  // reduce2_kernel: row-chunk stage incl. the 8-byte shift of unaligned rows
  // rows of the largest group when every run of rows sharing a weight row is cut into
  // ceil(n / (nw*RPW)) pieces of nearly equal size
  std::vector<int> runs;
  for (int i = 0; i < nrows;) {
    int j = i;
    while (j < nrows && rows[j].wrow == rows[i].wrow) ++j;
    runs.push_back(j - i);
    i = j;
  }
  auto max_rows_of = [&](int n) {
    int m = 1;
    for (int len : runs) {
      const int pieces = (len + n * rpw - 1) / (n * rpw);
      m = std::max(m, (len + pieces - 1) / pieces);
    }
    return m;
  };
  auto smem_of = [&](int n) {
    return gen == 2 ? (size_t)smem2_layout(n, NSW, stage2, max_rows_of(n)).total : (size_t)smem_layout(n, ng).total;
  };
  // reduce2_kernel: one row per warp while the rows of a group fit NW_MAX2 warps, otherwise NW_MAX2
  // warps of which the first max_rows - nw take two rows (ECCKD_REDUCE_NW overrides, timing only)
  int nw = gen == 2 ? std::min(NW_MAX2, longest) : std::min(NW_MAX, (longest + rpw - 1) / rpw);
  if (gen == 2)
    if (const char* env = getenv("ECCKD_REDUCE_NW")) nw = std::max(1, std::min(NW_MAX2, atoi(env)));
  while (nw > 1 && smem_of(nw) > smem_max) --nw;
  ECCKD_REQUIRE(smem_of(nw) <= smem_max, ECCKD_PARAMETER_ERROR,
                "average_optical_depth: ng = %d needs %zu B of shared memory (limit %zu)", ng, smem_of(nw), smem_max);
  // balance: split a run of n rows into ceil(n/(nw*RPW)) groups of nearly equal size
  const int cap = nw * rpw;
  std::vector<GroupDesc> groups;
  for (int i = 0; i < nrows;) {
    int j = i;
    while (j < nrows && rows[j].wrow == rows[i].wrow) ++j;
    const int n = j - i;
    const int pieces = (n + cap - 1) / cap;
    for (int q = 0; q < pieces; ++q) {
      const int b0 = i + (int)(((int64_t)n * q) / pieces), b1 = i + (int)(((int64_t)n * (q + 1)) / pieces);
      GroupDesc gd;
      memset(&gd, 0, sizeof(gd));
      gd.nrows = b1 - b0;
      gd.wrow = rows[i].wrow;
      for (int r = b0; r < b1; ++r) {
        gd.od[r - b0] = rows[r].od;
        gd.row[r - b0] = rows[r].row;
        gd.is_log[r - b0] = (rows[r].method == M_LOG) ? 1 : 0;
        gd.bulk_ok[r - b0] = aligned16(rows[r].od) ? 1 : 0;
      }
      groups.push_back(gd);
    }
    i = j;
  }
  const int ngroups = (int)groups.size();
  const int threads = (nw + NPW) * 32;
  // reduce2_kernel: as many ring slots as shared memory leaves room for (the ring couples the consumer
  // warps: a slot is refilled when ALL of them have released it, so its depth is their slack)
  int nsw = NSW;
  const int max_rows = max_rows_of(nw);
  if (gen == 2)
    while (nsw < 8 && (size_t)smem2_layout(nw, nsw + 1, stage2, max_rows).total <= smem_max) ++nsw;
  const size_t smem = gen == 2 ? (size_t)smem2_layout(nw, nsw, stage2, max_rows).total : smem_of(nw);

  ReduceParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.ngroups = ngroups;
  rp.nw = nw;
  rp.nchunks = plan->nchunks;
  rp.nwav = nwav;
  rp.ng = ng;
  rp.ch = plan->ch;
  rp.round_off = plan->d_round_off;
  rp.rounds = plan->d_rounds;
  rp.balanced = plan->balanced ? 1 : 0;
  rp.nsw = nsw;
  rp.max_rows = max_rows;
  // the floating row needs warp nw-2 to own second-row stages and warp nw-1 to own none (packed stages)
  rp.float_row = (gen == 2 && nw >= 2 && max_rows == 2 * nw - 1 && !getenv("ECCKD_NO_FLOAT_ROW")) ? 1 : 0;
  rp.stage_bytes = stage2;
  rp.s1_warp = (rp.float_row && nw >= 4 && !getenv("ECCKD_S1_LAST")) ? nw - 3 : nw - 1;
  rp.weight = d_weight;
  rp.w_stride = w_stride;
  rp.weight_bulk_ok = (wmode != W_PLANCK && aligned16(d_weight) && (wmode == W_VECTOR || (w_stride % 2) == 0)) ? 1 : 0;
  rp.wn = d_wn;
  rp.dwn = d_dwn;
  trans_const_init(&rp.tc, kd);
  trans_const256_init(&rp.tc256, kd);
  trans_const64_init(&rp.tc64, kd);
#if defined(ECCKD_TAB_LOGN)
  trans_constT_init(&rp.tct, kd, ECCKD_TAB_LOGN, ECCKD_TAB_QDEG);
  ECCKD_TRY(exp2tab_get(ctx, ECCKD_TAB_LOGN, &rp.exp2tab));
#endif
#ifdef ECCKD_TIMING_KNOBS  // builds for tools/variants.sh only: the knobs skip work, the results are wrong
  if (const char* e = getenv("ECCKD_DEBUG")) rp.debug = atoi(e);
#endif
  // wavenumber windows: the per-wavenumber data a group re-reads (wavenumber grids or weight vector,
  // plan records: ~20 B per wavenumber) should stay in L2 between groups; 24 MB per window does
  int nwin = 1;
  if (ngroups > 1) {
    const double per_wav = (wmode == W_PLANCK ? 16.0 : (wmode == W_VECTOR ? 8.0 : 0.0)) + (double)sizeof(PlanRound) / plan->ch;
    nwin = (int)std::min<int64_t>(64, std::max<int64_t>(1, (int64_t)(per_wav * (double)nwav / 24.0e6 + 0.999)));
    if (const char* e = getenv("ECCKD_NWIN")) nwin = std::max(1, atoi(e));
    nwin = (int)std::min<int64_t>(nwin, plan->nchunks);
  }
  const int cw = (int)((plan->nchunks + nwin - 1) / nwin);
  nwin = (int)((plan->nchunks + cw - 1) / cw);
  rp.nwin = nwin;
  rp.cw = cw;

  int ncta = 0;
  ECCKD_CUDA_OK(launch_reduce(fkind, wmode, gen, ctx, rp, threads, smem, &ncta, true));
  const int64_t T = (int64_t)ngroups * plan->nchunks;
  ncta = (int)std::min<int64_t>(ncta, std::max<int64_t>(1, T));
  // (window, group) pairs a CTA's contiguous piece can touch
  auto vg_of = [&](int64_t item) {
    const int64_t per_win = (int64_t)ngroups * cw;
    const int64_t v = item / per_win;
    const int64_t n_v = std::min<int64_t>(plan->nchunks, v * cw + cw) - v * cw;
    return v * ngroups + (item - v * per_win) / n_v;
  };
  int maxg = 1;
  for (int b = 0; b < ncta; ++b) {
    const int64_t lo = (T * b) / ncta, hi = (T * (b + 1)) / ncta;
    if (hi > lo) maxg = std::max(maxg, (int)(vg_of(hi - 1) - vg_of(lo) + 1));
  }
  rp.ncta = ncta;
  rp.maxg = maxg;

  // device scratch: groups in ws0, CTA pieces in ws1/ws2, temperatures in ws3
  void* d_groups = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 0, (size_t)ngroups * sizeof(GroupDesc), &d_groups));
  ECCKD_TRY(upload_small(ctx, d_groups, groups.data(), (size_t)ngroups * sizeof(GroupDesc)));
  const size_t npieces = (size_t)ncta * maxg * nw * rpw;
  void* d_partial = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, npieces * ng * 4 * sizeof(double), &d_partial));
  void* d_pz = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 2, npieces * ng * sizeof(int32_t), &d_pz));
  void* d_temp = nullptr;
  if (wmode == W_PLANCK) {
    ECCKD_TRY(ws_reserve(ctx, 3, (size_t)n_temperature * sizeof(double), &d_temp));
    // the kernel wants (h/k)/T per weight row (planck_function.cpp:31-33,50)
    std::vector<double> hkt((size_t)n_temperature);
    for (int i = 0; i < n_temperature; ++i) hkt[i] = (6.62606896e-34 / 1.3806504e-23) / h_temperature[i];
    ECCKD_TRY(upload_small(ctx, d_temp, hkt.data(), (size_t)n_temperature * sizeof(double)));
    if (gen == 2 && fkind == F_TRANS) {  // the producers' table form (planck_weight_tab)
      void* d_temp2 = nullptr;
      ECCKD_TRY(ws_reserve(ctx, 6, (size_t)n_temperature * sizeof(double), &d_temp2));
      for (int i = 0; i < n_temperature; ++i)
        hkt[i] = (double)(((long double)6.62606896e-34 / (long double)1.3806504e-23) * (100.0L * 2.99792458e8L) / (long double)h_temperature[i]);
      ECCKD_TRY(upload_small(ctx, d_temp2, hkt.data(), (size_t)n_temperature * sizeof(double)));
      rp.temperature2 = static_cast<const double*>(d_temp2);
      void* d_pref = nullptr;
      ECCKD_TRY(ws_reserve(ctx, 7, (size_t)nwav * sizeof(double), &d_pref));
      planck_prefactor_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(nwav, d_wn, d_dwn, static_cast<double*>(d_pref));
      ECCKD_LAUNCH_CHECK(ctx);
      rp.pref = static_cast<const double*>(d_pref);
    }
  }
  rp.groups = static_cast<const GroupDesc*>(d_groups);
  rp.temperature = static_cast<const double*>(d_temp);
  rp.partial = static_cast<double*>(d_partial);
  rp.partial_z = static_cast<int32_t*>(d_pz);

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->timing) {
    ECCKD_CUDA_OK(cudaEventCreate(&ev0));
    ECCKD_CUDA_OK(cudaEventCreate(&ev1));
    ECCKD_CUDA_OK(cudaEventRecord(ev0, ctx->stream));
  }
  ECCKD_CUDA_OK(launch_reduce(fkind, wmode, gen, ctx, rp, threads, smem, &ncta, false));
  ctx->launches++;
  if (ctx->timing) {
    ECCKD_CUDA_OK(cudaEventRecord(ev1, ctx->stream));
    ctx->reduce_events.emplace_back(ev0, ev1);
  }

  CombineParams cp;
  cp.groups = rp.groups;
  cp.ngroups = ngroups;
  cp.nw = nw;
  cp.rpw = rpw;
  cp.ncta = ncta;
  cp.maxg = maxg;
  cp.nwin = nwin;
  cp.cw = cw;
  cp.nchunks = plan->nchunks;
  cp.ng = ng;
  cp.shared_s1 = !(fkind == F_LOG || fkind == F_HYBRID);
  cp.float_row = rp.float_row;
  cp.s1_warp = rp.s1_warp;
  cp.partial = rp.partial;
  cp.partial_z = rp.partial_z;
  cp.sums = d_sums;
  cp.minmax = d_minmax;
  const int64_t nthreads = (int64_t)ngroups * nw * rpw * ng;
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
    *wmode = (w_stride == 0) ? W_VECTOR : W_MATRIX;  // stride 0 == spread<0>(ssi, nlay)
    ECCKD_REQUIRE(w_stride == 0 || w_stride >= plan->nwav, ECCKD_PARAMETER_ERROR, "weight stride smaller than nwav");
  } else if (weight_mode == ECCKD_WEIGHT_VECTOR) {
    ECCKD_REQUIRE(d_weight != nullptr, ECCKD_PARAMETER_ERROR, "weight vector is NULL");
    *wmode = W_VECTOR;
  } else if (weight_mode == ECCKD_WEIGHT_PLANCK) {
    ECCKD_REQUIRE(h_temperature && d_wn && d_dwn, ECCKD_PARAMETER_ERROR,
                  "Planck weights need temperature, wavenumber and d_wavenumber");
    *wmode = W_PLANCK;
  } else {
    set_error("weight_mode %d not understood", weight_mode);
    return ECCKD_PARAMETER_ERROR;
  }
  return ECCKD_OK;
}

}  // namespace ecckd

using namespace ecckd;

// ------------------------------------------------------------------ C ABI
// Plan buffers come from the stream-ordered allocator: cudaMalloc/cudaFree synchronise the whole
// device, which would serialise a plan built on the compute stream with the host-to-device copies
// the host-buffer entry points run concurrently on their copy stream.
static void plan_free(ecckd_gpoint_plan* p) {
  cudaStream_t st = p->ctx ? p->ctx->stream : nullptr;
  if (p->d_round_off) cudaFreeAsync(p->d_round_off, st);
  if (p->d_rounds) cudaFreeAsync(p->d_rounds, st);
  if (p->d_count) cudaFreeAsync(p->d_count, st);
  if (p->d_scratch) cudaFreeAsync(p->d_scratch, st);
  if (p->d_g_copy) cudaFreeAsync(p->d_g_copy, st);
  if (p->legacy) plan_free(p->legacy);
  delete p;
}

// free_layout: build the plan reduce_kernel needs (no window constraint, lanes of a g point adjacent)
static int plan_create_impl(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng, bool free_layout,
                            ecckd_gpoint_plan** out) {
  ECCKD_REQUIRE(ctx && d_g_point && out, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_create: NULL argument");
  ECCKD_REQUIRE(nwav > 0 && ng > 0 && ng <= MAX_NG, ECCKD_PARAMETER_ERROR,
                "ecckd_gpoint_plan_create: need nwav > 0 and 0 < ng <= %d (got nwav=%lld ng=%d)", MAX_NG,
                (long long)nwav, ng);
  *out = nullptr;
  ecckd_gpoint_plan* p = new ecckd_gpoint_plan();
  p->ctx = ctx;
  p->nwav = nwav;
  p->ng = ng;
  // Chunk length: the plan packs every chunk into rounds of 32 lanes x 16 slots with each g point
  // padded to whole lanes, so the longest chunk whose g points still fit one round is the cheapest;
  // it depends on how many g points meet inside a chunk.  Count the rounds for a few candidates
  // (longest first), with and without the window constraint of reduce2_kernel, and keep the
  // cheapest (a windowed round costs ~0.77 of a free one: no shuffle fold).
  static const int candidates[] = {CH_MAX, 448, 384, 320, 256};
  const int64_t max_chunks = (nwav + 31) / 32;  // the shortest chunk length ECCKD_CH may ask for
  ECCKD_REQUIRE(max_chunks < ((int64_t)1 << 31), ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_create: nwav too large");
  uint32_t* d_nrounds = nullptr;
  unsigned long long* d_nfree = nullptr;
  int32_t* d_rminmax = nullptr;
  unsigned* d_sinfo = nullptr;  // balanced plans: the per-wavenumber words handed from the fill to the colouring kernel
  cudaError_t e = cudaMallocAsync(&p->d_round_off, (size_t)(max_chunks + 1) * sizeof(uint32_t), ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nrounds, (size_t)max_chunks * sizeof(uint32_t), ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nfree, sizeof(unsigned long long), ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&p->d_count, (size_t)ng * sizeof(unsigned long long), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_count, 0, (size_t)ng * sizeof(unsigned long long), ctx->stream);
  uint32_t total = 0;
  std::vector<unsigned long long> cnt(ng);
  int best_ch = 0, best_w = WIN;
  int force_w = -1;  // ECCKD_PLAN_WINDOW=0: free plans only (reduce_kernel), =1: windowed plans only
  if (const char* env = getenv("ECCKD_PLAN_WINDOW")) force_w = atoi(env) ? 1 : 0;
  if (free_layout) force_w = 0;
  bool balance = true;  // ECCKD_PLAN_BALANCE=0: windowed plans keep the adjacent-lane placement (comparisons)
  if (const char* env = getenv("ECCKD_PLAN_BALANCE")) balance = atoi(env) != 0;
  if (const char* env = getenv("ECCKD_CH")) best_ch = std::min(CH_MAX, std::max(32, atoi(env) & ~31));
  // rounds of a plan with chunk length ch and window w (h_total), and without a window (h_free)
  // (the wavenumbers per g point do not depend on ch: they are counted by the first pass only; the round offsets
  // left by the last pass are those of (last_ch, last_w))
  bool counted = false;
  int last_ch = 0, last_w = -1;
  uint32_t last_total = 0;
  auto count_rounds = [&](int ch, int w, bool with_counts, uint32_t* h_total, unsigned long long* h_free) {
    const int64_t nch = (nwav + ch - 1) / ch;
    cudaError_t err = cudaSuccess;
    with_counts = !counted;
    counted = true;
    last_ch = ch;
    last_w = w;
    if (h_free) err = cudaMemsetAsync(d_nfree, 0, sizeof(unsigned long long), ctx->stream);
    if (err != cudaSuccess) return err;
    plan_count_kernel<<<(unsigned)nch, ch, 0, ctx->stream>>>(d_g_point, nwav, ng, w, d_nrounds,
                                                              h_free ? d_nfree : nullptr,
                                                              with_counts ? p->d_count : nullptr);
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(nch, d_nrounds, p->d_round_off);
    ctx->launches += 2;
    err = cudaGetLastError();
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(h_total, p->d_round_off + nch, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess && h_free)
      err = cudaMemcpyAsync(h_free, d_nfree, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err == cudaSuccess) last_total = *h_total;
    return err;
  };
  const double windowed_round_cost = 0.77;
  auto cost_of = [&](uint32_t tot_w, unsigned long long tot_free, int* w_out) {
    const double cw = windowed_round_cost * (double)tot_w, cf = (double)tot_free;
    const bool use_w = force_w < 0 ? (cw <= cf) : (force_w == 1);
    *w_out = use_w ? WIN : 0;
    return use_w ? cw : cf;
  };
  if (e == cudaSuccess && best_ch != 0) {
    uint32_t tot = 0;
    unsigned long long tot_free = 0;
    e = count_rounds(best_ch, WIN, false, &tot, &tot_free);
    if (e == cudaSuccess) cost_of(tot, tot_free, &best_w);
  }
  if (e == cudaSuccess && best_ch == 0) {
    double best_cost = 0.0;
    for (const int& ch : candidates) {
      uint32_t tot = 0;
      unsigned long long tot_free = 0;
      e = count_rounds(ch, WIN, false, &tot, &tot_free);
      if (e != cudaSuccess) break;
      int w = 0;
      const double cost = cost_of(tot, tot_free, &w);  // every round costs the same 32 x 16 slots per row
      if (best_ch == 0 || cost < best_cost * 0.995) {
        best_ch = ch;
        best_cost = cost;
        best_w = w;
      }
      // the next (shorter) candidate needs at least one round per chunk: stop when that cannot win
      const size_t idx = (size_t)(&ch - candidates);
      if (idx + 1 < sizeof(candidates) / sizeof(candidates[0]) &&
          best_cost <= 0.995 * windowed_round_cost * (double)((nwav + candidates[idx + 1] - 1) / candidates[idx + 1]))
        break;
    }
  }
  if (e == cudaSuccess) {
    p->ch = best_ch;
    p->windowed = best_w > 0;
    p->nchunks = (nwav + best_ch - 1) / best_ch;
    if (last_ch == best_ch && last_w == best_w)
      total = last_total;  // the pass that chose this plan left its offsets
    else
      e = count_rounds(best_ch, best_w, true, &total, nullptr);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(cnt.data(), p->d_count, (size_t)ng * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) {
    p->nrounds = total;
    const size_t nr = std::max<size_t>(1, total);
    e = cudaMallocAsync(&p->d_rounds, (nr + 1) * sizeof(PlanRound), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_rminmax, 2 * nr * sizeof(int32_t), ctx->stream);
    if (e == cudaSuccess) {
      plan_init_kernel<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, ctx->stream>>>((int64_t)nr, p->d_rounds, d_rminmax,
                                                                                    d_rminmax + nr);
      p->balanced = best_w > 0 && balance;
      if (p->balanced) {
        e = cudaMallocAsync(&d_sinfo, (size_t)p->nchunks * best_ch * sizeof(unsigned), ctx->stream);
        if (e == cudaSuccess) {
          plan_fill_balanced_kernel<<<(unsigned)p->nchunks, best_ch, 0, ctx->stream>>>(
              d_g_point, nwav, ng, best_w, p->d_round_off, p->d_rounds, d_rminmax, d_rminmax + nr, d_sinfo);
          plan_colour_kernel<<<(unsigned)((2 * nr + COL_WARPS - 1) / COL_WARPS), COL_WARPS * 32, 0, ctx->stream>>>(
              (int64_t)total, best_ch, p->d_round_off, p->d_rounds, d_sinfo);
          ctx->launches++;
        }
      } else
        plan_fill_kernel<<<(unsigned)p->nchunks, best_ch, 0, ctx->stream>>>(d_g_point, nwav, ng, best_w, p->d_round_off,
                                                                            p->d_rounds, d_rminmax, d_rminmax + nr);
      ctx->launches += 2;
      if (best_w > 0) {
        plan_window_kernel<<<1, 32, 0, ctx->stream>>>((int64_t)nr, best_w, d_rminmax, d_rminmax + nr, p->d_rounds);
        ctx->launches++;
      }
      e = cudaGetLastError();
    }
    if (e == cudaSuccess && p->balanced) {
      // reduce_kernel (kinds with per-row weight sums) cannot run on a balanced plan: keep the map so
      // that a free-layout plan can be built the first time such a kind is asked for
      e = cudaMallocAsync(&p->d_g_copy, (size_t)nwav * sizeof(int32_t), ctx->stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(p->d_g_copy, d_g_point, (size_t)nwav * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // d_g_point may be freed by the caller
  }
  if (d_rminmax) cudaFreeAsync(d_rminmax, ctx->stream);
  if (d_sinfo) cudaFreeAsync(d_sinfo, ctx->stream);
  if (d_nfree) cudaFreeAsync(d_nfree, ctx->stream);
  if (d_nrounds) cudaFreeAsync(d_nrounds, ctx->stream);
  if (e != cudaSuccess) {
    set_error("ecckd_gpoint_plan_create: %s", cudaGetErrorString(e));
    plan_free(p);
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  p->h_count.assign(cnt.begin(), cnt.end());
  *out = p;
  return ECCKD_OK;
}

// The plan-shaping environment knobs (read when a plan is built) as one word: a cached plan is only
// reused under the settings it was built with.
static uint64_t plan_env_signature() {
  uint64_t sig = 1469598103934665603ull;
  for (const char* name : {"ECCKD_PLAN_WINDOW", "ECCKD_PLAN_BALANCE", "ECCKD_CH"}) {
    const char* v = getenv(name);
    for (const char* c = v ? v : "-"; *c; ++c) sig = (sig ^ (unsigned char)*c) * 1099511628211ull;
    sig = (sig ^ 0xffu) * 1099511628211ull;
  }
  return sig;
}

ECCKD_API int ecckd_gpoint_plan_create(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng,
                                       ecckd_gpoint_plan** out) {
  ECCKD_DEVICE(ctx);
  const int rc = plan_create_impl(ctx, d_g_point, nwav, ng, false, out);
  if (rc == ECCKD_OK && out && *out) (*out)->env_sig = plan_env_signature();
  return rc;
}

// Test / diagnostic hook: the plan's rounds as the kernels see them.  h_rounds receives up to
// capacity_rounds records of 1168 bytes (uint16 off[16][32], uint32 meta[32], steps, chunk, flags, win);
// info = {nrounds, chunk length, windowed, balanced}.
ECCKD_API int ecckd_gpoint_plan_export(ecckd_gpoint_plan* plan, void* h_rounds, int64_t capacity_rounds,
                                       int64_t* info) {
  ECCKD_DEVICE(plan ? plan->ctx : nullptr);
  ECCKD_REQUIRE(plan && info, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_export: NULL argument");
  info[0] = plan->nrounds;
  info[1] = plan->ch;
  info[2] = plan->windowed ? 1 : 0;
  info[3] = plan->balanced ? 1 : 0;
  if (h_rounds && capacity_rounds > 0) {
    const int64_t n = std::min<int64_t>(capacity_rounds, plan->nrounds);
    ECCKD_CUDA_OK(cudaMemcpyAsync(h_rounds, plan->d_rounds, (size_t)n * sizeof(PlanRound), cudaMemcpyDeviceToHost,
                                  plan->ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(plan->ctx->stream));
  }
  return ECCKD_OK;
}

namespace ecckd {
__global__ void debug_eval_kernel(int which, TransConst K, TransConst256 K2, TransConst64 K3, double kd, int64_t n,
                                  const double* __restrict__ x, double* __restrict__ out) {
  __shared__ __align__(16) double s_tab[256];
  __shared__ __align__(16) double s_tab64[64 * 16];
  trans_tab_init(s_tab);
  trans_tab64_init(s_tab64);
  __syncthreads();
  const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_tab);
  const unsigned char* tab64 = reinterpret_cast<const unsigned char*>(s_tab64);
  const unsigned lane_off = (threadIdx.x & 15u) << 3;
  const TransLane tl = trans_lane_init();
  // whole warps stay in the loop (the evaluations shuffle): pad with a harmless argument
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const double v = i < n ? x[i] : 1.0;
    const double r = which == 0 ? trans_one_minus_exp(v, K, tl)
                   : which == 2 ? trans256_one_minus_exp(v, K2, tab)
                   : which == 3 ? trans64_one_minus_exp(v, K3, tab64, lane_off)
                                : planck_elem_fast((6.62606896e-34 / 1.3806504e-23) / kd, v, 1.0, tl);
    if (i < n) out[i] = r;
  }
}
}  // namespace ecckd

ECCKD_API int ecckd_debug_eval(ecckd_ctx* ctx, int which, double kd, int64_t n, const double* x, double* out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && x && out && n > 0 && which >= 0 && which <= 3 && kd > 0.0, ECCKD_PARAMETER_ERROR,
                "ecckd_debug_eval: bad argument");
  double* d = nullptr;
  ECCKD_CUDA_OK(cudaMalloc(&d, (size_t)2 * n * sizeof(double)));
  TransConst K;
  trans_const_init(&K, kd);
  TransConst256 K2;
  trans_const256_init(&K2, kd);
  TransConst64 K3;
  trans_const64_init(&K3, kd);
  cudaError_t e = cudaMemcpyAsync(d, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    debug_eval_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, ctx->stream>>>(which, K, K2, K3, kd, n, d, d + n);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) {
    set_error("ecckd_debug_eval: %s", cudaGetErrorString(e));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_destroy(ecckd_gpoint_plan* plan) {
  ECCKD_DEVICE(plan ? plan->ctx : nullptr);
  if (!plan) return ECCKD_OK;
  if (plan->ctx) cudaStreamSynchronize(plan->ctx->stream);
  plan_free(plan);
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_counts(ecckd_gpoint_plan* plan, int64_t* h_count) {
  ECCKD_DEVICE(plan ? plan->ctx : nullptr);
  ECCKD_REQUIRE(plan && h_count, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_counts: NULL argument");
  for (int g = 0; g < plan->ng; ++g) h_count[g] = plan->h_count[g];
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_plan_counts_dev(ecckd_gpoint_plan* plan, int64_t** d_count) {
  ECCKD_DEVICE(plan ? plan->ctx : nullptr);
  ECCKD_REQUIRE(plan && d_count, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_plan_counts_dev: NULL argument");
  *d_count = reinterpret_cast<int64_t*>(plan->d_count);
  return ECCKD_OK;
}

ECCKD_API int ecckd_planck_function_dev(ecckd_ctx* ctx, int nt, const double* h_temperature, int64_t nwav,
                                        const double* d_wn, const double* d_dwn, double* d_ans,
                                        int64_t ans_stride) {
  ECCKD_DEVICE(ctx);
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
  ECCKD_DEVICE(ctx);
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
  ECCKD_DEVICE(ctx);
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
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && plan, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: NULL argument");
  ECCKD_REQUIRE(nspec > 0 && nlay > 0, ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth: nspec and nlay must be > 0");
  // partial sums live in the plan-independent scratch slot 3 tail; sized [nrows][ng][NSUM + NMINMAX]
  const size_t nrg = (size_t)nspec * nlay * plan->ng;
  if (plan->scratch_bytes < nrg * (NSUM + NMINMAX) * sizeof(double)) {
    if (plan->d_scratch) ECCKD_CUDA_OK(cudaFreeAsync(plan->d_scratch, ctx->stream));
    plan->d_scratch = nullptr;
    plan->scratch_bytes = 0;
    ECCKD_CUDA_OK(cudaMallocAsync(&plan->d_scratch, nrg * (NSUM + NMINMAX) * sizeof(double), ctx->stream));
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
  ECCKD_DEVICE(ctx);
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

// float32 spectra (how the CKDMIP files store optical depth, merge_spectra.cpp:108; the reference widens them
// on read, DataFileEngineNetcdf.cpp:393-397) widened on the device: exact, and half the PCIe bytes.
__global__ void widen_f32_kernel(int64_t n, const float* __restrict__ src, double* __restrict__ dst) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(src + i);
    *reinterpret_cast<double2*>(dst + i) = make_double2((double)v.x, (double)v.y);
  } else if (i < n) {
    dst[i] = (double)src[i];
  }
}

__global__ void g_point_differs_kernel(int64_t n, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                       int32_t* __restrict__ differs) {
  bool d = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d |= a[i] != b[i];
  if (__syncthreads_or(d) && threadIdx.x == 0) *differs = 1;
}

// What an ecckd_lut_builder keeps on the device between blocks: the g-point map and its plan, and the
// per-wavenumber weight inputs (wavenumber + spacing for Planck weights, or one weight row).
struct ResidentGrid {
  const int32_t* d_g;
  const double* d_w0;  // wavenumber_cm_1, or the weight row (ssi)
  const double* d_w1;  // d_wavenumber_cm_1 (Planck weights only)
  ecckd_gpoint_plan* plan;
};

// Host-buffer driver shared by the two reference-facing entry points below.
// Weights: either planck_fl (matrix, or one row when planck_stride == 0) or
// Planck function of temperature_fl computed in the kernel.
static int host_average(ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav,
                        const double* pressure_fl, const double* pressure_hl, const int32_t* g_point,
                        const double* optical_depth, int64_t od_stride, const double* planck_fl,
                        int64_t planck_stride, const double* temperature_fl, const double* wavenumber_cm_1,
                        const double* d_wavenumber_cm_1, const char* averaging_method, double* molar_abs,
                        double* min_molar_abs, double* max_molar_abs, int64_t out_stride, int32_t* warn,
                        const float* optical_depth_f32 = nullptr, const ResidentGrid* grid = nullptr) {
  const int method = ecckd_averaging_method_from_string(averaging_method);
  if (method < 0) {
    set_error("averaging_method \"%s\" not understood", averaging_method ? averaging_method : "(null)");
    return ECCKD_PARAMETER_ERROR;
  }
  ECCKD_REQUIRE(nlay > 0 && nwav > 0 && ng > 0 && od_stride >= nwav && out_stride >= ng, ECCKD_PARAMETER_ERROR,
                "ecckd_average_optical_depth_to_g_point: bad dimensions");
  cudaStream_t st = ctx->stream;
  const bool fused = planck_fl == nullptr;
  const bool vecw = !fused && planck_stride == 0;
  // device staging (even row strides so 16 B copies are legal), cached in the context
  const int64_t dstride = (nwav + 1) & ~int64_t(1);
  const size_t od_elems = (size_t)nlay * dstride;
  const bool resident_w = grid && (fused || vecw);
  const size_t w_elems =
      resident_w ? 0 : (fused ? (size_t)2 * dstride : (vecw ? (size_t)dstride : (size_t)nlay * dstride));
  const size_t out_elems = (size_t)nlay * ng;
  void* arena = nullptr;
  const size_t f32_elems = optical_depth_f32 ? (od_elems + 1) / 2 : 0;  // float staging, counted in doubles
  ECCKD_TRY(ws_reserve(ctx, 4, (od_elems + w_elems + 3 * out_elems + f32_elems) * sizeof(double), &arena));
  double* d_od = static_cast<double*>(arena);
  double* d_w = d_od + od_elems;
  double* d_out = d_w + w_elems;
  float* d_f32 = reinterpret_cast<float*>(d_out + 3 * out_elems);
  void* iarena = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 5, (size_t)(nwav + ng + 1) * sizeof(int32_t), &iarena));
  int32_t* d_g_stage = static_cast<int32_t*>(iarena);
  int32_t* d_warn = d_g_stage + nwav;  // [ng] warnings + 1 "map differs from the cached one" flag
  const int32_t* d_g = grid ? grid->d_g : d_g_stage;
  const double* d_wn = resident_w ? grid->d_w0 : d_w;            // wavenumbers, or the weight row
  const double* d_dwn = resident_w ? grid->d_w1 : d_w + dstride;
  // Layers are independent, so the optical depths (and a weight matrix) go up in slabs of layers on a
  // second stream while the slabs already there are reduced: the call costs the PCIe transfer plus one
  // slab's reduction instead of transfer + plan + reduction in sequence.
  if (!ctx->copy_stream) ECCKD_CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  cudaStream_t cs = ctx->copy_stream;
  ECCKD_CUDA_OK(cudaStreamSynchronize(st));  // the staging arena may still be read by earlier work on st
  // The small per-wavenumber inputs go FIRST and on the copy stream: host-to-device copies are served
  // in submission order per stream, and queued on the compute stream they were only served after the
  // last slab (measured: the plan and every reduction then started when all copies had landed, 12 ms
  // per call).  The compute stream waits for them through an event.
  int rc = ECCKD_OK;
  ecckd_gpoint_plan* plan = nullptr;
  const bool trace = getenv("ECCKD_TRACE") != nullptr;
  auto now_ms = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  const double t_begin = now_ms();
  cudaEvent_t small_ready = nullptr;
  cudaError_t e = cudaSuccess;
  if (!grid) e = cudaMemcpyAsync(d_g_stage, g_point, (size_t)nwav * sizeof(int32_t), cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_warn, 0, (size_t)(ng + 1) * sizeof(int32_t), cs);
  if (resident_w) {
  } else if (e == cudaSuccess && fused) {
    e = cudaMemcpyAsync(d_w, wavenumber_cm_1, nwav * sizeof(double), cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_w + dstride, d_wavenumber_cm_1, nwav * sizeof(double), cudaMemcpyHostToDevice, cs);
  } else if (e == cudaSuccess && vecw) {
    e = cudaMemcpyAsync(d_w, planck_fl, nwav * sizeof(double), cudaMemcpyHostToDevice, cs);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&small_ready, trace ? cudaEventDefault : cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(small_ready, cs);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, small_ready, 0);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(cs);
    if (small_ready) cudaEventDestroy(small_ready);
    set_error("ecckd_average_optical_depth_to_g_point: %s", cudaGetErrorString(e));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  // slabs of decreasing size (weights nslab, nslab-1, ..., 1): what is left to do when the last copy
  // lands is the reduction of a slab of one or two layers
  const int nslab = std::min(nlay, 8);
  std::vector<int> slab_begin(nslab + 1, 0);
  {
    const int64_t tri = (int64_t)nslab * (nslab + 1) / 2;
    int64_t acc_w = 0;
    for (int s = 0; s < nslab; ++s) {
      acc_w += nslab - s;
      slab_begin[s + 1] = std::max(slab_begin[s] + 1, (int)((nlay * acc_w + tri / 2) / tri));
    }
    slab_begin[nslab] = nlay;
    for (int s = nslab - 1; s > 0; --s) slab_begin[s] = std::min(slab_begin[s], slab_begin[s + 1] - 1);
  }
  // One loop: enqueue the copy of slab s, then the reduction of slab s (the plan is built while slab 0
  // is on its way).  The order matters: a reduction call uploads its small descriptors with host-to-device
  // copies of its own, and those are served in submission order with the slabs - enqueued after ALL
  // slab copies (as the first version did) every reduction waited for the last slab to land.
  std::vector<cudaEvent_t> ready(nslab, nullptr);
  std::vector<cudaEvent_t> reduced;
  auto drop_events = [&]() {
    for (cudaEvent_t ev : ready)
      if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : reduced) cudaEventDestroy(ev);
    reduced.clear();
    if (small_ready) cudaEventDestroy(small_ready);
    small_ready = nullptr;
  };
  const bool mm = min_molar_abs && max_molar_abs;
  double t_enq = 0.0, t_plan = 0.0;
  for (int s = 0; s < nslab && e == cudaSuccess && rc == ECCKD_OK; ++s) {
    const int l0 = slab_begin[s], l1 = slab_begin[s + 1];
    if (optical_depth_f32) {
      e = cudaMemcpy2DAsync(d_f32 + (size_t)l0 * dstride, dstride * sizeof(float),
                            optical_depth_f32 + (size_t)l0 * od_stride, od_stride * sizeof(float), nwav * sizeof(float),
                            l1 - l0, cudaMemcpyHostToDevice, cs);
      if (e == cudaSuccess) {
        const int64_t nel = (int64_t)(l1 - l0) * dstride;  // pad columns carry stale values nobody reads
        widen_f32_kernel<<<(unsigned)((nel / 2 + 256) / 256), 256, 0, cs>>>(nel, d_f32 + (size_t)l0 * dstride,
                                                                          d_od + (size_t)l0 * dstride);
        ctx->launches++;
        e = cudaGetLastError();
      }
    } else {
      e = cudaMemcpy2DAsync(d_od + (size_t)l0 * dstride, dstride * sizeof(double),
                            optical_depth + (size_t)l0 * od_stride, od_stride * sizeof(double), nwav * sizeof(double),
                            l1 - l0, cudaMemcpyHostToDevice, cs);
    }
    if (e == cudaSuccess && !fused && !vecw)
      e = cudaMemcpy2DAsync(d_w + (size_t)l0 * dstride, dstride * sizeof(double), planck_fl + (size_t)l0 * planck_stride,
                            planck_stride * sizeof(double), nwav * sizeof(double), l1 - l0, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ready[s], trace ? cudaEventDefault : cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ready[s], cs);
    if (e != cudaSuccess) break;
    if (s == 0) {
      t_enq = now_ms();
      if (grid) {
        plan = grid->plan;
      } else {
        // the plan of the previous call is reused when the g-point map is the same, element for element
        // (compared on the device: 2 x 4 B per wavenumber of HBM reads, ~15 us at 2^23) and so are the
        // plan-shaping settings; the host waits here for the small inputs only, with slab 0 in flight
        static const bool use_cache = !(getenv("ECCKD_PLAN_CACHE") && atoi(getenv("ECCKD_PLAN_CACHE")) == 0);
        const uint64_t sig = plan_env_signature();
        bool hit = false;
        if (use_cache && ctx->cached_plan && ctx->cached_nwav == nwav && ctx->cached_ng == ng &&
            ctx->cached_plan->env_sig == sig) {
          g_point_differs_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(nwav, d_g, ctx->cached_g, d_warn + ng);
          ctx->launches++;
          int32_t differs = 1;
          e = cudaMemcpyAsync(&differs, d_warn + ng, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          if (e != cudaSuccess) break;
          hit = differs == 0;
        }
        if (hit) {
          plan = ctx->cached_plan;
          ctx->plan_cache_hits++;
        } else {
          rc = ecckd_gpoint_plan_create(ctx, d_g, nwav, ng, &plan);
          if (rc != ECCKD_OK) break;
          ctx->plan_cache_misses++;
          if (use_cache) {
            if (ctx->cached_plan) ecckd_gpoint_plan_destroy(ctx->cached_plan);
            ctx->cached_plan = nullptr;
            if (ctx->cached_nwav != nwav || !ctx->cached_g) {
              if (ctx->cached_g) cudaFree(ctx->cached_g);
              ctx->cached_g = nullptr;
              e = cudaMalloc(&ctx->cached_g, (size_t)nwav * sizeof(int32_t));
            }
            if (e == cudaSuccess)
              e = cudaMemcpyAsync(ctx->cached_g, d_g, (size_t)nwav * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) {
              ecckd_gpoint_plan_destroy(plan);
              plan = nullptr;
              break;
            }
            ctx->cached_plan = plan;
            ctx->cached_nwav = nwav;
            ctx->cached_ng = ng;
          }
        }
      }
      t_plan = now_ms();
    }
    e = cudaStreamWaitEvent(st, ready[s], 0);
    if (e != cudaSuccess) break;
    rc = ecckd_average_optical_depth_to_g_point_dev(
        ctx, plan, reference_surface_vmr, l1 - l0, pressure_fl + l0, pressure_hl + l0, d_od + (size_t)l0 * dstride, dstride,
        fused ? ECCKD_WEIGHT_PLANCK : (vecw ? ECCKD_WEIGHT_VECTOR : ECCKD_WEIGHT_MATRIX),
        fused ? nullptr : (vecw ? d_wn : d_w + (size_t)l0 * dstride), (fused || vecw) ? 0 : dstride,
        fused ? temperature_fl + l0 : nullptr, fused ? d_wn : nullptr, fused ? d_dwn : nullptr, method,
        d_out + (size_t)l0 * ng, mm ? d_out + out_elems + (size_t)l0 * ng : nullptr,
        mm ? d_out + 2 * out_elems + (size_t)l0 * ng : nullptr, ng, d_warn);
    if (trace) {
      cudaEvent_t ev;
      cudaEventCreate(&ev);
      cudaEventRecord(ev, st);
      reduced.push_back(ev);
    }
  }
  if (e != cudaSuccess) {
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(st);
    drop_events();
    if (plan && !grid && plan != ctx->cached_plan) ecckd_gpoint_plan_destroy(plan);
    set_error("ecckd_average_optical_depth_to_g_point: %s", cudaGetErrorString(e));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  // Results: one contiguous device-to-host copy of {molar_abs, min, max} into the context's pinned
  // staging buffer, rows copied to the caller's (possibly strided, pageable) arrays on the host after
  // the synchronisation.  (Three 2-D copies into pageable memory, one row of 512 B at a time, cost ~4 ms.)
  const size_t res_bytes = (mm ? 3 : 1) * out_elems * sizeof(double);
  const bool staged_out = res_bytes + 256 + (size_t)ng * sizeof(int32_t) <= ctx->pinned_bytes;
  double* h_res = nullptr;
  int32_t* h_warn = nullptr;
  if (rc == ECCKD_OK && e == cudaSuccess) {
    if (staged_out) {
      e = cudaStreamSynchronize(st);  // the staging ring may hold descriptor uploads still in flight
      ctx->pinned_off = 0;
      h_res = static_cast<double*>(ctx->pinned);
      h_warn = reinterpret_cast<int32_t*>(static_cast<char*>(ctx->pinned) + ((res_bytes + 255) & ~(size_t)255));
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_res, d_out, res_bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && warn)
        e = cudaMemcpyAsync(h_warn, d_warn, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    } else {
      e = cudaMemcpy2DAsync(molar_abs, out_stride * sizeof(double), d_out, ng * sizeof(double), ng * sizeof(double), nlay,
                            cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && mm) {
        e = cudaMemcpy2DAsync(min_molar_abs, out_stride * sizeof(double), d_out + out_elems, ng * sizeof(double),
                              ng * sizeof(double), nlay, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
          e = cudaMemcpy2DAsync(max_molar_abs, out_stride * sizeof(double), d_out + 2 * out_elems, ng * sizeof(double),
                                ng * sizeof(double), nlay, cudaMemcpyDeviceToHost, st);
      }
      if (e == cudaSuccess && warn)
        e = cudaMemcpyAsync(warn, d_warn, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
  }
  const double t_launched = now_ms();
  cudaStreamSynchronize(cs);
  const double t_copied = now_ms();
  if (trace) {
    cudaStreamSynchronize(st);
    for (int s2 = 0; s2 < nslab; ++s2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, small_ready, ready[s2]);
      float ms2 = 0.f;
      if ((size_t)s2 < reduced.size()) cudaEventElapsedTime(&ms2, small_ready, reduced[s2]);
      fprintf(stderr, "  slab %d (layers %d-%d) landed %.2f ms, reduced %.2f ms after the small inputs\n", s2,
              slab_begin[s2], slab_begin[s2 + 1] - 1, ms, ms2);
    }
  }
  drop_events();
  const cudaError_t e2 = cudaStreamSynchronize(st);
  if (h_res && rc == ECCKD_OK && e == cudaSuccess && e2 == cudaSuccess) {
    double* dst[3] = {molar_abs, mm ? min_molar_abs : nullptr, mm ? max_molar_abs : nullptr};
    for (int a = 0; a < (mm ? 3 : 1); ++a)
      for (int l = 0; l < nlay; ++l)
        memcpy(dst[a] + (size_t)l * out_stride, h_res + a * out_elems + (size_t)l * ng, (size_t)ng * sizeof(double));
    if (warn) memcpy(warn, h_warn, (size_t)ng * sizeof(int32_t));
  }
  const double t_done = now_ms();
  if (plan && !grid && plan != ctx->cached_plan) ecckd_gpoint_plan_destroy(plan);
  if (trace)
    fprintf(stderr, "host_average: enqueue copies %.2f, plan ready %.2f, launches done %.2f, copies landed %.2f, "
            "all done %.2f, plan freed %.2f ms\n", t_enq - t_begin, t_plan - t_begin, t_launched - t_begin,
            t_copied - t_begin, t_done - t_begin, now_ms() - t_begin);
  if (rc != ECCKD_OK) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) {
    set_error("ecckd_average_optical_depth_to_g_point: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_average_optical_depth_to_g_point(ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay,
                                                     int64_t nwav, const double* pressure_fl,
                                                     const double* pressure_hl, const int32_t* g_point,
                                                     const double* optical_depth, int64_t od_stride,
                                                     const double* planck_fl, int64_t planck_stride,
                                                     const char* averaging_method, double* molar_abs,
                                                     double* min_molar_abs, double* max_molar_abs, int64_t out_stride,
                                                     int32_t* warn) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && pressure_fl && pressure_hl && g_point && optical_depth && planck_fl && molar_abs,
                ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth_to_g_point: NULL argument");
  return host_average(ctx, ng, reference_surface_vmr, nlay, nwav, pressure_fl, pressure_hl, g_point, optical_depth,
                      od_stride, planck_fl, planck_stride, nullptr, nullptr, nullptr, averaging_method, molar_abs,
                      min_molar_abs, max_molar_abs, out_stride, warn);
}

ECCKD_API int ecckd_planck_average_optical_depth_to_g_point(
    ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav, const double* pressure_fl,
    const double* pressure_hl, const int32_t* g_point, const double* optical_depth, int64_t od_stride,
    const double* temperature_fl, const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
    const char* averaging_method, double* molar_abs, double* min_molar_abs, double* max_molar_abs, int64_t out_stride,
    int32_t* warn) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && pressure_fl && pressure_hl && g_point && optical_depth && temperature_fl && wavenumber_cm_1 &&
                    d_wavenumber_cm_1 && molar_abs,
                ECCKD_PARAMETER_ERROR, "ecckd_planck_average_optical_depth_to_g_point: NULL argument");
  return host_average(ctx, ng, reference_surface_vmr, nlay, nwav, pressure_fl, pressure_hl, g_point, optical_depth,
                      od_stride, nullptr, 0, temperature_fl, wavenumber_cm_1, d_wavenumber_cm_1, averaging_method,
                      molar_abs, min_molar_abs, max_molar_abs, out_stride, warn);
}

ECCKD_API int ecckd_planck_average_optical_depth_to_g_point_f32(
    ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav, const double* pressure_fl,
    const double* pressure_hl, const int32_t* g_point, const float* optical_depth, int64_t od_stride,
    const double* temperature_fl, const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
    const char* averaging_method, double* molar_abs, double* min_molar_abs, double* max_molar_abs, int64_t out_stride,
    int32_t* warn) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && pressure_fl && pressure_hl && g_point && optical_depth && temperature_fl && wavenumber_cm_1 &&
                    d_wavenumber_cm_1 && molar_abs,
                ECCKD_PARAMETER_ERROR, "ecckd_planck_average_optical_depth_to_g_point_f32: NULL argument");
  return host_average(ctx, ng, reference_surface_vmr, nlay, nwav, pressure_fl, pressure_hl, g_point, nullptr, od_stride,
                      nullptr, 0, temperature_fl, wavenumber_cm_1, d_wavenumber_cm_1, averaging_method, molar_abs,
                      min_molar_abs, max_molar_abs, out_stride, warn, optical_depth);
}

// One (gas, concentration, temperature) block of the create_look_up_table driver loop:
// src/ecckd/create_look_up_table.cpp:293-344 (and :368-415, :448-495 for the other concentration
// dependences, which differ only in what is read).  From the half-level profile of the block it forms
// pressure_fl (:305) and the pressure-weighted full-level temperature (:316-317), then averages with
// Planck weights at that temperature (longwave, :320-331) or with the solar spectrum replicated over
// the layers (shortwave, ssi != NULL, :333-341).
ECCKD_API int ecckd_create_lut_block(ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav,
                                     const double* pressure_hl, const double* temperature_hl, const int32_t* g_point,
                                     const double* optical_depth, int64_t od_stride, const double* wavenumber_cm_1,
                                     const double* d_wavenumber_cm_1, const double* ssi, const char* averaging_method,
                                     double* temperature_fl, double* molar_abs, double* min_molar_abs,
                                     double* max_molar_abs, int64_t out_stride, int32_t* warn) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && pressure_hl && temperature_hl && g_point && optical_depth && molar_abs && nlay > 0,
                ECCKD_PARAMETER_ERROR, "ecckd_create_lut_block: NULL argument");
  ECCKD_REQUIRE(ssi || (wavenumber_cm_1 && d_wavenumber_cm_1), ECCKD_PARAMETER_ERROR,
                "ecckd_create_lut_block: longwave blocks need the wavenumber grid, shortwave ones ssi");
  std::vector<double> p_fl(nlay), t_fl(nlay);
  for (int i = 0; i < nlay; ++i) {
    p_fl[i] = 0.5 * (pressure_hl[i] + pressure_hl[i + 1]);                                   // :305
    const double t_x_p0 = temperature_hl[i] * pressure_hl[i], t_x_p1 = temperature_hl[i + 1] * pressure_hl[i + 1];
    t_fl[i] = 0.5 * (t_x_p0 + t_x_p1) / p_fl[i];                                             // :316-317
    if (temperature_fl) temperature_fl[i] = t_fl[i];
  }
  if (ssi) {
    return host_average(ctx, ng, reference_surface_vmr, nlay, nwav, p_fl.data(), pressure_hl, g_point, optical_depth,
                        od_stride, ssi, 0, nullptr, nullptr, nullptr, averaging_method, molar_abs, min_molar_abs,
                        max_molar_abs, out_stride, warn);
  }
  return host_average(ctx, ng, reference_surface_vmr, nlay, nwav, p_fl.data(), pressure_hl, g_point, optical_depth,
                      od_stride, nullptr, 0, t_fl.data(), wavenumber_cm_1, d_wavenumber_cm_1, averaging_method, molar_abs,
                      min_molar_abs, max_molar_abs, out_stride, warn);
}

// The block loop of create_look_up_table (create_look_up_table.cpp:249-591) reads the g-point map once
// (:84-148), the wavenumber grid with the first block and then calls the averaging for every (gas,
// concentration, temperature) block with the SAME map and grid.  An ecckd_lut_builder is that loop's
// state on the device: map, plan and grid (or the solar spectrum) are uploaded once, every block then
// costs its own optical depths over PCIe and nothing else.
struct ecckd_lut_builder {
  ecckd_ctx* ctx = nullptr;
  int ng = 0;
  int64_t nwav = 0;
  bool sw = false;
  int32_t* d_g = nullptr;
  double* d_w = nullptr;  // [2][dstride] wavenumber, spacing (longwave) or [dstride] ssi (shortwave)
  int64_t dstride = 0;
  ecckd_gpoint_plan* plan = nullptr;
};

ECCKD_API int ecckd_lut_builder_destroy(ecckd_lut_builder* b) {
  if (!b) return ECCKD_OK;
  ECCKD_DEVICE(b->ctx);
  cudaStreamSynchronize(b->ctx->stream);
  if (b->plan) ecckd_gpoint_plan_destroy(b->plan);
  if (b->d_g) cudaFree(b->d_g);
  if (b->d_w) cudaFree(b->d_w);
  delete b;
  return ECCKD_OK;
}

ECCKD_API int ecckd_lut_builder_create(ecckd_ctx* ctx, int ng, int64_t nwav, const int32_t* g_point,
                                       const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
                                       const double* ssi, ecckd_lut_builder** out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && g_point && out && ng > 0 && nwav > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_lut_builder_create: NULL argument or bad dimensions");
  ECCKD_REQUIRE(ssi || (wavenumber_cm_1 && d_wavenumber_cm_1), ECCKD_PARAMETER_ERROR,
                "ecckd_lut_builder_create: longwave builders need the wavenumber grid, shortwave ones ssi");
  *out = nullptr;
  ecckd_lut_builder* b = new ecckd_lut_builder;
  b->ctx = ctx;
  b->ng = ng;
  b->nwav = nwav;
  b->sw = ssi != nullptr;
  b->dstride = (nwav + 1) & ~int64_t(1);
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMalloc(&b->d_g, (size_t)nwav * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_w, (size_t)(b->sw ? 1 : 2) * b->dstride * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_g, g_point, (size_t)nwav * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(b->d_w, b->sw ? ssi : wavenumber_cm_1, (size_t)nwav * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && !b->sw)
    e = cudaMemcpyAsync(b->d_w + b->dstride, d_wavenumber_cm_1, (size_t)nwav * sizeof(double), cudaMemcpyHostToDevice, st);
  int rc = ECCKD_OK;
  if (e == cudaSuccess) rc = ecckd_gpoint_plan_create(ctx, b->d_g, nwav, ng, &b->plan);
  if (e == cudaSuccess && rc == ECCKD_OK) e = cudaStreamSynchronize(st);  // the host arrays may go away
  if (e != cudaSuccess || rc != ECCKD_OK) {
    if (e != cudaSuccess) set_error("ecckd_lut_builder_create: %s", cudaGetErrorString(e));
    ecckd_lut_builder_destroy(b);
    return rc != ECCKD_OK ? rc : ECCKD_UNEXPECTED_EXCEPTION;
  }
  *out = b;
  return ECCKD_OK;
}

// One block: the arguments of ecckd_create_lut_block minus what the builder holds.  Exactly one of
// optical_depth (float64) and optical_depth_f32 (the float32 values the spectra files store, widened on
// the device: same results, half the PCIe bytes) is given.
ECCKD_API int ecckd_lut_builder_block(ecckd_lut_builder* b, double reference_surface_vmr, int nlay,
                                      const double* pressure_hl, const double* temperature_hl,
                                      const double* optical_depth, const float* optical_depth_f32, int64_t od_stride,
                                      const char* averaging_method, double* temperature_fl, double* molar_abs,
                                      double* min_molar_abs, double* max_molar_abs, int64_t out_stride, int32_t* warn) {
  ECCKD_REQUIRE(b && pressure_hl && molar_abs && nlay > 0 && ((optical_depth != nullptr) != (optical_depth_f32 != nullptr)),
                ECCKD_PARAMETER_ERROR, "ecckd_lut_builder_block: NULL argument, or not exactly one optical depth array");
  ECCKD_REQUIRE(b->sw || temperature_hl, ECCKD_PARAMETER_ERROR, "ecckd_lut_builder_block: longwave blocks need temperature_hl");
  ECCKD_DEVICE(b->ctx);
  std::vector<double> p_fl(nlay), t_fl(nlay, 0.0);
  for (int i = 0; i < nlay; ++i) {
    p_fl[i] = 0.5 * (pressure_hl[i] + pressure_hl[i + 1]);  // create_look_up_table.cpp:305
    if (temperature_hl) {                                    // :316-317
      const double t_x_p0 = temperature_hl[i] * pressure_hl[i], t_x_p1 = temperature_hl[i + 1] * pressure_hl[i + 1];
      t_fl[i] = 0.5 * (t_x_p0 + t_x_p1) / p_fl[i];
    }
    if (temperature_fl) temperature_fl[i] = t_fl[i];
  }
  const ResidentGrid grid{b->d_g, b->d_w, b->sw ? nullptr : b->d_w + b->dstride, b->plan};
  // shortwave: one weight row (planck_fl = the resident ssi, stride 0); longwave: Planck weights in the kernel
  static const double resident_row = 0.0;  // non-NULL marker: the row itself is on the device already
  return host_average(b->ctx, b->ng, reference_surface_vmr, nlay, b->nwav, p_fl.data(), pressure_hl, nullptr, optical_depth,
                      od_stride, b->sw ? &resident_row : nullptr, 0, b->sw ? nullptr : t_fl.data(), nullptr, nullptr,
                      averaging_method, molar_abs, min_molar_abs, max_molar_abs, out_stride, warn, optical_depth_f32, &grid);
}

ECCKD_API int ecckd_merge_axpy_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, double* d_optical_depth, int64_t stride,
                                   const double* d_od, int64_t od_stride, double scaling,
                                   const double* h_scaling_profile) {
  ECCKD_DEVICE(ctx);
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
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_g_point && d_wn && d_dwn && d_out, ECCKD_PARAMETER_ERROR, "ecckd_gpoint_fraction: NULL argument");
  double bin0, dwav;
  ECCKD_TRY(check_uniform_bins(nbin, h_w1, h_w2, &bin0, &dwav));
  // scratch: [ng*nbin + ng] 128-bit cells, then the largest magnitude
  void* d_tmp = nullptr;
  const size_t n = (size_t)ng * nbin + ng;
  ECCKD_TRY(ws_reserve(ctx, 1, n * sizeof(Fix128) + 16, &d_tmp));
  ECCKD_CUDA_OK(cudaMemsetAsync(d_tmp, 0, n * sizeof(Fix128) + 16, ctx->stream));
  Fix128* cells = static_cast<Fix128*>(d_tmp);
  unsigned long long* maxbits = reinterpret_cast<unsigned long long*>(cells + n);
  const unsigned nblk = (unsigned)((nwav + 255) / 256);
  hist_max_kernel<<<nblk, 256, 0, ctx->stream>>>(ng, nwav, d_g_point, d_dwn, maxbits);
  ECCKD_LAUNCH_CHECK(ctx);
  gfrac_hist_kernel<<<nblk, 256, 0, ctx->stream>>>(ng, nwav, d_g_point, d_wn, d_dwn, nbin, bin0, dwav, maxbits, cells);
  ECCKD_LAUNCH_CHECK(ctx);
  gfrac_norm_kernel<<<(ng * nbin + 255) / 256, 256, 0, ctx->stream>>>(ng, nbin, cells, maxbits, d_out);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_ssi_intervals_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wn, const double* d_ssi, int nbin,
                                      const double* h_w1, const double* h_w2, double* d_out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_wn && d_ssi && d_out, ECCKD_PARAMETER_ERROR, "ecckd_ssi_intervals: NULL argument");
  double bin0, dwav;
  ECCKD_TRY(check_uniform_bins(nbin, h_w1, h_w2, &bin0, &dwav));
  void* d_tmp = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, (size_t)nbin * sizeof(Fix128) + 16, &d_tmp));
  ECCKD_CUDA_OK(cudaMemsetAsync(d_tmp, 0, (size_t)nbin * sizeof(Fix128) + 16, ctx->stream));
  Fix128* cells = static_cast<Fix128*>(d_tmp);
  unsigned long long* maxbits = reinterpret_cast<unsigned long long*>(cells + nbin);
  const unsigned nblk = (unsigned)((nwav + 255) / 256);
  hist_max_kernel<<<nblk, 256, 0, ctx->stream>>>(0, nwav, nullptr, d_ssi, maxbits);
  ECCKD_LAUNCH_CHECK(ctx);
  ssi_hist_kernel<<<nblk, 256, 0, ctx->stream>>>(nwav, d_wn, d_ssi, nbin, bin0, dwav, maxbits, cells);
  ECCKD_LAUNCH_CHECK(ctx);
  fix128_out_kernel<<<(nbin + 255) / 256, 256, 0, ctx->stream>>>(nbin, cells, maxbits, d_out);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// Sums of nrow per-wavenumber rows over the wavenumbers of each g point (lw_spectra.cpp:250-255,
// scale_lut.cpp:119-123: sum(flux(__, find(g_point == ig)), 1)): the reduction with f = 1 and the rows as
// the weight matrix, i.e. the same fixed-order kernels as the averaging.
// The merge loop of read_merged_spectrum.cpp:149-168 for sources with a scalar scaling, in ONE pass:
// optical_depth = 0, then += od_k * scaling_k (or += od_k when scaling_k == 1) for k = 0 .. nsrc-1 in that
// order - the same operations in the same order as the loop of AXPY calls, without re-reading and
// re-writing the running sum nsrc times.
constexpr int MERGE_MAX_SRC = 8;
struct MergeSrc {
  const double* od[MERGE_MAX_SRC];
  int64_t stride[MERGE_MAX_SRC];
  double scaling[MERGE_MAX_SRC];
  int nsrc;
};
__global__ void merge_spectra_kernel(int nlay, int64_t nwav, double* __restrict__ dst, int64_t stride, const MergeSrc m) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwav) return;
  for (int l = blockIdx.y; l < nlay; l += gridDim.y) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < MERGE_MAX_SRC; ++k) {
      if (k < m.nsrc) {
        const double v = m.od[k][(int64_t)l * m.stride[k] + w];
        acc = (m.scaling[k] != 1.0) ? fma(v, m.scaling[k], acc) : acc + v;  // read_merged_spectrum.cpp:160-167
      }
    }
    dst[(int64_t)l * stride + w] = acc;
  }
}
// the same sums two wavenumbers per load and four per thread (rows and strides aligned to 16 bytes): the scalar
// kernel moves 3.4 TB/s on the 3.6 GB blocks of a LW build
__global__ void __launch_bounds__(256) merge_spectra_vec_kernel(int nlay, int64_t nwav, double* __restrict__ dst,
                                                                 int64_t stride, const MergeSrc m) {
  const int64_t npair = nwav >> 1;
  for (int l = blockIdx.y; l < nlay; l += gridDim.y) {
    double2 acc[2];
    int64_t i[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      i[u] = (int64_t)blockIdx.x * 512 + u * 256 + threadIdx.x;
      acc[u] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < MERGE_MAX_SRC; ++k) {
      if (k < m.nsrc) {
        const double2* src = reinterpret_cast<const double2*>(m.od[k] + (int64_t)l * m.stride[k]);
        double2 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) v[u] = i[u] < npair ? __ldcs(src + i[u]) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          // read_merged_spectrum.cpp:160-167
          acc[u].x = (m.scaling[k] != 1.0) ? fma(v[u].x, m.scaling[k], acc[u].x) : acc[u].x + v[u].x;
          acc[u].y = (m.scaling[k] != 1.0) ? fma(v[u].y, m.scaling[k], acc[u].y) : acc[u].y + v[u].y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (i[u] < npair) reinterpret_cast<double2*>(dst + (int64_t)l * stride)[i[u]] = acc[u];
    if ((nwav & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // the odd last wavenumber
      const int64_t w = nwav - 1;
      double a = 0.0;
      for (int k = 0; k < m.nsrc; ++k) {
        const double v = m.od[k][(int64_t)l * m.stride[k] + w];
        a = (m.scaling[k] != 1.0) ? fma(v, m.scaling[k], a) : a + v;
      }
      dst[(int64_t)l * stride + w] = a;
    }
  }
}
ECCKD_API int ecckd_merge_spectra_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, int nsrc, const double* const* h_d_od,
                                      const int64_t* h_od_stride, const double* h_scaling, double* d_optical_depth,
                                      int64_t stride) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && h_d_od && h_od_stride && h_scaling && d_optical_depth, ECCKD_PARAMETER_ERROR,
                "ecckd_merge_spectra: NULL argument");
  ECCKD_REQUIRE(nlay > 0 && nwav > 0 && nsrc > 0 && nsrc <= MERGE_MAX_SRC && stride >= nwav, ECCKD_PARAMETER_ERROR,
                "ecckd_merge_spectra: bad dimensions (1 to %d sources)", MERGE_MAX_SRC);
  MergeSrc m;
  memset(&m, 0, sizeof(m));
  m.nsrc = nsrc;
  for (int k = 0; k < nsrc; ++k) {
    ECCKD_REQUIRE(h_d_od[k] != nullptr && h_od_stride[k] >= nwav, ECCKD_PARAMETER_ERROR, "ecckd_merge_spectra: source %d", k);
    m.od[k] = h_d_od[k];
    m.stride[k] = h_od_stride[k];
    m.scaling[k] = h_scaling[k] < 0.0 ? 1.0 : h_scaling[k];  // :146-148
  }
  bool vec = aligned16(d_optical_depth) && (stride % 2) == 0 && nwav >= 2;
  for (int k = 0; k < nsrc; ++k) vec = vec && aligned16(m.od[k]) && (m.stride[k] % 2) == 0;
  if (vec) {
    dim3 grid((unsigned)((nwav / 2 + 511) / 512), (unsigned)std::min(nlay, 64));
    merge_spectra_vec_kernel<<<grid, 256, 0, ctx->stream>>>(nlay, nwav, d_optical_depth, stride, m);
  } else {
    dim3 grid((unsigned)((nwav + 255) / 256), (unsigned)std::min(nlay, 64));
    merge_spectra_kernel<<<grid, 256, 0, ctx->stream>>>(nlay, nwav, d_optical_depth, stride, m);
  }
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

ECCKD_API int ecckd_gpoint_sum_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nrow, const double* d_values,
                                   int64_t stride, double* d_out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && plan && d_values && d_out && nrow > 0 && stride >= plan->nwav, ECCKD_PARAMETER_ERROR,
                "ecckd_gpoint_sum: bad argument");
  std::vector<RowDesc> rows(nrow);
  std::vector<OutDesc> outs(nrow);
  for (int t = 0; t < nrow; ++t) {
    RowDesc& r = rows[t];
    r.od = d_values;  // any valid row: F_NONE ignores the values
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
  const size_t nrg = (size_t)nrow * plan->ng;
  ECCKD_CUDA_OK(cudaMalloc(&scratch, nrg * (NSUM + NMINMAX) * sizeof(double)));
  double* d_sums = static_cast<double*>(scratch);
  double* d_minmax = d_sums + nrg * NSUM;
  int rc = reduce_partial(ctx, plan, rows, F_NONE, 1.0, W_MATRIX, d_values, stride, nullptr, 0, nullptr, nullptr, d_sums,
                          d_minmax);
  if (rc == ECCKD_OK)
    rc = finalize_rows(ctx, plan->ng, reinterpret_cast<const long long*>(plan->d_count), outs, 1.0, d_sums, d_minmax);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(scratch);
  return rc;
}

// Sum of every row over ALL its n elements (sum(flux, 1) of lw_spectra.cpp:233-234) in a fixed order:
// 256 blocks per row each add a contiguous slice (fixed per-thread strides, fixed tree), one more block
// adds the 256 partial sums in order.
constexpr int ROWSUM_BLOCKS = 256;
__global__ void __launch_bounds__(256) row_sum_partial_kernel(int64_t n, const double* __restrict__ v, int64_t stride,
                                                              double* __restrict__ partial) {
  const int row = blockIdx.y;
  const int64_t per = (n + ROWSUM_BLOCKS - 1) / ROWSUM_BLOCKS;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  const double* x = v + (int64_t)row * stride;
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) s += x[i];
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)row * ROWSUM_BLOCKS + blockIdx.x] = sh[0];
}
__global__ void row_sum_final_kernel(int nrow, const double* __restrict__ partial, double* __restrict__ out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrow) return;
  double s = 0.0;
  for (int b = 0; b < ROWSUM_BLOCKS; ++b) s += partial[(size_t)row * ROWSUM_BLOCKS + b];
  out[row] = s;
}
ECCKD_API int ecckd_row_sum_dev(ecckd_ctx* ctx, int nrow, int64_t n, const double* d_values, int64_t stride, double* d_out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_values && d_out && nrow > 0 && n > 0 && stride >= n, ECCKD_PARAMETER_ERROR, "ecckd_row_sum: bad argument");
  void* d_partial = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 1, (size_t)nrow * ROWSUM_BLOCKS * sizeof(double), &d_partial));
  row_sum_partial_kernel<<<dim3(ROWSUM_BLOCKS, nrow), 256, 0, ctx->stream>>>(n, d_values, stride, static_cast<double*>(d_partial));
  ECCKD_LAUNCH_CHECK(ctx);
  row_sum_final_kernel<<<(nrow + 63) / 64, 64, 0, ctx->stream>>>(nrow, static_cast<const double*>(d_partial), d_out);
  ECCKD_LAUNCH_CHECK(ctx);
  return ECCKD_OK;
}

// The per-profile body of the lw_spectra tool (lw_spectra.cpp:208-255, the branch with g points): Planck
// function at the half levels, the longwave sweep over every wavenumber with unit surface emissivity,
// broadband fluxes, the transmission-averaged optical depth per g point (reference_surface_vmr = 0: raw
// optical depth, Planck weights at the pressure-weighted layer temperatures) and the fluxes summed over
// each g point's wavenumbers.  Scratch: 3 arrays of [nlay+1][nwav] + [nwav] doubles.
__global__ void fill_ones_kernel(int64_t n, double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = 1.0;
}
ECCKD_API int ecckd_lw_spectra_profile_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nlay, const double* d_wn,
                                           const double* d_dwn, const double* h_pressure_hl, const double* h_temperature_hl,
                                           const double* d_optical_depth, int64_t od_stride, double* d_flux_dn_lw,
                                           double* d_flux_up_lw, double* d_spectral_od, double* d_spectral_flux_dn,
                                           double* d_spectral_flux_up) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && plan && d_wn && d_dwn && h_pressure_hl && h_temperature_hl && d_optical_depth && nlay > 0,
                ECCKD_PARAMETER_ERROR, "ecckd_lw_spectra_profile: bad argument");
  ECCKD_REQUIRE(od_stride == plan->nwav, ECCKD_PARAMETER_ERROR,
                "ecckd_lw_spectra_profile: optical_depth must be a contiguous [nlay][nwav] array (the flux solver's layout)");
  const int64_t nwav = plan->nwav;
  const int ng = plan->ng;
  const size_t nhl = (size_t)(nlay + 1) * nwav;
  double* scratch = nullptr;
  ECCKD_CUDA_OK(cudaMalloc(&scratch, (3 * nhl + (size_t)nwav) * sizeof(double)));
  double* planck_hl = scratch;
  double* flux_dn = planck_hl + nhl;
  double* flux_up = flux_dn + nhl;
  double* ones = flux_up + nhl;
  int rc = ecckd_planck_function_dev(ctx, nlay + 1, h_temperature_hl, nwav, d_wn, d_dwn, planck_hl, nwav);  // :211-216
  if (rc == ECCKD_OK) {
    fill_ones_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(nwav, ones);  // :217-218
    ctx->launches++;
    // surf_planck = planck_function(temperature_hl(end)) = the last row of planck_hl
    rc = ecckd_radiative_transfer_lw_dev(ctx, 1, nlay, nwav, planck_hl, d_optical_depth, ones, planck_hl + (size_t)nlay * nwav,
                                         flux_dn, flux_up);  // :223-224 (od_stride must equal nwav for the solver)
  }
  if (rc == ECCKD_OK && d_flux_dn_lw) rc = ecckd_row_sum_dev(ctx, nlay + 1, nwav, flux_dn, nwav, d_flux_dn_lw);  // :230
  if (rc == ECCKD_OK && d_flux_up_lw) rc = ecckd_row_sum_dev(ctx, nlay + 1, nwav, flux_up, nwav, d_flux_up_lw);  // :231
  if (rc == ECCKD_OK && d_spectral_od) {  // :241-249
    std::vector<double> p_fl(nlay), t_fl(nlay);
    for (int l = 0; l < nlay; ++l) {
      p_fl[l] = 0.5 * (h_pressure_hl[l] + h_pressure_hl[l + 1]);
      t_fl[l] = 0.5 * (h_temperature_hl[l] * h_pressure_hl[l] + h_temperature_hl[l + 1] * h_pressure_hl[l + 1]) / p_fl[l];
    }
    rc = ecckd_average_optical_depth_to_g_point_dev(ctx, plan, 0.0, nlay, p_fl.data(), h_pressure_hl, d_optical_depth, od_stride,
                                                    ECCKD_WEIGHT_PLANCK, nullptr, 0, t_fl.data(), d_wn, d_dwn,
                                                    ECCKD_AVG_TRANSMISSION, d_spectral_od, nullptr, nullptr, ng, nullptr);
  }
  if (rc == ECCKD_OK && d_spectral_flux_dn) rc = ecckd_gpoint_sum_dev(ctx, plan, nlay + 1, flux_dn, nwav, d_spectral_flux_dn);  // :250-255
  if (rc == ECCKD_OK && d_spectral_flux_up) rc = ecckd_gpoint_sum_dev(ctx, plan, nlay + 1, flux_up, nwav, d_spectral_flux_up);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(scratch);
  return rc;
}

// scale_lut.cpp:117-133 and :180-181 on the per-g flux sums of ecckd_gpoint_sum_dev (host-sized arrays):
// od_best(iz, ig) = -mu0 log(flux(iz+1, ig) / flux(iz, ig)), or -1 where the flux at the layer base is not
// positive; scaling = od_best / od_total, 1 where od_best <= 0.
ECCKD_API int ecckd_scale_lut_scaling(int nz, int ng, double mu0, const double* h_flux_dn_gpoint, const double* h_od_total,
                                      double* h_od_best, double* h_scaling) {
  ECCKD_REQUIRE(nz > 0 && ng > 0 && h_flux_dn_gpoint && h_od_best, ECCKD_PARAMETER_ERROR, "ecckd_scale_lut_scaling: bad argument");
  for (int ig = 0; ig < ng; ++ig) {
    double flux_top = h_flux_dn_gpoint[ig];
    for (int iz = 0; iz < nz; ++iz) {
      const double flux_base = h_flux_dn_gpoint[(size_t)(iz + 1) * ng + ig];
      h_od_best[(size_t)iz * ng + ig] = (flux_base <= 0.0) ? -1.0 : -mu0 * std::log(flux_base / flux_top);
      flux_top = flux_base;
    }
  }
  if (h_scaling) {
    ECCKD_REQUIRE(h_od_total != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_scale_lut_scaling: od_total missing");
    for (size_t i = 0; i < (size_t)nz * ng; ++i) h_scaling[i] = (h_od_best[i] <= 0.0) ? 1.0 : h_od_best[i] / h_od_total[i];
  }
  return ECCKD_OK;
}

#ifdef ECCKD_PLANCK_TAB
// The Planck look-up table (create_look_up_table.cpp:565-587) is a reduction whose weights differ in every row
// (one temperature each) and whose values are all 1: on reduce2_kernel every row is a group of its own, fed by two
// producer warps (14.3 ms for 231 temperatures x 2^23 wavenumbers).  Here the temperatures are the lanes: a warp
// walks a range of wavenumbers, every lane evaluates the weight at its own temperature (planck_weight_tab, the
// per-wavenumber factor and the g point are warp-uniform and come staged in shared memory) and adds it to its own
// column of the warp's [g point][lane] sums - no two lanes share a cell, the order of the additions is the order of
// the wavenumbers.  A block takes PL_RANGE wavenumbers (fixed, so that the sums do not depend on the grid) for up to
// 8 tiles of 32 temperatures; planck_lut_combine_kernel adds the blocks' sums in block order.
constexpr int PL_STAGE = 1024;
constexpr int PL_RANGE = 8192;
#ifndef ECCKD_PL_ILP
#define ECCKD_PL_ILP 8
#endif
constexpr int PL_ILP = ECCKD_PL_ILP;
constexpr int PL_TAB_BYTES = (1 << ECCKD_TAB_LOGN) * 8;
__global__ void __launch_bounds__(256) planck_lut_kernel(int64_t nwav, int ng, int nt, int ntp,
                                                          const int32_t* __restrict__ g_point,
                                                          const double* __restrict__ wn, const double* __restrict__ pref,
                                                          const double* __restrict__ hkct,
                                                          const double* __restrict__ exp2tab,
                                                          double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  unsigned char* tab = pl_smem;
  double2* s_np = reinterpret_cast<double2*>(pl_smem + PL_TAB_BYTES);  // (wavenumber, prefactor)
  int* s_g = reinterpret_cast<int*>(s_np + PL_STAGE);
  double* bins_all = reinterpret_cast<double*>(s_g + PL_STAGE);
  const int nwarp = (int)blockDim.x >> 5, warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  trans_tabT_init<ECCKD_TAB_LOGN, 0>(reinterpret_cast<double*>(tab), exp2tab);  // the producers' form of the table
  double* bins = bins_all + (size_t)warp * ng * 32 + lane;  // [g point][lane]
  for (int g = 0; g < ng; ++g) bins[g * 32] = 0.0;
  const int t = ((int)blockIdx.y * nwarp + warp) * 32 + lane;
  const double hk = hkct[t < nt ? t : nt - 1];
  const int64_t w0 = (int64_t)blockIdx.x * PL_RANGE;
  const int64_t w1 = w0 + PL_RANGE < nwav ? w0 + PL_RANGE : nwav;
  for (int64_t base = w0; base < w1; base += PL_STAGE) {
    const int n = (int)(w1 - base < PL_STAGE ? w1 - base : PL_STAGE);
    __syncthreads();
    for (int i = threadIdx.x; i < PL_STAGE; i += blockDim.x) {
      const bool in = i < n;
      const int g = in ? g_point[base + i] : -1;
      // a wavenumber outside the map (or past the end) keeps a harmless argument and adds to nobody
      s_np[i] = make_double2(in ? wn[base + i] : 1.0, in ? pref[base + i] : 0.0);
      s_g[i] = (g >= 0 && g < ng) ? g * 32 : -1;
    }
    __syncthreads();
    // PL_ILP independent evaluations in flight (two warps per scheduler cannot hide the latency of one chain), then
    // their additions in wavenumber order
    for (int i = 0; i < PL_STAGE; i += PL_ILP) {
      if (i >= n) break;
      double w[PL_ILP];
      int o[PL_ILP];
#pragma unroll
      for (int m = 0; m < PL_ILP; m += 4) {
        const int4 q = *reinterpret_cast<const int4*>(s_g + i + m);
        o[m] = q.x, o[m + 1] = q.y, o[m + 2] = q.z, o[m + 3] = q.w;
      }
#pragma unroll
      for (int m = 0; m < PL_ILP; ++m) {
        const double2 a = s_np[i + m];
        w[m] = planck_weight_tab(hk, a.x, a.y, tab);
      }
#pragma unroll
      for (int m = 0; m < PL_ILP; ++m)
        if (o[m] >= 0) bins[o[m]] += w[m];
    }
  }
  if (t < nt) {
    double* out = partial + (size_t)blockIdx.x * ng * ntp + t;  // [block][g point][temperature]
    for (int g = 0; g < ng; ++g) out[(size_t)g * ntp] = bins[g * 32];
  }
}
// the blocks' sums in block order: eight contiguous runs of blocks per cell, then the eight run sums in run order
__global__ void __launch_bounds__(256) planck_lut_combine_kernel(int nblk, int ng, int nt, int ntp,
                                                                  const double* __restrict__ partial,
                                                                  double* __restrict__ out) {
  __shared__ double s_part[8][32];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;  // g * ntp + t
  const int per = (nblk + 7) / 8;
  const int b0 = part * per, b1 = min(nblk, b0 + per);
  double s = 0.0;
  if (i < ng * ntp)
    for (int b = b0; b < b1; ++b) s += partial[(size_t)b * ng * ntp + i];
  s_part[part][lane] = s;
  __syncthreads();
  if (part != 0 || i >= ng * ntp) return;
  const int g = i / ntp, t = i - g * ntp;
  if (t >= nt) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) s += s_part[k][lane];
  out[(size_t)t * ng + g] = s;
}
static int planck_lut_tiled(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng, const double* d_wn,
                            const double* d_dwn, int nlut, const double* h_temperature_lut, double* d_out, bool* done) {
  *done = false;
  int dev = 0, smem_max = 0;
  ECCKD_CUDA_OK(cudaGetDevice(&dev));
  ECCKD_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = (size_t)PL_TAB_BYTES + (size_t)PL_STAGE * 20;
  const size_t per_warp = (size_t)ng * 32 * sizeof(double);
  if (fixed + per_warp > (size_t)smem_max) return ECCKD_OK;  // too many g points for one warp's sums: the general path
  const int ntile = (nlut + 31) / 32;
  int nwarp = (int)std::min<size_t>(8, ((size_t)smem_max - fixed) / per_warp);
  nwarp = std::min(nwarp, ntile);
  const int ny = (ntile + nwarp - 1) / nwarp;
  const int ntp = ny * nwarp * 32;
  const int64_t nblk = (nwav + PL_RANGE - 1) / PL_RANGE;
  const size_t smem = fixed + per_warp * nwarp;
  std::vector<double> hkct((size_t)nlut);
  for (int i = 0; i < nlut; ++i)  // (h/k)(100 c)/T, as for the producers of reduce2_kernel
    hkct[i] = (double)(((long double)6.62606896e-34 / (long double)1.3806504e-23) * (100.0L * 2.99792458e8L) /
                       (long double)h_temperature_lut[i]);
  void* d_hk = nullptr;
  void* d_pref = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 6, (size_t)nlut * sizeof(double), &d_hk));
  ECCKD_TRY(ws_reserve(ctx, 7, (size_t)nwav * sizeof(double), &d_pref));
  ECCKD_TRY(upload_small(ctx, d_hk, hkct.data(), (size_t)nlut * sizeof(double)));
  const double* d_tab = nullptr;
  ECCKD_TRY(exp2tab_get(ctx, ECCKD_TAB_LOGN, &d_tab));
  double* d_partial = nullptr;
  ECCKD_CUDA_OK(cudaMallocAsync(&d_partial, (size_t)nblk * ng * ntp * sizeof(double), ctx->stream));
  planck_prefactor_kernel<<<(unsigned)((nwav + 255) / 256), 256, 0, ctx->stream>>>(nwav, d_wn, d_dwn,
                                                                                   static_cast<double*>(d_pref));
  cudaError_t e = cudaFuncSetAttribute(planck_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    planck_lut_kernel<<<dim3((unsigned)nblk, (unsigned)ny), nwarp * 32, smem, ctx->stream>>>(
        nwav, ng, nlut, ntp, d_g_point, d_wn, static_cast<const double*>(d_pref), static_cast<const double*>(d_hk),
        d_tab, d_partial);
    planck_lut_combine_kernel<<<(unsigned)((ng * ntp + 31) / 32), 256, 0, ctx->stream>>>((int)nblk, ng, nlut, ntp,
                                                                                         d_partial, d_out);
    ctx->launches += 3;
    e = cudaGetLastError();
  }
  cudaFreeAsync(d_partial, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ECCKD_CUDA_OK(e);
  *done = true;
  return ECCKD_OK;
}
#endif

ECCKD_API int ecckd_planck_lut_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, const double* d_wn, const double* d_dwn,
                                   int nlut, const double* h_temperature_lut, double* d_out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && plan && d_wn && d_dwn && h_temperature_lut && d_out && nlut > 0, ECCKD_PARAMETER_ERROR,
                "ecckd_planck_lut: bad argument");
#ifdef ECCKD_PLANCK_TAB
  // temperatures across the lanes (planck_lut_kernel) when the plan kept its g-point map
  if (plan->d_g_copy && !getenv("ECCKD_PLANCK_LUT_GENERAL")) {
    bool done = false;
    ECCKD_TRY(planck_lut_tiled(ctx, plan->d_g_copy, plan->nwav, plan->ng, d_wn, d_dwn, nlut, h_temperature_lut, d_out,
                               &done));
    if (done) return ECCKD_OK;
  }
#endif
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
  int rc = reduce_partial(ctx, plan, rows, F_NONE, 1.0, W_PLANCK, nullptr, 0, h_temperature_lut, nlut, d_wn, d_dwn,
                          d_sums, d_minmax);
  if (rc == ECCKD_OK)
    rc = finalize_rows(ctx, plan->ng, reinterpret_cast<const long long*>(plan->d_count), outs, 1.0, d_sums, d_minmax);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(scratch);
  return rc;
}

ECCKD_API int ecckd_compact_g_points_dev(ecckd_ctx* ctx, int ng, int64_t nwav, int32_t* d_g_point,
                                         int32_t* h_g_point_map, int* new_ng) {
  ECCKD_DEVICE(ctx);
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
  ECCKD_DEVICE(ctx);
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
