// reduce_rows.cuh - the per-round work of a consumer warp of the spectral reduction: evaluate the
// 16 plan slots of NR rows and fold the lanes of each g point (see spectral_reduce.cu).  Kept in a
// header so that tools/ubench/body.cu can time it in isolation.
#pragma once
#include "spectral_reduce.cuh"

namespace ecckd {

template <int FKIND>
__device__ __forceinline__ double f_transform(double tau, const TransConst& K, const TransLane& tl, bool is_log,
                                              bool& positive) {
  positive = true;
  if (FKIND == F_LINEAR) return tau;
  if (FKIND == F_TRANS) return trans_one_minus_exp(tau, K, tl);
  if (FKIND == F_SQRT) return sqrt(tau);
  if (FKIND == F_LOG) {
    positive = tau > 0.0;
    return positive ? log(tau) : 0.0;
  }
  if (FKIND == F_HYBRID) {
    // warp-uniform branch (is_log is a property of the row): the shuffles inside stay converged
    if (is_log) {
      positive = tau > 0.0;
      return positive ? log(tau) : 0.0;
    }
    return trans_one_minus_exp(tau, K, tl);
  }
  return 1.0;  // F_NONE: only the weight sum is wanted
}

// Per-lane state of one row over the 16 slots of a round.
//
// Minimum / maximum: a lane starts the round from the row's running extremes of its g point
// (the shared-memory accumulators) and tests every optical depth with ONE unsigned range check on
// its high word: lowb <= hi < lowb + span means "strictly between the extremes' high words", i.e.
// certainly no new extreme.  Only values failing the test (new extremes, values sharing the high
// word of an extreme, negative or non-finite values) take the exact FP64 compare path, which is
// rare once a g point has seen a few hundred values.  This costs 2 issue slots per value instead of
// the 8 of two DSETP + four FSEL, and a round in which no lane of the warp found a new extreme
// skips the cross-lane reduction of the extremes altogether.
struct Partial {
  double s0, s1, mn, mx;
  unsigned lowb, span;
  int zc;
  int upd;  // this lane found a new extreme in this round
};

// hi_cap: high words at or above it never pass the range check (reduce2_kernel's transmission kinds send
// optical depths beyond 708/kD to the exact path, which clamps them; see rows_slots)
__device__ __forceinline__ void minmax_bounds(Partial& p, const unsigned hi_cap = 0x7ff00000u) {
  const unsigned mnhi = (unsigned)__double2hiint(p.mn), mxhi = min((unsigned)__double2hiint(p.mx), hi_cap);
  // +0.0 cannot be undercut by a non-negative value: let high word 0 pass as "inside"
  const unsigned lowb = (p.mn == 0.0 && mnhi == 0u) ? 0u : mnhi + 1u;
  // negative or not yet initialised extremes (sign bit set / mn = +inf, mx = -inf): no fast path
  const bool ok = (int)(mnhi | (unsigned)__double2hiint(p.mx)) >= 0 && mxhi > lowb;
  p.lowb = lowb;
  p.span = ok ? mxhi - lowb : 0u;
}

// Starts a round: extremes from the accumulator entry {sum f*w, sum w, min, max} of the lane's g
// point (unused lanes: a range that accepts everything, so that they never ask for a reduction).
__device__ __forceinline__ void partial_begin(Partial& p, const double* __restrict__ acc_g, const bool lane_used) {
  p.s0 = 0.0;
  p.s1 = 0.0;
  p.zc = 0;
  p.upd = 0;
  if (lane_used) {
    const double2 mm = *reinterpret_cast<const double2*>(acc_g + 2);
    p.mn = mm.x;
    p.mx = mm.y;
    minmax_bounds(p);
  } else {
    p.mn = 0.0;
    p.mx = 0.0;
    p.lowb = 0u;
    p.span = 0xFFFFFFFFu;
  }
}

// The 16 slots of NR rows of one round, evaluated together: the rows share the lane's offsets and
// weights (read once per slot from the ring slot) and give the FP64 pipe NR x 4 independent
// dependency chains.  `stage0` is the first row's chunk in shared memory, row r is STAGE_BYTES
// further.  The slot loop handles 4 slots per trip and is not unrolled further: measured on the B200,
// every variant with a bigger body (8 slots per trip, a register software pipeline, a deferred fold
// threaded through the loop) was slower than this plain form.  `a` must come from partial_begin.
// TAB64: the transmission kinds use the replicated 64-entry table (K3, lane_off = (lane & 15) << 3)
// instead of the 256-entry one (K2).
template <int FKIND, int NR, bool TAB64 = false>
__device__ __forceinline__ void rows_slots(Partial (&a)[NR], const unsigned char* __restrict__ stage0,
                                           const uint16_t* __restrict__ offp, const double* __restrict__ wp,
                                           const int nv, const TransConst& K, const TransLane& tl,
                                           const bool (&is_log)[NR], const TransConst256& K2,
                                           const unsigned char* __restrict__ tab,
                                           const TransConst64* __restrict__ K3 = nullptr, const unsigned lane_off = 0u,
                                           const int row_stride = STAGE_BYTES,
                                           const TransConstT* __restrict__ KT = nullptr) {
#if defined(ECCKD_V_NOCLAMP) && defined(ECCKD_TAB_LOGN)
  // the transmission kinds of reduce2_kernel: no clamp of the optical depth on the fast path
  constexpr bool NOCLAMP = TAB64 && FKIND == F_TRANS;
#else
  constexpr bool NOCLAMP = false;
#endif
  const unsigned hi_cap = NOCLAMP ? (unsigned)KT->hi_max : 0x7ff00000u;
  constexpr bool LOGKIND = (FKIND == F_LOG || FKIND == F_HYBRID);
  // the loads of one trip (4 slots): plan offsets, weights, optical depths of the NR rows
  auto load_trip = [&](double (&tau)[NR][4], double (&w)[4], const int kk) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned o = offp[(kk + j) * 32];
      w[j] = wp[(kk + j) * 32];  // zero for padding slots
#pragma unroll
      for (int r = 0; r < NR; ++r) tau[r][j] = *reinterpret_cast<const double*>(stage0 + r * row_stride + o);
    }
  };
  auto compute_trip = [&](double (&tau)[NR][4], const double (&w)[4], const int kk) {
    // extremes: one range check per value, one (rare) branch per 4 x NR values
    bool outside = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      unsigned worst = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) worst = max(worst, (unsigned)__double2hiint(tau[r][j]) - a[r].lowb);
      outside = outside || (worst >= a[r].span);
    }
    if (outside) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        double mn = a[r].mn, mx = a[r].mx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mn = (tau[r][j] < mn) ? tau[r][j] : mn;
          mx = (tau[r][j] > mx) ? tau[r][j] : mx;
          // a NaN optical depth always lands here (its high word is outside every range): poison the
          // sum, as the reference's sum() would be, instead of letting the clamps below absorb it
          if (tau[r][j] != tau[r][j]) a[r].s0 = tau[r][j];
        }
        if (mn != a[r].mn || mx != a[r].mx) {
          a[r].mn = mn;
          a[r].mx = mx;
          a[r].upd = 1;
          minmax_bounds(a[r], hi_cap);
        }
        if (NOCLAMP) {  // optical depths beyond 708/kD (exp underflows: f = 1) come through here: clamp them now
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tau[r][j] = __hiloint2double(min(__double2hiint(tau[r][j]), KT->hi_max), __double2loint(tau[r][j]));
        }
      }
    }
    double f[NR][4];
    bool pos[NR][4];
    if (FKIND == F_TRANS) {
      TransMid m[NR][4];
      double T[NR][4];
#if defined(ECCKD_TAB_LOGN)
      // TAB64 callers: the 2^ECCKD_TAB_LOGR-fold replicated 2^ECCKD_TAB_LOGN-entry table; `lane_off` is its
      // shared-window address + the lane's copy offset
#define ECCKD_P1(x) (TAB64 ? (NOCLAMP ? transT_phase1_noclamp(x, *KT) : transT_phase1(x, *KT)) : trans256_phase1(x, K2))
#define ECCKD_P2(mm) (TAB64 ? transT_phase2<ECCKD_TAB_LOGN, ECCKD_TAB_LOGR>(mm, lane_off) : trans256_phase2(mm, tab))
#define ECCKD_P3(mm, TT) (TAB64 ? transT_phase3<ECCKD_TAB_QDEG>(mm, TT, *KT) : trans256_phase3(mm, TT, K2))
#elif defined(ECCKD_V_TAB256)
      // TAB64 callers: `tab` is the 8-fold replicated 256-entry table, lane_off = (lane & 7) << 3
#define ECCKD_P1(x) (TAB64 ? ECCKD_P1_256(x) : trans256_phase1(x, K2))
#ifdef ECCKD_V_I2F
#define ECCKD_P1_256(x) trans256_phase1_i2f(x, K2)
#else
#define ECCKD_P1_256(x) trans256_phase1(x, K2)
#endif
#define ECCKD_P2(mm) (TAB64 ? trans256r_phase2(mm, tab, lane_off) : trans256_phase2(mm, tab))
#define ECCKD_P3(mm, TT) trans256_phase3(mm, TT, K2)
#else
#ifdef ECCKD_V_I2F
#define ECCKD_P1(x) (TAB64 ? trans64_phase1_i2f(x, *K3) : trans256_phase1(x, K2))
#else
#define ECCKD_P1(x) (TAB64 ? trans64_phase1(x, *K3) : trans256_phase1(x, K2))
#endif
#define ECCKD_P2(mm) (TAB64 ? trans64_phase2(mm, tab, lane_off) : trans256_phase2(mm, tab))
#define ECCKD_P3(mm, TT) (TAB64 ? trans64_phase3(mm, TT, *K3) : trans256_phase3(mm, TT, K2))
#endif
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) m[r][j] = ECCKD_P1(tau[r][j]);
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) T[r][j] = ECCKD_P2(m[r][j]);
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[r][j] = ECCKD_P3(m[r][j], T[r][j]);
          pos[r][j] = true;
        }
#undef ECCKD_P1
#undef ECCKD_P2
#undef ECCKD_P3
    } else {
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[r][j] = f_transform<FKIND>(tau[r][j], K, tl, is_log[r], pos[r][j]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double we = w[j];
        if (LOGKIND) {
          we = pos[r][j] ? we : 0.0;
          a[r].zc += (kk + j < nv && !pos[r][j]) ? 1 : 0;
          a[r].s1 += we;
        }
        a[r].s0 = fma(f[r][j], we, a[r].s0);
      }
    }
  };
#ifdef ECCKD_V_TAUPF
  // two register sets in turn: the loads of the next trip are in flight while this one is evaluated
  double tA[NR][4], wA[4], tB[NR][4], wB[4];
  load_trip(tA, wA, 0);
#pragma unroll 1
  for (int kk = 0; kk < SLOTS; kk += 8) {
    load_trip(tB, wB, kk + 4);
    compute_trip(tA, wA, kk);
    if (kk + 8 < SLOTS) load_trip(tA, wA, kk + 8);
    compute_trip(tB, wB, kk + 4);
  }
#elif defined(ECCKD_V_ROT)
  // rotating pipeline: the next trip's loads are issued before this trip is evaluated, one code copy
  // (the register moves at the end of a trip are the price)
  double tau[NR][4], w[4];
  load_trip(tau, w, 0);
#pragma unroll 1
  for (int kk = 0; kk < SLOTS; kk += 4) {
    double tn[NR][4], wn[4];
    load_trip(tn, wn, (kk + 4) & (SLOTS - 1));  // (the last trip re-reads trip 0: unused)
    compute_trip(tau, w, kk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = wn[j];
#pragma unroll
      for (int r = 0; r < NR; ++r) tau[r][j] = tn[r][j];
    }
  }
#else
  // no register software pipeline: the loads of a trip are issued at its top and the other warps
  // of the scheduler cover their latency (a rotating pipeline cost 52 register moves per trip)
#ifdef ECCKD_V_UNROLL4
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int kk = 0; kk < SLOTS; kk += 4) {
    double tau[NR][4], w[4];
    load_trip(tau, w, kk);
    compute_trip(tau, w, kk);
  }
#endif
}

// Segmented suffix reduction over adjacent lanes of the same g point (lane predicates from the
// plan) for NR rows at once, then the head lane of every g point folds the results into the rows'
// accumulators (row r at acc + r*ng*4).  The extremes are only reduced and stored when some lane of
// the warp found a new one in this round (warp-uniform decision).
template <int NR, bool S0, bool S1, bool MM, bool ZC>
__device__ __forceinline__ void fold_partials(Partial (&a)[NR], const unsigned meta, const int nsteps,
                                              double* __restrict__ acc, int32_t* __restrict__ accz, const int ng) {
  if (!S0 && !S1 && !ZC) {  // only the extremes, and only if one changed: nothing to do in most rounds
    int u = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) u |= a[r].upd;
    if (!__any_sync(0xffffffffu, u)) return;
  }
  bool mm_any = false;
  if (MM) {
    int u = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) u |= a[r].upd;
    mm_any = __any_sync(0xffffffffu, u);
  }
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    if (st < nsteps) {  // warp-uniform
      const int d = 1 << st;
      const bool take = (meta >> (17 + st)) & 1u;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (S0) {
          const double t0 = __shfl_down_sync(0xffffffffu, a[r].s0, d);
          if (take) a[r].s0 += t0;
        }
        if (S1) {
          const double t1 = __shfl_down_sync(0xffffffffu, a[r].s1, d);
          if (take) a[r].s1 += t1;
        }
        if (ZC) {
          const int tz = __shfl_down_sync(0xffffffffu, a[r].zc, d);
          if (take) a[r].zc += tz;
        }
      }
    }
  }
  if (MM && mm_any) {
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      if (st < nsteps) {
        const int d = 1 << st;
        const bool take = (meta >> (17 + st)) & 1u;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double tn = __shfl_down_sync(0xffffffffu, a[r].mn, d);
          const double tx = __shfl_down_sync(0xffffffffu, a[r].mx, d);
          a[r].mn = (take && tn < a[r].mn) ? tn : a[r].mn;
          a[r].mx = (take && tx > a[r].mx) ? tx : a[r].mx;
        }
      }
    }
  }
  const unsigned g = meta & 0xFFFu;
  if (g != NO_G && ((meta >> 22) & 1u)) {  // head lane of its g point in this round
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double2* q = reinterpret_cast<double2*>(acc + ((size_t)r * ng + g) * 4);
      if (S0 || S1) {
        double2 a01 = q[0];
        if (S0) a01.x += a[r].s0;
        if (S1) a01.y += a[r].s1;
        q[0] = a01;
      }
      if (MM && mm_any) q[1] = make_double2(a[r].mn, a[r].mx);  // already include the previous extremes
      if (ZC) accz[r * ng + g] += a[r].zc;
    }
  }
}

}  // namespace ecckd
