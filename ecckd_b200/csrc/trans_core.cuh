// trans_core.cuh - 1 - exp(-kd*tau) for the transmission averaging methods
// (average_optical_depth.cpp:47-70: `1.0 - exp(-LW_DIFFUSIVITY*OD_SCALING*k * optical_depth)`),
// written for the FP64 pipe of sm_100a: 11 FP64 instructions per value instead of ~18.
//
//   x = -kd*tau = (32 k + j) ln2/32 + r,  |r| <= ln2/64
//   exp(x) = 2^k * 2^(j/32) * (1 + p(r)),  p(r) = r + r^2/2 + ... + r^6/720   (|r^7/7!| < 4e-18)
//
// * n = 32 k + j comes out of one FMA with the 1.5*2^52 magic constant; the reduction is done in
//   units of tau (rho = tau + n*ln2/(32 kd), r = -kd*rho) with the polynomial coefficients
//   pre-multiplied by (-kd)^i on the host, so the multiplication by -kd costs nothing;
// * 2^(j/32) is a 32-entry table held one entry per LANE and looked up with two warp shuffles
//   (a shared-memory table would bank-conflict ~3-way); the table's high words are stored minus
//   (j << 15) so that adding (n << 15) to them applies 2^k and cancels j in one integer IMAD;
// * 1 - exp(x) = 1 - T*(1 + p) with T = 2^k 2^(j/32), two FMAs: absolute error below 2.3e-16 (two
//   ulp of 1; the reference's 1.0 - exp() is within ~1.7e-16), swept against expm1 on the GPU by
//   tests/test_trans_core_gpu.py through the C ABI hook ecckd_debug_eval.
// * tau is clamped to 708/kd through its high word (one integer min): exp() would be below 1e-307
//   there and 1 - exp() is exactly 1 long before.  Not handled (unphysical): tau < -708/kd and NaN.
// All 32 lanes of the warp must call trans_one_minus_exp together (it shuffles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ecckd {

struct TransConst {
  double a;        // -kd * 32/ln2
  double b;        // ln2 / (32 kd)
  double c[6];     // c[i] = (-kd)^(i+1) / (i+1)!
  int32_t hi_max;  // high word of 708/kd
  int32_t pad_;
};

// 2^(j/32), j = 0..31, correctly rounded (generated with mpmath, see tools/gen_exp2_table.py)
__constant__ double EXP2_32[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

inline void trans_const_init(TransConst* k, double kd) {
  const long double ln2 = 0.693147180559945309417232121458176568L;
  const long double lkd = (long double)kd;
  k->a = (double)(-lkd * 32.0L / ln2);
  k->b = (double)(ln2 / (32.0L * lkd));
  long double f = 1.0L, pw = 1.0L;
  for (int i = 0; i < 6; ++i) {
    pw *= -lkd;
    f *= (long double)(i + 1);
    k->c[i] = (double)(pw / f);
  }
  const double tmax = 708.0 / kd;
  uint64_t bits;
  memcpy(&bits, &tmax, 8);
  k->hi_max = (int32_t)(bits >> 32);
  k->pad_ = 0;
}

// per-lane table entry: {low word, high word - (lane << 15)} of 2^(lane/32)
struct TransLane {
  int lo, hi_adj;
};
__device__ __forceinline__ TransLane trans_lane_init() {
  const int lane = threadIdx.x & 31;
  const double t = EXP2_32[lane];
  TransLane tl;
  tl.lo = __double2loint(t);
  tl.hi_adj = __double2hiint(t) - (lane << 15);
  return tl;
}

__device__ __forceinline__ double trans_one_minus_exp(double tau, const TransConst& K, const TransLane& tl) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  const int n = __double2loint(t);
  const double nf = t - magic;
  const double rho = fma(nf, K.b, tc);
  double q = K.c[5];
  q = fma(q, rho, K.c[4]);
  q = fma(q, rho, K.c[3]);
  q = fma(q, rho, K.c[2]);
  q = fma(q, rho, K.c[1]);
  q = fma(q, rho, K.c[0]);
  const double p = q * rho;
  const int thi = __shfl_sync(0xffffffffu, tl.hi_adj, n);
  const int tlo = __shfl_sync(0xffffffffu, tl.lo, n);
  const double T = __hiloint2double(thi + (n << 15), tlo);
  const double om = 1.0 - T;
  return fma(-T, p, om);
}

// The same in three phases, so that a caller evaluating several values can run phase 1 for all
// of them, then all the shuffles (one convergence check), then phase 3: with the single-call form
// ptxas serialises the first value's polynomial in front of the shuffle block.
struct TransMid {
  double rho;
  int n;
};
__device__ __forceinline__ TransMid trans_phase1(double tau, const TransConst& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma(t - magic, K.b, tc);
  return m;
}
__device__ __forceinline__ double trans_phase2(const TransMid& m, const TransLane& tl) {
#ifdef ECCKD_UBENCH_NOSHFL  // timing experiment only (tools/ubench): no table look-up
  return __hiloint2double(tl.hi_adj + (m.n << 15), tl.lo);
#endif
  const int thi = __shfl_sync(0xffffffffu, tl.hi_adj, m.n);
  const int tlo = __shfl_sync(0xffffffffu, tl.lo, m.n);
  return __hiloint2double(thi + (m.n << 15), tlo);
}
__device__ __forceinline__ double trans_phase3(const TransMid& m, const double T, const TransConst& K) {
  double q = K.c[5];
  q = fma(q, m.rho, K.c[4]);
  q = fma(q, m.rho, K.c[3]);
  q = fma(q, m.rho, K.c[2]);
  q = fma(q, m.rho, K.c[1]);
  q = fma(q, m.rho, K.c[0]);
  // 1 - T*(1 + p): two FMAs; absolute error ~1e-16 like the reference's 1 - exp()
  return fma(-T, fma(q, m.rho, 1.0), 1.0);
}

// exp(x) for 0 <= x < 708 with the same table: x = (32 k + j) ln2/32 + r (two-constant
// Cody-Waite reduction), exp(x) = 2^k 2^(j/32) (1 + p(r)).  Below 1 ulp; used for the Planck weights.
__device__ __forceinline__ double exp_table(double x, const TransLane& tl) {
  const double magic = 6755399441055744.0;
  const double t = fma(x, 0x1.71547652b82fep+5, magic);  // 32/ln2
  const int n = __double2loint(t);
  const double nf = t - magic;
  double r = fma(nf, -0x1.62e42fee00000p-6, x);  // -ln2/32, high part (21 trailing zero bits)
  r = fma(nf, -0x1.a39ef35793c76p-38, r);  // low part
  double q = 1.0 / 720.0;
  q = fma(q, r, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double pr = q * r;
  const int thi = __shfl_sync(0xffffffffu, tl.hi_adj, n);
  const int tlo = __shfl_sync(0xffffffffu, tl.lo, n);
  const double T = __hiloint2double(thi + (n << 15), tlo);
  return fma(T, pr, T);
}

}  // namespace ecckd
