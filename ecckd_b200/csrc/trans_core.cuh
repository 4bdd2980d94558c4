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

// 2^(j/256), j = 0..255, correctly rounded (mpmath, tools/gen_exp2_table.py): the consumer warps of
// the spectral reduction keep it in shared memory and get by with a degree-4 polynomial
// (|r| <= ln2/512: r^5/120 < 4e-17), two FP64 instructions and two shuffles less per value than the
// 32-entry lane table above.
__constant__ double EXP2_256[256] = {
    0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0,
    0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0, 0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0,
    0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,
    0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0,
    0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0, 0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0,
    0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,
    0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0,
    0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0, 0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0,
    0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,
    0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0,
    0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0, 0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0,
    0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,
    0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0,
    0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0, 0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0,
    0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,
    0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0,
    0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0, 0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0,
    0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,
    0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0,
    0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0, 0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0,
    0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,
    0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0,
    0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0, 0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0,
    0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,
    0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0,
    0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0, 0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0,
    0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,
    0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0,
    0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0, 0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0,
    0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,
    0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0,
    0x1.6623882552225p+0, 0x1.671c1c70833f6p+0, 0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,
    0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0,
    0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0, 0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0,
    0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,
    0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0,
    0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0, 0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0,
    0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,
    0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0, 0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0,
    0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,
    0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0,
    0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0, 0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0,
    0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,
    0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0,
    0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0, 0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0,
    0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,
    0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0,
    0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0, 0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,
    0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0,
    0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0,
    0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,
    0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0,
    0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0, 0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0,
    0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,
    0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0, 0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0,
    0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,
    0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0,
    0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0, 0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0,
    0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,
    0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0};

struct TransConst256 {
  double a;        // -kd * 256/ln2
  double b;        // ln2 / (256 kd)
  double c[4];     // c[i] = (-kd)^(i+1) / (i+1)!
  int32_t hi_max;  // high word of 708/kd
  int32_t pad_;
};

inline void trans_const256_init(TransConst256* k, double kd) {
  const long double ln2 = 0.693147180559945309417232121458176568L;
  const long double lkd = (long double)kd;
  k->a = (double)(-lkd * 256.0L / ln2);
  k->b = (double)(ln2 / (256.0L * lkd));
  long double f = 1.0L, pw = 1.0L;
  for (int i = 0; i < 4; ++i) {
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

constexpr int TRANS_TAB_BYTES = 256 * 8;
// fills a 2 KB shared-memory table: entry j = 2^(j/256) with (j << 12) taken off its high word, so
// that adding (n << 12) applies 2^k and cancels j (n = 256 k + j).  Call with all threads of the CTA.
__device__ __forceinline__ void trans_tab_init(double* tab) {
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    const double t = EXP2_256[j];
    tab[j] = __hiloint2double(__double2hiint(t) - (j << 12), __double2loint(t));
  }
}

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

// The three phases with the 256-entry shared-memory table.
__device__ __forceinline__ TransMid trans256_phase1(double tau, const TransConst256& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma(t - magic, K.b, tc);
  return m;
}
__device__ __forceinline__ double trans256_phase2(const TransMid& m, const unsigned char* __restrict__ tab) {
  const uint2 e = *reinterpret_cast<const uint2*>(tab + (__byte_perm((unsigned)m.n, 0u, 0x4440u) << 3));  // (n & 255) * 8
  return __hiloint2double((int)e.y + (m.n << 12), (int)e.x);
}
__device__ __forceinline__ double trans256_phase3(const TransMid& m, const double T, const TransConst256& K) {
  double q = K.c[3];
  q = fma(q, m.rho, K.c[2]);
  q = fma(q, m.rho, K.c[1]);
  q = fma(q, m.rho, K.c[0]);
  return fma(-T, fma(q, m.rho, 1.0), 1.0);
}
__device__ __forceinline__ double trans256_one_minus_exp(double tau, const TransConst256& K,
                                                         const unsigned char* __restrict__ tab) {
  const TransMid m = trans256_phase1(tau, K);
  return trans256_phase3(m, trans256_phase2(m, tab), K);
}

// The same with a 64-entry table REPLICATED 16 times (8 KB of shared memory): copy c = lane & 15 of
// entry j sits at byte ((j*16 + c)*8), so the 16 lanes of a half-warp always read 16 different
// 8-byte bank pairs whatever their j -- the look-up is one conflict-free wavefront per half-warp,
// where the single 256-entry table above cost ~3.  |r| <= ln2/128 needs a degree-5 polynomial
// (r^6/720 < 3.6e-17): one FP64 instruction more per value, one integer instruction less.
struct TransConst64 {
  double a;        // -kd * 64/ln2
  double b;        // ln2 / (64 kd)
  double c[5];     // c[i] = (-kd)^(i+1) / (i+1)!
  int32_t hi_max;  // high word of 708/kd
  int32_t pad_;
};

inline void trans_const64_init(TransConst64* k, double kd) {
  const long double ln2 = 0.693147180559945309417232121458176568L;
  const long double lkd = (long double)kd;
  k->a = (double)(-lkd * 64.0L / ln2);
  k->b = (double)(ln2 / (64.0L * lkd));
  long double f = 1.0L, pw = 1.0L;
  for (int i = 0; i < 5; ++i) {
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

constexpr int TRANS_TAB64_BYTES = 64 * 16 * 8;
// entry (j, c) = 2^(j/64) with (j << 14) taken off its high word: adding (n << 14), n = 64 k + j,
// applies 2^k and cancels j.  Call with all threads of the CTA.
__device__ __forceinline__ void trans_tab64_init(double* tab) {
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
    const int j = i >> 4;
    const double t = EXP2_256[4 * j];
    tab[i] = __hiloint2double(__double2hiint(t) - (j << 14), __double2loint(t));
  }
}
__device__ __forceinline__ TransMid trans64_phase1(double tau, const TransConst64& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma(t - magic, K.b, tc);
  return m;
}
// lane_off = (lane & 15) << 3
__device__ __forceinline__ double trans64_phase2(const TransMid& m, const unsigned char* __restrict__ tab,
                                                 const unsigned lane_off) {
  const unsigned idx = (((unsigned)m.n << 7) & 0x1f80u) | lane_off;
  const uint2 e = *reinterpret_cast<const uint2*>(tab + idx);
  return __hiloint2double((int)e.y + (m.n << 14), (int)e.x);
}
__device__ __forceinline__ double trans64_phase3(const TransMid& m, const double T, const TransConst64& K) {
  double q = K.c[4];
  q = fma(q, m.rho, K.c[3]);
  q = fma(q, m.rho, K.c[2]);
  q = fma(q, m.rho, K.c[1]);
  q = fma(q, m.rho, K.c[0]);
  return fma(-T, fma(q, m.rho, 1.0), 1.0);
}
__device__ __forceinline__ double trans64_one_minus_exp(double tau, const TransConst64& K,
                                                        const unsigned char* __restrict__ tab, const unsigned lane_off) {
  const TransMid m = trans64_phase1(tau, K);
  return trans64_phase3(m, trans64_phase2(m, tab, lane_off), K);
}

// The 256-entry table REPLICATED 8 times (16 KB): copy c = lane & 7 of entry j at byte ((j*8 + c)*8).
// Lanes l and l + 8 of a half-warp share a copy and collide only when their j differ and have the
// same parity (an entry row is 64 bytes, so consecutive j alternate between the two halves of the
// banks): ~1.9 wavefronts per half-warp against 1 for the 64-entry table, in exchange for a
// degree-4 polynomial (one FP64 instruction less per value).
constexpr int TRANS_TAB256R_BYTES = 256 * 8 * 8;
__device__ __forceinline__ void trans_tab256r_init(double* tab) {
  for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) {
    const int j = i >> 3;
    const double t = EXP2_256[j];
    tab[i] = __hiloint2double(__double2hiint(t) - (j << 12), __double2loint(t));
  }
}
// lane_off8 = (lane & 7) << 3
__device__ __forceinline__ double trans256r_phase2(const TransMid& m, const unsigned char* __restrict__ tab,
                                                   const unsigned lane_off8) {
  const unsigned idx = (((unsigned)m.n << 6) & 0x3fc0u) | lane_off8;
  const uint2 e = *reinterpret_cast<const uint2*>(tab + idx);
  return __hiloint2double((int)e.y + (m.n << 12), (int)e.x);
}
#ifdef ECCKD_V_I2F
// range reduction with the integer -> FP64 conversion (XU pipe) in place of the FP64 subtraction
__device__ __forceinline__ TransMid trans256_phase1_i2f(double tau, const TransConst256& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma((double)m.n, K.b, tc);
  return m;
}
__device__ __forceinline__ TransMid trans64_phase1_i2f(double tau, const TransConst64& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma((double)m.n, K.b, tc);
  return m;
}
#endif

// ---------------------------------------------------------------------------------------------
// The general form: a table of N = 2^LOGN entries 2^(j/N) replicated R = 2^LOGR times in shared memory
// (copy c = lane & (R-1) of entry j at byte ((j*R + c)*8)), and exp(x) = 1 + x q(x) on |x| <= ln2/(2N) with
// q the MINIMAX polynomial of (e^x - 1)/x (tools/gen_expm1_minimax.py), one degree lower than the Taylor
// polynomial of the same accuracy.  FP64 instructions per value: 3 (range reduction) + QDEG + 2.
//   N = 128,  R = 16, q degree 3: conflict-free look-ups, |error of exp| <= 1.5e-16
//   N = 1024, R = 2,  q degree 2: ~2.6 wavefronts per half-warp,   <= 1.4e-16
//   N = 2048, R = 1,  q degree 2: ~3.1 wavefronts per half-warp,   <= 8.5e-18
// The table itself (N doubles, correctly rounded on the host in extended precision) lives in global
// memory (ecckd_ctx::d_exp2tab); a CTA adjusts and replicates it into shared memory at start-up.
struct TransConstT {
  double a;        // -kd * N/ln2
  double b;        // ln2 / (N kd)
  double c[4];     // c[i] = q_i (-kd)^(i+1)
  int32_t hi_max;  // high word of 708/kd
  int32_t pad_;
};
static const double EXPM1_Q3_128[4] = {0x1.ffffffffffe08p-1, 0x1.ffffffffffd60p-2, 0x1.55555d887706dp-3, 0x1.55555d8840eb4p-5};
static const double EXPM1_Q2_1024[3] = {0x1.0000000000000p+0, 0x1.0000001ebfbe0p-1, 0x1.5555557e54fd5p-3};
static const double EXPM1_Q2_2048[3] = {0x1.0000000000000p+0, 0x1.00000007afef8p-1, 0x1.5555555f953f5p-3};
__constant__ double EXPM1_Q2_2048_D[3] = {0x1.0000000000000p+0, 0x1.00000007afef8p-1, 0x1.5555555f953f5p-3};
inline void trans_constT_init(TransConstT* k, double kd, int logn, int qdeg) {
  const long double ln2 = 0.693147180559945309417232121458176568L;
  const long double lkd = (long double)kd;
  const long double N = (long double)(1 << logn);
  const double* q = logn == 7 ? EXPM1_Q3_128 : (logn == 10 ? EXPM1_Q2_1024 : EXPM1_Q2_2048);
  k->a = (double)(-lkd * N / ln2);
  k->b = (double)(ln2 / (N * lkd));
  long double pw = 1.0L;
  for (int i = 0; i < 4; ++i) {
    pw *= -lkd;
    k->c[i] = i <= qdeg ? (double)((long double)q[i] * pw) : 0.0;
  }
  const double tmax = 708.0 / kd;
  uint64_t bits;
  memcpy(&bits, &tmax, 8);
  k->hi_max = (int32_t)(bits >> 32);
  k->pad_ = 0;
}
template <int LOGN, int LOGR>
__device__ __forceinline__ void trans_tabT_init(double* tab, const double* __restrict__ g_exp2) {
  for (int i = threadIdx.x; i < (1 << (LOGN + LOGR)); i += blockDim.x) {
    const int j = i >> LOGR;
    const double t = g_exp2[j];
    tab[i] = __hiloint2double(__double2hiint(t) - (j << (20 - LOGN)), __double2loint(t));
  }
}
__device__ __forceinline__ TransMid transT_phase1(double tau, const TransConstT& K) {
  const double magic = 6755399441055744.0;
  const double tc = __hiloint2double(min(__double2hiint(tau), K.hi_max), __double2loint(tau));
  const double t = fma(tc, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma(t - magic, K.b, tc);
  return m;
}
// (the caller guarantees tau < 708/kd * (1 + 2^-20))
__device__ __forceinline__ TransMid transT_phase1_noclamp(double tau, const TransConstT& K) {
  const double magic = 6755399441055744.0;
  const double t = fma(tau, K.a, magic);
  TransMid m;
  m.n = __double2loint(t);
  m.rho = fma(t - magic, K.b, tau);
  return m;
}
// tab_lane = shared-window address of the table + ((lane & (R-1)) << 3).  Mask, then scale-and-add
// (LOP3 + IMAD/LEA) and a load through the shared window: written in PTX because the compiler turns the
// C form into shift, mask and a separate add of the base.
template <int LOGN, int LOGR>
__device__ __forceinline__ double transT_phase2(const TransMid& m, const unsigned tab_lane) {
  unsigned addr, lo, hi;
  asm("{\n\t.reg .u32 j;\n\tand.b32 j, %1, %2;\n\tmad.lo.u32 %0, j, %3, %4;\n\t}\n"
      : "=r"(addr)
      : "r"(m.n), "n"((1 << LOGN) - 1), "n"(8 << LOGR), "r"(tab_lane));
  asm("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(lo), "=r"(hi) : "r"(addr));
  return __hiloint2double((int)hi + (m.n << (20 - LOGN)), (int)lo);
}
template <int QDEG>
__device__ __forceinline__ double transT_phase3(const TransMid& m, const double T, const TransConstT& K) {
  double q = K.c[QDEG];
#pragma unroll
  for (int i = QDEG - 1; i >= 0; --i) q = fma(q, m.rho, K.c[i]);
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
