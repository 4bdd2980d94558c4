// common.cuh - internals shared by the sm_100a kernels behind include/ecckd_b200.h
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ecckd_b200.h"

#define ECCKD_API extern "C" __attribute__((visibility("default")))

// src/ecckd/constants.h:22-26
#define ECCKD_ACCEL_GRAVITY 9.80665
#define ECCKD_SPECIFIC_HEAT_AIR 1004.0
#define ECCKD_LW_DIFFUSIVITY 1.66
#define ECCKD_MOLAR_MASS_DRY_AIR 28.970

namespace ecckd {

void set_error(const char* fmt, ...);

// A grow-only device scratch buffer owned by the context (avoids cudaMalloc on
// the hot path; all work on one context is stream-ordered so reuse is safe).
struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};

}  // namespace ecckd

struct ecckd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int num_sms = 148;
  size_t smem_optin = 0;
  int64_t launches = 0;
  ecckd::Workspace ws[8];   // device scratch slots (0-3 kernels' descriptors/partials, 4-5 host-API staging)
  bool timing = false;      // record CUDA events around the streaming reduction kernel
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> reduce_events;
  cudaStream_t copy_stream = nullptr;  // host-buffer entry points: H2D copies overlapped with the kernels
  void* pinned = nullptr;   // small pinned staging buffer for descriptor uploads
  size_t pinned_bytes = 0;
  size_t pinned_off = 0;    // bump offset; wraps after a stream synchronise
  // Host-buffer entry points keep the plan of the last g-point map they saw (and a device copy of
  // that map to compare the next call's map with, element for element): the create_look_up_table
  // block loop passes the same map for every block of a gas.
  struct ecckd_gpoint_plan* cached_plan = nullptr;
  int32_t* cached_g = nullptr;
  int64_t cached_nwav = 0;
  int cached_ng = 0;
  int64_t plan_cache_hits = 0, plan_cache_misses = 0;
  double* d_exp2tab = nullptr;  // 2^(j/N) table of the transmission averaging kernels (spectral_reduce.cu)
  int exp2tab_logn = 0;
};

namespace ecckd {

// Returns device scratch slot `slot` with at least `bytes` bytes (256-aligned).
int ws_reserve(ecckd_ctx* ctx, int slot, size_t bytes, void** out);
// Stages `bytes` from pageable host memory through the context's pinned buffer
// and enqueues the H2D copy on the context stream.
int upload_small(ecckd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

}  // namespace ecckd

namespace ecckd {
// Every entry point that takes a context (or an object created on one) runs with that context's
// device current and restores the caller's device on return: one process may hold contexts on
// several GPUs, and workspace allocation, plan/model creation and kernel launches all act on the
// CURRENT device.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const ecckd_ctx* ctx) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
      switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
}  // namespace ecckd
#define ECCKD_DEVICE(ctx) ecckd::DeviceGuard _ecckd_device_guard(ctx)

#define ECCKD_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ecckd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                \
                       cudaGetErrorString(_e));                                          \
      return ECCKD_UNEXPECTED_EXCEPTION;                                                 \
    }                                                                                    \
  } while (0)

#define ECCKD_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ecckd::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define ECCKD_LAUNCH_CHECK(ctx)                  \
  do {                                           \
    (ctx)->launches++;                           \
    ECCKD_CUDA_OK(cudaGetLastError());           \
  } while (0)

#define ECCKD_TRY(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != ECCKD_OK) return _rc; \
  } while (0)

namespace ecckd {
// Sums over the 32 lanes of FOUR values per lane at once (a transposed reduction): the first two
// steps exchange half of the values for the partner's other half, so 6 shuffles do the work of the
// 20 that four separate xor trees take.  Afterwards lane 0 holds the sum of v[0], lane 16 of v[1],
// lane 8 of v[2], lane 24 of v[3] (every lane of the same (bit 4, bit 3) group holds that sum too).
// The order of the additions is fixed: bitwise reproducible.
__device__ __forceinline__ double warp_sum4_transposed(const double (&v)[4], const int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  // step 1 (xor 16): lanes with bit 4 clear keep v[0], v[2]; the others v[1], v[3]
  const double a0 = (hi16 ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, hi16 ? v[0] : v[1], 16);
  const double a1 = (hi16 ? v[3] : v[2]) + __shfl_xor_sync(0xffffffffu, hi16 ? v[2] : v[3], 16);
  // step 2 (xor 8): lanes with bit 3 clear keep a0, the others a1
  double b = (hi8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi8 ? a0 : a1, 8);
  b += __shfl_xor_sync(0xffffffffu, b, 4);
  b += __shfl_xor_sync(0xffffffffu, b, 2);
  b += __shfl_xor_sync(0xffffffffu, b, 1);
  return b;
}
// index (0..3) of the value whose sum warp_sum4_transposed leaves in this lane
__device__ __forceinline__ int warp_sum4_slot(const int lane) { return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1); }
}  // namespace ecckd
