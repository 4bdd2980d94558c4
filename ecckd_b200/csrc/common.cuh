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
