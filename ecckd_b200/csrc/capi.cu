// capi.cu - context, memory and error plumbing of the C ABI (include/ecckd_b200.h)
#include <cstdarg>

#include <algorithm>

#include "common.cuh"

namespace ecckd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ws_reserve(ecckd_ctx* ctx, int slot, size_t bytes, void** out) {
  Workspace& w = ctx->ws[slot];
  if (w.bytes < bytes) {
    // All users are stream-ordered on ctx->stream; drain before freeing.
    if (w.ptr) {
      ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      ECCKD_CUDA_OK(cudaFree(w.ptr));
      w.ptr = nullptr;
      w.bytes = 0;
    }
    size_t want = (bytes + 255) & ~size_t(255);
    ECCKD_CUDA_OK(cudaMalloc(&w.ptr, want));
    w.bytes = want;
  }
  *out = w.ptr;
  return ECCKD_OK;
}

__global__ void stage_copy_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
    const size_t n16 = bytes >> 4;
    for (size_t k = i; k < n16; k += stride) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(src)[k];
    for (size_t k = (n16 << 4) + i; k < bytes; k += stride) dst[k] = src[k];
  } else {
    for (size_t k = i; k < bytes; k += stride) dst[k] = src[k];
  }
}

int upload_small(ecckd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return ECCKD_OK;
  if (bytes > ctx->pinned_bytes) {
    // too big for the staging ring: synchronous pageable copy
    ECCKD_CUDA_OK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return ECCKD_OK;
  }
  size_t off = (ctx->pinned_off + 63) & ~size_t(63);
  if (off + bytes > ctx->pinned_bytes) {
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    off = 0;
  }
  char* stage = static_cast<char*>(ctx->pinned) + off;
  memcpy(stage, h_src, bytes);
  ctx->pinned_off = off + bytes;
  // A kernel reads the pinned staging buffer over the bus (it is mapped into the device address space)
  // instead of a host-to-device copy: copies are served by the copy engine in its own order, and a
  // descriptor upload queued behind gigabytes of spectra (the host-buffer averaging call streams them
  // on a second stream) held the kernels that needed it back until the last of them had landed.
  const unsigned threads = 256, blocks = (unsigned)std::min<size_t>(64, (bytes + threads * 16 - 1) / (threads * 16));
  stage_copy_kernel<<<blocks, threads, 0, ctx->stream>>>(static_cast<unsigned char*>(d_dst),
                                                        reinterpret_cast<const unsigned char*>(stage), bytes);
  ECCKD_CUDA_OK(cudaGetLastError());
  return ECCKD_OK;
}

}  // namespace ecckd

using namespace ecckd;

ECCKD_API int ecckd_version(void) { return ECCKD_B200_VERSION; }

ECCKD_API const char* ecckd_last_error(void) { return g_err; }

ECCKD_API int ecckd_averaging_method_from_string(const char* name) {
  // strings of src/ecckd/average_optical_depth.cpp:43-129
  static const char* names[] = {"linear",          "transmission", "transmission-2",
                                "transmission-3",  "transmission-10", "square-root",
                                "logarithmic",     "hybrid-logarithmic-transmission-3",
                                "total-transmission"};  // find_g_points.cpp:146 (interval fits only)
  if (!name) return -1;
  for (int i = 0; i < 9; ++i) {
    if (strcmp(name, names[i]) == 0) return i;
  }
  return -1;
}

ECCKD_API int ecckd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int usable = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++usable;
  }
  return usable;
}

ECCKD_API int ecckd_ctx_create(int device, void* stream, ecckd_ctx** out) {
  ECCKD_REQUIRE(out != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("ecckd_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  ECCKD_REQUIRE(device >= 0 && device < n, ECCKD_PARAMETER_ERROR, "ecckd_ctx_create: device %d out of range [0,%d)",
                device, n);
  cudaDeviceProp prop;
  ECCKD_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  ECCKD_REQUIRE(prop.major == 10, ECCKD_UNEXPECTED_EXCEPTION,
                "ecckd_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  // the caller's current device is restored on return (ECCKD_DEVICE does the same in every other
  // entry point): creating a second context must not redirect the first one's allocations
  ecckd_ctx probe;
  probe.device = device;
  ECCKD_DEVICE(&probe);
  {
    int cur = -1;
    ECCKD_REQUIRE(cudaGetDevice(&cur) == cudaSuccess && cur == device, ECCKD_UNEXPECTED_EXCEPTION,
                  "ecckd_ctx_create: cannot select device %d", device);
  }
  ecckd_ctx* ctx = new ecckd_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (stream) {
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->owns_stream = false;
  } else {
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
      delete ctx;
      set_error("cudaStreamCreateWithFlags failed: %s", cudaGetErrorString(se));
      return ECCKD_UNEXPECTED_EXCEPTION;
    }
    ctx->owns_stream = true;
  }
  {
    // plans and partial sums come from the stream-ordered pool (cudaMallocAsync): keep what it has
    // instead of returning it to the driver at every synchronisation (the default threshold of 0
    // made one host-buffer call in ten pay ~300 ms for a fresh allocation).  This is a property of
    // the DEVICE's default pool, i.e. visible to other users of cudaMallocAsync in the process
    // (documented in include/ecckd_b200.h); ECCKD_KEEP_POOL=0 leaves the pool alone.
    cudaMemPool_t pool = nullptr;
    const char* keep_env = getenv("ECCKD_KEEP_POOL");
    if (keep_env && atoi(keep_env) == 0) {
    } else
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      uint64_t keep = ~uint64_t(0);
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  ctx->pinned_bytes = size_t(4) << 20;
  cudaError_t pe = cudaHostAlloc(&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault);
  if (pe != cudaSuccess) {
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    set_error("cudaHostAlloc failed: %s", cudaGetErrorString(pe));
    return ECCKD_UNEXPECTED_EXCEPTION;
  }
  *out = ctx;
  return ECCKD_OK;
}

ECCKD_API int ecckd_ctx_destroy(ecckd_ctx* ctx) {
  if (!ctx) return ECCKD_OK;
  ECCKD_DEVICE(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->cached_plan) ecckd_gpoint_plan_destroy(ctx->cached_plan);
  if (ctx->cached_g) cudaFree(ctx->cached_g);
  if (ctx->d_exp2tab) cudaFree(ctx->d_exp2tab);
  for (auto& w : ctx->ws) {
    if (w.ptr) cudaFree(w.ptr);
  }
  for (auto& ev : ctx->reduce_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ECCKD_OK;
}

ECCKD_API int ecckd_ctx_set_stream(ecckd_ctx* ctx, void* stream) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_ctx_set_stream: ctx is NULL");
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (ctx->owns_stream) {
    cudaStreamDestroy(ctx->stream);
    ctx->owns_stream = false;
  }
  if (stream) {
    ctx->stream = static_cast<cudaStream_t>(stream);
  } else {
    ECCKD_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_synchronize(ecckd_ctx* ctx) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_synchronize: ctx is NULL");
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return ECCKD_OK;
}

ECCKD_API int ecckd_device_malloc(ecckd_ctx* ctx, size_t bytes, void** d_ptr) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && d_ptr, ECCKD_PARAMETER_ERROR, "ecckd_device_malloc: NULL argument");
  ECCKD_CUDA_OK(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return ECCKD_OK;
}

ECCKD_API int ecckd_device_free(ecckd_ctx* ctx, void* d_ptr) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_device_free: ctx is NULL");
  if (d_ptr) {
    ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ECCKD_CUDA_OK(cudaFree(d_ptr));
  }
  return ECCKD_OK;
}

ECCKD_API int ecckd_host_malloc_pinned(size_t bytes, void** h_ptr) {
  ECCKD_REQUIRE(h_ptr != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_host_malloc_pinned: NULL argument");
  ECCKD_CUDA_OK(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return ECCKD_OK;
}

ECCKD_API int ecckd_host_free_pinned(void* h_ptr) {
  if (h_ptr) ECCKD_CUDA_OK(cudaFreeHost(h_ptr));
  return ECCKD_OK;
}

ECCKD_API int ecckd_memcpy_h2d(ecckd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_memcpy_h2d: ctx is NULL");
  ECCKD_CUDA_OK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ECCKD_OK;
}

ECCKD_API int ecckd_memcpy_d2h(ecckd_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_memcpy_d2h: ctx is NULL");
  ECCKD_CUDA_OK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return ECCKD_OK;
}

ECCKD_API int64_t ecckd_kernel_launch_count(ecckd_ctx* ctx) { return ctx ? ctx->launches : 0; }

ECCKD_API int ecckd_ctx_plan_cache_stats(ecckd_ctx* ctx, int64_t* hits, int64_t* misses) {
  ECCKD_REQUIRE(ctx, ECCKD_PARAMETER_ERROR, "ecckd_ctx_plan_cache_stats: NULL context");
  if (hits) *hits = ctx->plan_cache_hits;
  if (misses) *misses = ctx->plan_cache_misses;
  return ECCKD_OK;
}

ECCKD_API int ecckd_ctx_set_timing(ecckd_ctx* ctx, int on) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx != nullptr, ECCKD_PARAMETER_ERROR, "ecckd_ctx_set_timing: ctx is NULL");
  ctx->timing = on != 0;
  return ECCKD_OK;
}

ECCKD_API int ecckd_ctx_reduce_kernel_time(ecckd_ctx* ctx, double* total_ms, int64_t* launches) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && total_ms && launches, ECCKD_PARAMETER_ERROR, "ecckd_ctx_reduce_kernel_time: NULL argument");
  ECCKD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  double sum = 0.0;
  for (auto& ev : ctx->reduce_events) {
    float ms = 0.f;
    ECCKD_CUDA_OK(cudaEventElapsedTime(&ms, ev.first, ev.second));
    sum += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *total_ms = sum;
  *launches = (int64_t)ctx->reduce_events.size();
  ctx->reduce_events.clear();
  return ECCKD_OK;
}
