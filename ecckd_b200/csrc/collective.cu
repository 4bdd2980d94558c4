// Cross-GPU exchange of the spectral reduction (BASELINE north_star: "NCCL over
// NVLink only for the small per-g-point partial accumulators").
//
// The wavenumber axis is cut into nseg_total SEGMENTS whose boundaries do not depend on the number of
// ranks; a rank owns a contiguous run of segments and reduces each with its own g-point plan
// (ecckd_average_optical_depth_partial_dev into that segment's slot).  ecckd_average_optical_depth_exchange_dev
// then makes every slot visible on every rank (one grouped NCCL broadcast per owner, in place) and combines the
// slots IN SEGMENT ORDER.  A segment's partials depend only on the segment's data, the combination order is
// fixed, so the result is bitwise identical on every rank and for every number of ranks - unlike a ring or
// tree all-reduce, whose summation order follows the communicator.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded in the process when the host
// is PyTorch, else the system one), so that libecckd_b200.so has no link-time dependency on it and
// single-GPU users never load it.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "common.cuh"

using namespace ecckd;

namespace {

struct NcclId {
  char bytes[128];
};
using nccl_comm_t = void*;
enum { kNcclInt64 = 4, kNcclFloat64 = 8 };  // ncclDataType_t values (nccl.h)

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib) return ECCKD_OK;
  const char* names[] = {getenv("ECCKD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  }
  ECCKD_REQUIRE(lib, ECCKD_PROCESSING_ERROR, "NCCL not found (libnccl.so.2; set ECCKD_NCCL_LIB): %s", dlerror());
  NcclApi api;
  api.lib = lib;
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(lib, name);
    ok = ok && p != nullptr;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  ECCKD_REQUIRE(ok, ECCKD_PROCESSING_ERROR, "the NCCL library lacks a required symbol");
  g_nccl = api;
  return ECCKD_OK;
}

#define ECCKD_NCCL_OK(expr)                                                                             \
  do {                                                                                                  \
    const int _r = (expr);                                                                              \
    if (_r != 0) {                                                                                      \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");       \
      return ECCKD_UNEXPECTED_EXCEPTION;                                                                \
    }                                                                                                   \
  } while (0)

// out[i] = ((seg0[i] + seg1[i]) + seg2[i]) + ...   (sums, counts), min over segments (minmax)
__global__ void combine_segments_kernel(int nseg, int64_t n_sums, const double* __restrict__ seg_sums,
                                        double* __restrict__ sums, int64_t n_minmax,
                                        const double* __restrict__ seg_minmax, double* __restrict__ minmax,
                                        int64_t n_count, const long long* __restrict__ seg_count,
                                        long long* __restrict__ count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n_sums; i += stride) {
    double a = seg_sums[i];
    for (int s = 1; s < nseg; ++s) a += seg_sums[(int64_t)s * n_sums + i];
    sums[i] = a;
  }
  for (int64_t i = t; i < n_minmax; i += stride) {
    double a = seg_minmax[i];
    for (int s = 1; s < nseg; ++s) a = fmin(a, seg_minmax[(int64_t)s * n_minmax + i]);
    minmax[i] = a;
  }
  for (int64_t i = t; i < n_count; i += stride) {
    long long a = seg_count[i];
    for (int s = 1; s < nseg; ++s) a += seg_count[(int64_t)s * n_count + i];
    count[i] = a;
  }
}

}  // namespace

struct ecckd_comm {
  ecckd_ctx* ctx = nullptr;
  nccl_comm_t comm = nullptr;
  int nranks = 1, rank = 0;
  bool owned = false;
};

ECCKD_API int ecckd_comm_unique_id(void* id128) {
  ECCKD_REQUIRE(id128, ECCKD_PARAMETER_ERROR, "ecckd_comm_unique_id: NULL argument");
  ECCKD_TRY(nccl_load());
  NcclId id;
  ECCKD_NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.bytes, sizeof(id.bytes));
  return ECCKD_OK;
}

ECCKD_API int ecckd_comm_create(ecckd_ctx* ctx, int nranks, int rank, const void* id128, ecckd_comm** out) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, ECCKD_PARAMETER_ERROR,
                "ecckd_comm_create: bad argument");
  *out = nullptr;
  ECCKD_TRY(nccl_load());
  NcclId id;
  memcpy(id.bytes, id128, sizeof(id.bytes));
  nccl_comm_t c = nullptr;
  ECCKD_NCCL_OK(g_nccl.CommInitRank(&c, nranks, id, rank));
  ecckd_comm* h = new ecckd_comm;
  h->ctx = ctx;
  h->comm = c;
  h->nranks = nranks;
  h->rank = rank;
  h->owned = true;
  *out = h;
  return ECCKD_OK;
}

ECCKD_API int ecckd_comm_from_nccl(ecckd_ctx* ctx, void* nccl_comm, int nranks, int rank, ecckd_comm** out) {
  ECCKD_REQUIRE(ctx && nccl_comm && out && nranks >= 1 && rank >= 0 && rank < nranks, ECCKD_PARAMETER_ERROR,
                "ecckd_comm_from_nccl: bad argument");
  *out = nullptr;
  ECCKD_TRY(nccl_load());
  ecckd_comm* h = new ecckd_comm;
  h->ctx = ctx;
  h->comm = nccl_comm;
  h->nranks = nranks;
  h->rank = rank;
  h->owned = false;
  *out = h;
  return ECCKD_OK;
}

ECCKD_API int ecckd_comm_destroy(ecckd_comm* comm) {
  if (!comm) return ECCKD_OK;
  ECCKD_DEVICE(comm->ctx);
  if (comm->owned && comm->comm) {
    cudaStreamSynchronize(comm->ctx->stream);
    g_nccl.CommDestroy(comm->comm);
  }
  delete comm;
  return ECCKD_OK;
}

ECCKD_API int ecckd_average_optical_depth_exchange_dev(ecckd_ctx* ctx, ecckd_comm* comm, int nseg_total,
                                                       const int32_t* h_seg_begin, int64_t n_sums, int64_t n_minmax,
                                                       int64_t n_count, double* d_seg_sums, double* d_seg_minmax,
                                                       int64_t* d_seg_count, double* d_sums, double* d_minmax,
                                                       int64_t* d_count) {
  ECCKD_DEVICE(ctx);
  ECCKD_REQUIRE(ctx && nseg_total >= 1 && n_sums >= 0 && n_minmax >= 0 && n_count >= 0, ECCKD_PARAMETER_ERROR,
                "ecckd_average_optical_depth_exchange: bad dimensions");
  ECCKD_REQUIRE((n_sums == 0 || (d_seg_sums && d_sums)) && (n_minmax == 0 || (d_seg_minmax && d_minmax)) &&
                    (n_count == 0 || (d_seg_count && d_count)),
                ECCKD_PARAMETER_ERROR, "ecckd_average_optical_depth_exchange: NULL array");
  if (comm && comm->nranks > 1) {
    ECCKD_REQUIRE(comm->ctx == ctx, ECCKD_PARAMETER_ERROR,
                  "ecckd_average_optical_depth_exchange: the communicator belongs to another context");
    ECCKD_REQUIRE(h_seg_begin && h_seg_begin[0] == 0 && h_seg_begin[comm->nranks] == nseg_total, ECCKD_PARAMETER_ERROR,
                  "ecckd_average_optical_depth_exchange: h_seg_begin must run from 0 to nseg_total over the ranks");
    for (int r = 0; r < comm->nranks; ++r)
      ECCKD_REQUIRE(h_seg_begin[r] <= h_seg_begin[r + 1], ECCKD_PARAMETER_ERROR,
                    "ecckd_average_optical_depth_exchange: h_seg_begin must not decrease");
    // every owner broadcasts its run of slots in place; one group = one fused NCCL operation
    ECCKD_NCCL_OK(g_nccl.GroupStart());
    int rc = 0;
    for (int r = 0; r < comm->nranks && rc == 0; ++r) {
      const int64_t s0 = h_seg_begin[r], ns = h_seg_begin[r + 1] - s0;
      if (ns == 0) continue;
      if (n_sums && rc == 0)
        rc = g_nccl.Broadcast(d_seg_sums + s0 * n_sums, d_seg_sums + s0 * n_sums, (size_t)(ns * n_sums), kNcclFloat64, r,
                              comm->comm, ctx->stream);
      if (n_minmax && rc == 0)
        rc = g_nccl.Broadcast(d_seg_minmax + s0 * n_minmax, d_seg_minmax + s0 * n_minmax, (size_t)(ns * n_minmax),
                              kNcclFloat64, r, comm->comm, ctx->stream);
      if (n_count && rc == 0)
        rc = g_nccl.Broadcast(d_seg_count + s0 * n_count, d_seg_count + s0 * n_count, (size_t)(ns * n_count), kNcclInt64, r,
                              comm->comm, ctx->stream);
    }
    const int rc_end = g_nccl.GroupEnd();
    ECCKD_NCCL_OK(rc);
    ECCKD_NCCL_OK(rc_end);
  }
  const int64_t n = std::max(n_sums, std::max(n_minmax, n_count));
  if (n > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8);
    combine_segments_kernel<<<grid, 256, 0, ctx->stream>>>(nseg_total, n_sums, d_seg_sums, d_sums, n_minmax, d_seg_minmax,
                                                           d_minmax, n_count, reinterpret_cast<const long long*>(d_seg_count),
                                                           reinterpret_cast<long long*>(d_count));
    ECCKD_LAUNCH_CHECK(ctx);
  }
  return ECCKD_OK;
}
