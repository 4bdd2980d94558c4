// lds64_conflicts.cu - which lanes of a warp share a shared-memory wavefront for 64-bit loads?
// Times LDS.64 with fixed per-lane address patterns (one warp, 8 independent loads in flight) and
// prints cycles per load next to the wavefront count predicted by two models:
//   half : the warp is served as lanes 0-15 then 16-31; a half costs max over the 16 8-byte bank
//          pairs of the distinct addresses in that pair
//   full : one group of 32 lanes; cost = max(2, max over bank pairs of distinct addresses)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds64_conflicts lds64_conflicts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <set>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void probe(const int* __restrict__ pat, int npat, long long* __restrict__ cycles, double* sink) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += 32) sm[i] = i;
  __syncwarp();
  for (int p = 0; p < npat; ++p) {
    const int a = pat[p * 32 + threadIdx.x];  // element index < 512
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += sm[a + 512 * u + ((it & 1) << 12)];
    }
    const long long t1 = clock64();
    double s = 0;
    for (int u = 0; u < 8; ++u) s += acc[u];
    sink[p * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[p] = t1 - t0;
  }
}

static int model(const int* a, bool half) {
  auto cost = [&](int lo, int hi) {
    int mx = 0;
    for (int r = 0; r < 16; ++r) {
      std::set<int> s;
      for (int l = lo; l < hi; ++l)
        if ((a[l] & 15) == r) s.insert(a[l]);
      mx = std::max(mx, (int)s.size());
    }
    return mx;
  };
  if (half) return std::max(1, cost(0, 16)) + std::max(1, cost(16, 32));
  return std::max(2, cost(0, 32));
}

int main() {
  std::vector<int> pat;
  auto add = [&](int (*f)(int)) { for (int l = 0; l < 32; ++l) pat.push_back(f(l)); };
  add([](int l) { return l; });                              // consecutive
  add([](int l) { return (l >> 1) + 16 * (l & 1); });        // pairs (l/2), two rows: half model 4, even/odd split 2
  add([](int l) { return (l & 15) + 16 * (l >> 4) * 3; });   // second half repeats the pairs of the first: 2
  add([](int l) { return (l & 7) + 16 * (l >> 3); });        // 8 pairs, 4 rows: 4 under both
  add([](int l) { return l < 16 ? l : 16 * (l - 15); });     // second half all in pair 0: half 17, full 17
  add([](int l) { return (l & 1) ? 16 * (l >> 1) + 1 : (l >> 1); });  // odd lanes conflict in pair 1 (16 rows), even lanes spread
  add([](int l) { return l < 8 ? 16 * l : (l < 16 ? l : (l < 24 ? 16 * (l - 16) + 8 : l - 8)); });
  srand(7);
  for (int p = 0; p < 25; ++p)
    for (int l = 0; l < 32; ++l) pat.push_back(rand() % 480);
  const int npat = (int)pat.size() / 32;
  int* d_pat;
  long long* d_cyc;
  double* d_sink;
  cudaMalloc(&d_pat, pat.size() * 4);
  cudaMalloc(&d_cyc, npat * 8);
  cudaMalloc(&d_sink, npat * 32 * 8);
  cudaMemcpy(d_pat, pat.data(), pat.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8);
  probe<<<1, 32, 8192 * 8>>>(d_pat, npat, d_cyc, d_sink);
  probe<<<1, 32, 8192 * 8>>>(d_pat, npat, d_cyc, d_sink);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  std::vector<long long> cyc(npat);
  cudaMemcpy(cyc.data(), d_cyc, npat * 8, cudaMemcpyDeviceToHost);
  printf("pattern cycles_per_LDS64 half_model full_model\n");
  for (int p = 0; p < npat; ++p)
    printf("%2d %8.3f %3d %3d\n", p, cyc[p] / (256.0 * 8), model(&pat[p * 32], true), model(&pat[p * 32], false));
  return 0;
}
