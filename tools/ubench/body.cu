// Times the consumer-warp round body (rows_slots + fold_partials) of the spectral reduction in
// isolation: W warps per SM loop over a resident shared-memory stage, no copies, no ring.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ecckd_b200/csrc -o body body.cu
#include <cstdio>
#include "reduce_rows.cuh"
using namespace ecckd;

template <int NR, bool FOLD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) body_kernel(TransConst K, int iters, double* out, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  // layout: per warp: 2 stages, one wperm + PlanRound; acc
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem + RING_SLOT_BYTES + (size_t)warp * (NR * STAGE_BYTES + 64 * 4 * 8 * NR);
  double* stage = reinterpret_cast<double*>(base);
  double* wperm = reinterpret_cast<double*>(smem);   // one ring slot shared by all warps (every warp writes the same values)
  PlanRound* pr = reinterpret_cast<PlanRound*>(smem + WPERM_BYTES);
  double* acc = reinterpret_cast<double*>(base + NR * STAGE_BYTES);
  for (int i = lane; i < NR * CH_MAX; i += 32) stage[i] = 0.001 * ((i * 7919) % 1000) + 1e-4;
  for (int k = 0; k < SLOTS; ++k) {
    wperm[k * 32 + lane] = 1.0 + 0.01 * k;
    pr->off[k][lane] = (uint16_t)(((k * 32 + lane) % CH_MAX) * 8);
  }
  // five segments of lanes: 12, 8, 6, 4, 2
  const int seg_start[6] = {0, 12, 20, 26, 30, 32};
  int sgi = 0;
  while (lane >= seg_start[sgi + 1]) ++sgi;
  const int follow = seg_start[sgi + 1] - 1 - lane;
  unsigned mask = 0;
  for (int s = 0; s < 5; ++s)
    if ((1 << s) <= follow) mask |= 1u << s;
  pr->meta[lane] = (unsigned)sgi | (16u << 12) | (mask << 17) | ((lane == seg_start[sgi]) ? (1u << 22) : 0u);
  for (int i = lane; i < 64 * 4 * NR; i += 32) acc[i] = (i & 3) == 2 ? 1e-5 : ((i & 3) == 3 ? 10.0 : 0.0);  // extremes established: the usual state
  __syncwarp();
  const TransLane tl = trans_lane_init();
  __shared__ __align__(16) double s_tab[256];
  trans_tab_init(s_tab);
  TransConst256 K2;
  K2.a = K.a * 8.0; K2.b = K.b / 8.0; K2.hi_max = K.hi_max;
  for (int i = 0; i < 4; ++i) K2.c[i] = K.c[i];
  int32_t accz_dummy[2];
  const bool is_log[NR] = {};
  __syncthreads();
  const long long t0 = clock64();
  double sink = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    Partial a[NR];
    const unsigned meta = pr->meta[lane];
    for (int r = 0; r < NR; ++r) partial_begin(a[r], acc + ((size_t)r * 64 + (meta & 0xFFFu)) * 4, true);
    rows_slots<F_TRANS, NR>(a, reinterpret_cast<const unsigned char*>(stage), &pr->off[0][lane], wperm + lane, 16, K, tl,
                            is_log, K2, reinterpret_cast<const unsigned char*>(s_tab));
    if (FOLD) fold_partials<NR, true, false, true, false>(a, meta, 4, acc, accz_dummy, 64);
    else sink += a[0].s0 + a[NR - 1].mn;
    __syncwarp();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink + acc[lane];
}

template <int NR, bool FOLD, int MAXT>
void run(int warps) {
  TransConst K;
  trans_const_init(&K, 1.66);
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8);
  cudaMalloc(&cyc, 148 * 8);
  const size_t smem = RING_SLOT_BYTES + (size_t)warps * (NR * STAGE_BYTES + 64 * 4 * 8 * NR);
  cudaFuncSetAttribute(body_kernel<NR, FOLD, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  for (int r = 0; r < 2; ++r) {
    body_kernel<NR, FOLD, MAXT><<<148, warps * 32, smem>>>(K, iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double c = (double)h[0] / iters;
  const double elems = (double)warps * NR * SLOTS * 32;
  printf("NR=%d fold=%d maxthreads %4d warps %2d: %8.1f cycles per round -> %5.3f elem/cycle/SM (70%% of the HBM roofline needs 2.04)\n", NR,
         (int)FOLD, MAXT, warps, c, elems / c);
}

int main() {
  for (int w : {8, 12}) run<2, true, 384>(w);
  for (int w : {8, 12, 16}) run<1, true, 512>(w);
  return 0;
}
