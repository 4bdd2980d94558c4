// Micro-benchmarks for B200 (sm_100a): dependent-issue latency and per-SM throughput of the
// instructions the spectral reduction leans on.  nvcc -arch=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP, int CHAINS>
__global__ void lat_kernel(double* out, long long* cyc, double a, double b) {
  double x[CHAINS];
  for (int i = 0; i < CHAINS; ++i) x[i] = a + i + threadIdx.x;
  int lane = threadIdx.x & 31;
  __shared__ double sm[1024];
  sm[threadIdx.x % 1024] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) x[i] = fma(x[i], b, a);                                  // DFMA
      if (OP == 1) x[i] = x[i] + b;                                         // DADD
      if (OP == 3) x[i] = __shfl_sync(0xffffffffu, x[i], (lane + 1) & 31);  // 2 SHFL
      if (OP == 4) x[i] = (x[i] < b) ? x[i] : b;                            // min: DSETP + 2 FSEL
      if (OP == 5) x[i] = sm[(__double2loint(x[i]) & 1023)];                // LDS.64 dependent
      if (OP == 6) {                                                        // int64 min
        long long v = __double_as_longlong(x[i]), w = __double_as_longlong(b);
        x[i] = __longlong_as_double(v < w ? v : w + it);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, int warps) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int r = 0; r < 2; ++r) {
    lat_kernel<OP, CHAINS><<<148, warps * 32>>>(out, cyc, 1.0, 0.9999);
    cudaDeviceSynchronize();
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0] / N_IT;
  printf("%-18s warps %2d chains %d: %7.2f cycles/iter -> %6.2f cycles per op, %6.3f warp-ops/cycle/SM\n", name, warps, CHAINS,
         c, c / CHAINS, warps * CHAINS / c);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0, 1>("DFMA", 1); run<0, 2>("DFMA", 1); run<0, 4>("DFMA", 1); run<0, 8>("DFMA", 1);
  run<0, 8>("DFMA", 4); run<0, 8>("DFMA", 8); run<0, 8>("DFMA", 12); run<0, 4>("DFMA", 12); run<0, 2>("DFMA", 12); run<0, 1>("DFMA", 12);
  run<1, 1>("DADD", 1); run<1, 8>("DADD", 12);
  run<4, 1>("min DSETP+2FSEL", 1); run<4, 8>("min DSETP+2FSEL", 1); run<4, 8>("min DSETP+2FSEL", 12);
  run<6, 1>("int64 min", 1); run<6, 8>("int64 min", 12);
  run<3, 1>("SHFL x2", 1); run<3, 8>("SHFL x2", 12);
  run<5, 1>("LDS.64 dep", 1); run<5, 8>("LDS.64 dep", 12);
  return 0;
}
