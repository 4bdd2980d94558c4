// FP64 pipe throughput on B200 as a function of where DFMA operands come from.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 2048
template <int V>
__global__ void k(double* out, long long* cyc, double a, double b, const double* __restrict__ g) {
  double x[8], y[8], z[8];
  for (int i = 0; i < 8; ++i) {
    x[i] = a + i + threadIdx.x;
    y[i] = g[i + threadIdx.x] ;
    z[i] = g[64 + i + threadIdx.x];
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (V == 0) x[i] = fma(x[i], b, a);            // reg, const, const
      if (V == 1) x[i] = fma(x[i], y[i], a);         // reg, reg, const
      if (V == 2) x[i] = fma(x[i], y[i], z[i]);      // reg, reg, reg
      if (V == 3) x[i] = fma(y[i], z[i], x[i]);      // accumulate form
      if (V == 4) x[i] = fma(x[i], y[(i + 1) & 7], z[(i + 3) & 7]);
      if (V == 5) x[i] = x[i] * y[i];                // DMUL reg reg
      if (V == 6) x[i] = x[i] + y[i];                // DADD reg reg
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V>
void run(const char* name, const double* g) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int w : {4, 12}) {
    for (int r = 0; r < 2; ++r) {
      k<V><<<148, w * 32>>>(out, cyc, 1.0, 0.9999, g);
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / N_IT;
    printf("%-26s warps %2d: %7.2f cycles per 8 ops -> %5.3f warp-instr/cycle/SM\n", name, w, c, w * 8 / c);
  }
}
int main() {
  double* g;
  cudaMalloc(&g, 4096 * 8);
  cudaMemset(g, 0, 4096 * 8);
  run<0>("DFMA reg,const,const", g);
  run<1>("DFMA reg,reg,const", g);
  run<2>("DFMA reg,reg,reg", g);
  run<3>("DFMA acc (y*z+x)", g);
  run<4>("DFMA reg,reg',reg''", g);
  run<5>("DMUL reg,reg", g);
  run<6>("DADD reg,reg", g);
  return 0;
}
