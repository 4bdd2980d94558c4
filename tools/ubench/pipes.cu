// Which pipe do I2F.F64 / DSETP / 64-bit integer min run on?  Each variant adds work to a loop
// that already saturates the FP64 pipe with 8 DFMA chains; if the added instruction is an FP64-pipe
// instruction the loop gets proportionally slower.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 2048

template <int V>
__global__ void k(double* out, long long* cyc, double a, double b) {
  double x[8];
  long long m[8];
  int acc = 0;
  for (int i = 0; i < 8; ++i) {
    x[i] = a + i + threadIdx.x;
    m[i] = threadIdx.x * 77 + i;
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = fma(x[i], b, a);
      if (V == 1) acc ^= __double2hiint((double)(acc + i + it));                     // I2F.F64
      if (V == 2) acc += (x[i] < a + 0.5) ? 1 : 0;                                      // DSETP (+DADD hoisted)
      if (V == 3) m[i] = min(m[i], (long long)__double_as_longlong(x[i]) ^ (long long)it); // int64 min
      if (V == 4) acc ^= __double2hiint(x[i]) >> 31;                                    // ALU only
      if (V == 5) acc ^= __float2int_rn((float)x[i]);                                   // F2F.F32.F64
    }
  }
  long long t1 = clock64();
  double s = acc;
  for (int i = 0; i < 8; ++i) s += x[i] + (double)m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V>
void run(const char* name) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int r = 0; r < 2; ++r) {
    k<V><<<148, 12 * 32>>>(out, cyc, 1.0, 0.9999);
    cudaDeviceSynchronize();
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s %8.2f cycles per iteration (12 warps x 8 DFMA chains)\n", name, (double)h[0] / N_IT);
}
int main() {
  run<0>("8 DFMA");
  run<1>("8 DFMA + 8 I2F.F64");
  run<2>("8 DFMA + 8 DSETP");
  run<3>("8 DFMA + 8 int64 min");
  run<4>("8 DFMA + 8 ALU");
  run<5>("8 DFMA + 8 F2F.F32.F64+F2I");
  return 0;
}
