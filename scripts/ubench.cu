// scripts/ubench.cu -- dependent-chain latencies on the target GPU (one warp), used to reason about the
// serial critical path of the Gauss-Newton kernels.  nvcc -arch=sm_100a -O3 -fmad=false -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, float fa) {
  double x = a;
  float f = fa;
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, b, a);                       // DFMA
    if (OP == 1) x = (double)(float)x + a;                // F2F.F32.F64 + F2F.F64.F32 + DADD
    if (OP == 2) f = fmaf(f, fa, fa);                    // FFMA
    if (OP == 3) x = (double)__frcp_rn((float)x) + a;    // cvt + MUFU + cvt + DADD
    if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1) + a;   // 2 SHFL + DADD
    if (OP == 5) { idx = (int)sm[idx & 31] & 31; }        // LDS.64 dependent + cvt
    if (OP == 6) x = x / b;                               // DDIV
    if (OP == 7) f = (float)(int)f + fa;                  // F2I + I2F + FADD
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  out[threadIdx.x] = x + f + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"DFMA", "F2F.f32.f64+F2F.f64.f32+DADD", "FFMA", "cvt+MUFU.RCP+cvt+DADD", "SHFL.f64+DADD", "LDS.64+F2I", "DDIV", "F2I+I2F+FADD"};
#define RUN(OP, THREADS) chain<OP><<<1, THREADS>>>(out, cyc, 1.0000001, 0.9999999, 1.0000001f); cudaDeviceSynchronize(); \
  printf("%-34s threads %4d : %.1f cycles/iter\n", names[OP], THREADS, (double)cyc[0] / N);
  RUN(0, 32) RUN(0, 320) RUN(0, 1024) RUN(1, 32) RUN(1, 320) RUN(2, 32) RUN(3, 32) RUN(3, 320) RUN(4, 32) RUN(4,320) RUN(5, 32) RUN(6, 32) RUN(6, 320) RUN(7, 32) RUN(7, 320)
  return 0;
}
