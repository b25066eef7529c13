// FP32 FMA-pipe throughput on sm_100a: scalar FFMA vs packed FFMA2 (independent chains, all operands in registers).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a0) {
  float2 acc[16];
  float2 a = make_float2(a0, a0 * 1.0001f), b = make_float2(0.5f, 0.25f);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
      else acc[i] = __ffma2_rn(acc[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 8 * 256 * (double)iters * 32;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9,
                      fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  }
  return 0;
}
