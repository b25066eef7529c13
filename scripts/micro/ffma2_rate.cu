// FP32 FMA-pipe throughput on sm_100a: scalar FFMA vs packed FFMA2 (all operands in registers).
//   mode 0/1: independent chains sharing both multiplicands (best case, operand-reuse cache hits)
//   mode 3  : the same pattern with scalar FFMA
//   mode 2  : the depthwise-convolution pattern: acc[ox] += v[xi] * w[kx], 7 accumulators per input value, sliding window
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a0) {
  float2 acc[32];
  float2 a = make_float2(a0, a0 * 1.0001f), b = make_float2(0.5f, 0.25f);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 w[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) w[i] = make_float2(a0 + i * 1e-4f, a0 - i * 1e-4f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 3) {
#pragma unroll
      for (int xi = 0; xi < 38; ++xi) {
        const float vx = acc[xi & 31].y * 1e-9f + a.x, vy = a.y;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ox = xi - kx;
          if (ox >= 0 && ox < 32) { acc[ox].x = fmaf(vx, w[kx].x, acc[ox].x); acc[ox].y = fmaf(vy, w[kx].y, acc[ox].y); }
        }
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int xi = 0; xi < 38; ++xi) {
        const float2 v = make_float2(acc[xi & 31].y * 1e-9f + a.x, a.y);   // a fresh value per input column
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ox = xi - kx;
          if (ox >= 0 && ox < 32) acc[ox] = __ffma2_rn(v, w[kx], acc[ox]);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
          else acc[i] = __ffma2_rn(acc[i], a, b);
        }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  const char* names[4] = {"FFMA  (shared multiplicands)", "FFMA2 (shared multiplicands)", "FFMA2 (depthwise sliding window)", "FFMA  (depthwise sliding window)"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 256>>>(out, iters, 0.999f);
      else if (mode == 1) k<1><<<148 * 4, 256>>>(out, iters, 0.999f);
      else if (mode == 2) k<2><<<148 * 4, 256>>>(out, iters, 0.999f);
      else k<3><<<148 * 4, 256>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 4 * 256 * (double)iters * 2 * 224;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s (%.1f FMA/clk/SM at 1.965 GHz)\n", names[mode], ms, 2 * fma / ms / 1e9,
                      fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  }
  return 0;
}
