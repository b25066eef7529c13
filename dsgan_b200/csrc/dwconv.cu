// Depthwise k x k convolution (k = 3,5,7,9; stride 1; pad k/2) — CUDA-core, HBM/L2-bound.
// Reference op sites: Block.dwconv (MixConvNeXtML.py:220) and MidMLKA.X3/X5/X7/X9 (:94-97).
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
template <typename T, int K>
__global__ void k_dwconv(const T* __restrict__ x, int ldx, const float* __restrict__ w,
                         const float* __restrict__ bias, T* __restrict__ y, int ldy, int H, int W, int C, int flip,
                         int acc, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pix = i / C;
  const int px = (int)(pix % W);
  const long long r = pix / W;
  const int py = (int)(r % H);
  const long long n = r / H;
  constexpr int P = K / 2;
  float wv[K * K];
#pragma unroll
  for (int j = 0; j < K * K; ++j) wv[j] = __ldg(w + c * K * K + (flip ? (K * K - 1 - j) : j));
  float s = bias ? __ldg(bias + c) : 0.f;
  const T* xb = x + n * H * W * ldx + c;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const int iy = py + ky - P;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const int ix = px + kx - P;
      if (ix < 0 || ix >= W) continue;
      s = fmaf(ldf(xb + ((long long)iy * W + ix) * ldx), wv[ky * K + kx], s);
    }
  }
  T* o = y + pix * ldy + c;
  if (acc) s += ldf(o);
  stf(o, s);
}

// grid: (channel tiles of CL, pixel chunks); block 256 = CL channel lanes x PL pixel lanes
template <typename T, int K>
__global__ void __launch_bounds__(256) k_dwconv_wgrad(const T* __restrict__ x, int ldx, const T* __restrict__ dy,
                                                       int lddy, float* __restrict__ dw, float* __restrict__ db,
                                                       int H, int W, int C, int cl, long long npix, long long chunk) {
  __shared__ float sh[K * K + 1][32];
  const int pl = 256 / cl;
  const int tc = threadIdx.x % cl, tp = threadIdx.x / cl;
  const int c = blockIdx.x * cl + tc;
  for (int j = threadIdx.x; j < (K * K + 1) * 32; j += 256) (&sh[0][0])[j] = 0.f;
  __syncthreads();
  constexpr int P = K / 2;
  float acc[K * K];
#pragma unroll
  for (int j = 0; j < K * K; ++j) acc[j] = 0.f;
  float accb = 0.f;
  const long long p0 = (long long)blockIdx.y * chunk, p1 = min(p0 + chunk, npix);
  if (c < C && tp < pl) {
    for (long long p = p0 + tp; p < p1; p += pl) {
      const float g = ldf(dy + p * lddy + c);
      accb += g;
      const int px = (int)(p % W);
      const long long r = p / W;
      const int py = (int)(r % H);
      const T* xb = x + (r - py) * W * ldx + c;  // image base
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = py + ky - P;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = px + kx - P;
          if (ix < 0 || ix >= W) continue;
          acc[ky * K + kx] = fmaf(g, ldf(xb + ((long long)iy * W + ix) * ldx), acc[ky * K + kx]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K * K; ++j) atomicAdd(&sh[j][tc], acc[j]);
    atomicAdd(&sh[K * K][tc], accb);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < (K * K + 1) * cl; j += 256) {
    const int tap = j / cl, lc = j % cl, cc = blockIdx.x * cl + lc;
    if (cc >= C) continue;
    if (tap < K * K) atomicAdd(dw + cc * K * K + tap, sh[tap][lc]);
    else if (db) atomicAdd(db + cc, sh[tap][lc]);
  }
}
}  // namespace

extern "C" {
int dsgan_dwconv_fwd(const void* x, int ld_x, const float* w, const float* bias, void* y, int ld_y, int dtype, int N,
                     int H, int W, int C, int k, int flip, int accumulate, void* stream) {
  const long long total = (long long)N * H * W * C;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = cdiv(total, 256);
#define DW_CASE(KK)                                                                                             \
  case KK:                                                                                                      \
    DS_DISPATCH_DT(dtype, (k_dwconv<T, KK><<<grid, 256, 0, s>>>((const T*)x, ld_x, w, bias, (T*)y, ld_y, H, W, C,  \
                                                                 flip, accumulate, total)));                    \
    break;
  switch (k) {
    DW_CASE(3) DW_CASE(5) DW_CASE(7) DW_CASE(9)
    default: set_error("dwconv: unsupported k=%d", k); return 1;
  }
#undef DW_CASE
  return DS_LAUNCHED("dwconv_fwd");
}

int dsgan_dwconv_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, float* dw, float* db, int dtype, int N,
                       int H, int W, int C, int k, void* stream) {
  const long long npix = (long long)N * H * W;
  const int cl = C < 32 ? C : 32;
  DS_REQUIRE(256 % cl == 0 || cl == 3, "dwconv_wgrad: C=%d unsupported channel tiling", C);
  long long chunk = 2048;
  dim3 grid(cdiv(C, cl), cdiv(npix, chunk));
  cudaStream_t s = (cudaStream_t)stream;
#define DWG_CASE(KK)                                                                                              \
  case KK:                                                                                                        \
    DS_DISPATCH_DT(dtype, (k_dwconv_wgrad<T, KK><<<grid, 256, 0, s>>>((const T*)x, ld_x, (const T*)dy, ld_dy, dw, db, \
                                                                       H, W, C, cl, npix, chunk)));               \
    break;
  switch (k) {
    DWG_CASE(3) DWG_CASE(5) DWG_CASE(7) DWG_CASE(9)
    default: set_error("dwconv_wgrad: unsupported k=%d", k); return 1;
  }
#undef DWG_CASE
  return DS_LAUNCHED("dwconv_wgrad");
}
}
