// InstanceNorm2d(affine=False, eps=1e-5, biased variance) fused with activation / residual / concat-slice
// writes; forward and backward.  HBM-bound: warp/thread-private partial sums, one atomic per (block, channel).
// Reference: networks.py:25 and every nn.InstanceNorm2d in MixConvNeXtML.py (Q1 in SURVEY.md §10).
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
constexpr float EPS = 1e-5f;
constexpr int CHUNK = 512;  // pixels per block

struct Lanes { int cl, pl, tc, tp; };
__device__ __forceinline__ Lanes lanes(int C) {
  Lanes l;
  l.cl = C < 64 ? C : 64;
  l.pl = blockDim.x / l.cl;
  l.tc = threadIdx.x % l.cl;
  l.tp = threadIdx.x / l.cl;
  return l;
}
__device__ __forceinline__ void mean_rstd(const float* st, float inv_hw, float& mean, float& rstd) {
  const float k = st[0], s = st[1], ss = st[2];
  const float m = s * inv_hw;
  float var = ss * inv_hw - m * m;
  var = var < 0.f ? 0.f : var;
  mean = k + m;
  rstd = rsqrtf(var + EPS);
}

template <typename T>
__global__ void k_in_stats(const T* __restrict__ x, int ldx, long long HW, int C, float* __restrict__ stats) {
  const Lanes l = lanes(C);
  if (l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * CHUNK, p1 = min(p0 + (long long)CHUNK, HW);
  const T* xb = x + (long long)n * HW * ldx;
  for (int c = l.tc; c < C; c += l.cl) {
    const float k = ldf(xb + c);  // shift = first pixel of the plane: keeps the sums well conditioned
    float s = 0.f, ss = 0.f;
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const float v = ldf(xb + p * ldx + c) - k;
      s += v;
      ss = fmaf(v, v, ss);
    }
    float* st = stats + ((long long)n * C + c) * 3;
    if (blockIdx.x == 0 && l.tp == 0) st[0] = k;
    atomicAdd(st + 1, s);
    atomicAdd(st + 2, ss);
  }
}

template <typename T>
__global__ void k_in_apply(const T* __restrict__ x, int ldx, const float* __restrict__ stats,
                           const T* __restrict__ res, int ldr, T* __restrict__ y, int ldy, long long HW, int C,
                           int act) {
  const Lanes l = lanes(C);
  if (l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * CHUNK, p1 = min(p0 + (long long)CHUNK, HW);
  const long long base = (long long)n * HW;
  const float inv = 1.0f / (float)HW;
  for (int c = l.tc; c < C; c += l.cl) {
    float mean, rstd;
    mean_rstd(stats + ((long long)n * C + c) * 3, inv, mean, rstd);
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      float u = (ldf(x + (base + p) * ldx + c) - mean) * rstd;
      if (res) u += ldf(res + (base + p) * ldr + c);
      stf(y + (base + p) * ldy + c, act_fwd(act, u));
    }
  }
}

template <typename T>
__global__ void k_in_bwd_stats(const T* __restrict__ x, int ldx, const float* __restrict__ stats,
                               const T* __restrict__ res, int ldr, const T* __restrict__ dy, int lddy, long long HW,
                               int C, int act, float* __restrict__ bst) {
  const Lanes l = lanes(C);
  if (l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * CHUNK, p1 = min(p0 + (long long)CHUNK, HW);
  const long long base = (long long)n * HW;
  const float inv = 1.0f / (float)HW;
  for (int c = l.tc; c < C; c += l.cl) {
    float mean, rstd;
    mean_rstd(stats + ((long long)n * C + c) * 3, inv, mean, rstd);
    float sg = 0.f, sgx = 0.f;
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const float xh = (ldf(x + (base + p) * ldx + c) - mean) * rstd;
      float g = ldf(dy + (base + p) * lddy + c);
      if (act) {
        float u = xh;
        if (res) u += ldf(res + (base + p) * ldr + c);
        g *= act_bwd(act, u);
      }
      sg += g;
      sgx = fmaf(g, xh, sgx);
    }
    atomicAdd(bst + ((long long)n * C + c) * 2, sg);
    atomicAdd(bst + ((long long)n * C + c) * 2 + 1, sgx);
  }
}

template <typename T>
__global__ void k_in_bwd_apply(const T* __restrict__ x, int ldx, const float* __restrict__ stats,
                               const T* __restrict__ res, int ldr, const T* __restrict__ dy, int lddy,
                               const float* __restrict__ bst, T* __restrict__ dx, int lddx, int acc_dx,
                               T* __restrict__ dres, int lddr, int acc_dres, long long HW, int C, int act) {
  const Lanes l = lanes(C);
  if (l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * CHUNK, p1 = min(p0 + (long long)CHUNK, HW);
  const long long base = (long long)n * HW;
  const float inv = 1.0f / (float)HW;
  for (int c = l.tc; c < C; c += l.cl) {
    float mean, rstd;
    mean_rstd(stats + ((long long)n * C + c) * 3, inv, mean, rstd);
    const float mg = bst[((long long)n * C + c) * 2] * inv, mgx = bst[((long long)n * C + c) * 2 + 1] * inv;
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const float xh = (ldf(x + (base + p) * ldx + c) - mean) * rstd;
      float g = ldf(dy + (base + p) * lddy + c);
      if (act) {
        float u = xh;
        if (res) u += ldf(res + (base + p) * ldr + c);
        g *= act_bwd(act, u);
      }
      float v = rstd * (g - mg - xh * mgx);
      T* o = dx + (base + p) * lddx + c;
      if (acc_dx) v += ldf(o);
      stf(o, v);
      if (dres) {
        T* r = dres + (base + p) * lddr + c;
        stf(r, acc_dres ? g + ldf(r) : g);
      }
    }
  }
}
}  // namespace

extern "C" {
int dsgan_inorm_stats(const void* x, int ld_x, int dtype, int N, long long HW, int C, float* stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, sizeof(float) * 3 * N * C, s);
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_stats<T><<<grid, 256, 0, s>>>((const T*)x, ld_x, HW, C, stats)));
  return DS_LAUNCHED("inorm_stats");
}
int dsgan_inorm_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, void* y, int ld_y,
                      int dtype, int N, long long HW, int C, int act, void* stream) {
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_apply<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, ld_x, stats, (const T*)res,
                                                                              ld_res, (T*)y, ld_y, HW, C, act)));
  return DS_LAUNCHED("inorm_apply");
}
int dsgan_inorm_bwd_stats(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, int dtype, int N, long long HW, int C, int act, float* bstats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(bstats, 0, sizeof(float) * 2 * N * C, s);
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_bwd_stats<T><<<grid, 256, 0, s>>>((const T*)x, ld_x, stats, (const T*)res, ld_res,
                                                               (const T*)dy, ld_dy, HW, C, act, bstats)));
  return DS_LAUNCHED("inorm_bwd_stats");
}
int dsgan_inorm_bwd_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, const float* bstats, void* dx, int ld_dx, int acc_dx, void* dres, int ld_dres,
                          int acc_dres, int dtype, int N, long long HW, int C, int act, void* stream) {
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_bwd_apply<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, ld_x, stats, (const T*)res, ld_res, (const T*)dy, ld_dy, bstats, (T*)dx, ld_dx,
                            acc_dx, (T*)dres, ld_dres, acc_dres, HW, C, act)));
  return DS_LAUNCHED("inorm_bwd_apply");
}
}
