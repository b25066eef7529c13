// Shared helpers for the dsgan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace dsgan {

typedef __nv_bfloat16 bf16;

enum { DT_F32 = 0, DT_BF16 = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_GELU = 3, ACT_SIGMOID = 4 };

// ---- error / bookkeeping (host) -------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
extern unsigned long long g_launches;
#define DS_LAUNCHED(name) (++::dsgan::g_launches, ::dsgan::check_launch(name))
#define DS_REQUIRE(cond, ...) do { if (!(cond)) { ::dsgan::set_error(__VA_ARGS__); return 1; } } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- element access -------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- activations ----------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float act_fwd(int act, float v) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LEAKY: return v > 0.f ? v : 0.2f * v;
    case ACT_GELU: return gelu_f(v);
    case ACT_SIGMOID: return 1.0f / (1.0f + __expf(-v));
    default: return v;
  }
}
// derivative; `a` is the activation OUTPUT for relu/leaky/sigmoid and the PRE-activation for gelu
__device__ __forceinline__ float act_bwd(int act, float a) {
  switch (act) {
    case ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case ACT_LEAKY: return a > 0.f ? 1.f : 0.2f;
    case ACT_GELU: return gelu_grad_f(a);
    case ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}

// ---- fast GELU for bf16 epilogues ------------------------------------------------------------------------------
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 resolution): 1 rcp + 1 ex2 + ~8 FMA instead of
// the ~30-instruction erff.  Used only where the result is rounded to bf16 (tcgen05 epilogues); the fp32 validation
// mode keeps erff.  gelu'(x) shares the exponential: phi(x) = exp(-x^2/2)/sqrt(2 pi).
__device__ __forceinline__ void gelu_parts_fast(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  const float e = __expf(-z * z);  // = exp(-x^2/2)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfz = 1.0f - p * t * e;  // erf(|x|/sqrt2)
  cdf = 0.5f * (1.0f + copysignf(erfz, x));
  pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float act_fwd_fast(int act, float v) {
  if (act == ACT_GELU) { float c, p; gelu_parts_fast(v, c, p); return v * c; }
  return act_fwd(act, v);
}
__device__ __forceinline__ float act_bwd_fast(int act, float a) {
  if (act == ACT_GELU) { float c, p; gelu_parts_fast(a, c, p); return fmaf(a, p, c); }
  return act_bwd(act, a);
}

// packed (two elements per instruction, sm_100 FFMA2/FMUL2/FADD2) version of the fast GELU parts
__device__ __forceinline__ void gelu_parts_fast2(float2 x, float2& cdf, float2& pdf) {
  const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  const float2 t = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
  const float2 a = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 e = make_float2(exp2f(a.x), exp2f(a.y));  // exp(-z^2) = exp(-x^2/2)
  float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  const float2 pte = __fmul2_rn(__fmul2_rn(p, t), e);
  const float2 erfz = make_float2(copysignf(1.0f - pte.x, x.x), copysignf(1.0f - pte.y, x.y));
  cdf = __ffma2_rn(make_float2(0.5f, 0.5f), erfz, make_float2(0.5f, 0.5f));
  pdf = __fmul2_rn(make_float2(0.39894228040143268f, 0.39894228040143268f), e);
}
__device__ __forceinline__ float2 act_fwd_fast2(int act, float2 u) {
  if (act == ACT_GELU) { float2 c, p; gelu_parts_fast2(u, c, p); return __fmul2_rn(u, c); }
  if (act == ACT_NONE) return u;
  return make_float2(act_fwd(act, u.x), act_fwd(act, u.y));
}
__device__ __forceinline__ float2 act_bwd_fast2(int act, float2 u) {
  if (act == ACT_GELU) { float2 c, p; gelu_parts_fast2(u, c, p); return __ffma2_rn(u, p, c); }
  if (act == ACT_NONE) return make_float2(1.f, 1.f);
  return make_float2(act_bwd(act, u.x), act_bwd(act, u.y));
}
// v[0..n) <- act(v) / v *= act'(a), two at a time for GELU
template <int N>
__device__ __forceinline__ void act_fwd_fast_vec(int act, float (&v)[N]) {
  if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      float2 c, p;
      const float2 x = make_float2(v[j], v[j + 1]);
      gelu_parts_fast2(x, c, p);
      const float2 r = __fmul2_rn(x, c);
      v[j] = r.x; v[j + 1] = r.y;
    }
  } else if (act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_fwd(act, v[j]);
  }
}
template <int N>
__device__ __forceinline__ void act_bwd_fast_mul(int act, float (&v)[N], const float (&a)[N]) {
  if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      float2 c, p;
      const float2 x = make_float2(a[j], a[j + 1]);
      gelu_parts_fast2(x, c, p);
      const float2 d = __ffma2_rn(x, p, c);
      v[j] *= d.x; v[j + 1] *= d.y;
    }
  } else if (act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] *= act_bwd(act, a[j]);
  }
}

// ---- reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; every thread gets the result. `sh` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

}  // namespace dsgan

#define DS_DISPATCH_DT(dt, ...)                                     \
  do {                                                              \
    if ((dt) == ::dsgan::DT_F32) { typedef float T; __VA_ARGS__; }  \
    else if ((dt) == ::dsgan::DT_BF16) { typedef ::dsgan::bf16 T; __VA_ARGS__; } \
    else { ::dsgan::set_error("bad dtype %d", (int)(dt)); return 1; } \
  } while (0)
