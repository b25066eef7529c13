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
// nn.GELU() is the erf form.  Where the result is rounded to bf16 (tcgen05 epilogues, the bf16 InstanceNorm kernels) erf is
// evaluated WITHOUT the special-function unit: erf(z) = z * P(t), t = 2 z^2 / ZM^2 - 1 in [-1, 1], z clamped to
// +-ZM = 3.3 (erf(3.3) = 1 - 3e-6), P = degree-10 least-squares fit in the Chebyshev-conditioned variable t, scaled so
// that erf(+-ZM) = +-1 exactly (the far tails of GELU are exactly x and 0).  fp32 Horner error: |erf| 3.6e-6,
// |GELU| 7.4e-6, |GELU'| 1.9e-6 over [-10, 10] (checked against float64) -- three orders below the bf16 rounding of
// the result.  All FFMA2-packed: ~10 issue slots per element and no MUFU (the previous A&S 7.1.26 form needed a
// reciprocal and an exponential per element and was MUFU-bound in the 4C-hidden GEMM epilogues).  gelu'(x) additionally
// needs phi(x) = exp(-x^2/2)/sqrt(2 pi): one ex2.approx.  The fp32 validation mode keeps erff.
#define DS_ERF_ZM 3.3f
#define DS_ERF_A 1.836547291e-01f
#define DS_ERF_C0 4.281360372e-01f
#define DS_ERF_C1 -2.116433188e-01f
#define DS_ERF_C2 1.520669164e-01f
#define DS_ERF_C3 -1.144481841e-01f
#define DS_ERF_C4 8.414954066e-02f
#define DS_ERF_C5 -5.929519792e-02f
#define DS_ERF_C6 3.648884681e-02f
#define DS_ERF_C7 -1.777309736e-02f
#define DS_ERF_C8 1.122431634e-02f
#define DS_ERF_C9 -9.507629913e-03f
#define DS_ERF_C10 3.632073768e-03f
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 f2dup(float v) { return make_float2(v, v); }
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)).  Scalar FFMAs with LITERAL coefficients: the Horner step p = p * t + c_i then has two
// register operands and an immediate addend (SASS `FFMA R, R, R, imm`), which issues at the full 128 FMA/clk/SM.  The
// packed form used in round 1 (FFMA2 with the coefficient as a third 64-bit register operand) is register-bandwidth bound at
// ~78 FMA/clk/SM (scripts/micro/ffma2_rate.cu) -- it made the fused MLP epilogues ALU-bound 1.6x earlier than necessary.
__device__ __forceinline__ float gelu_cdf_fast(float x) {
  float z = x * 0.70710678118654752f;
  z = fminf(fmaxf(z, -DS_ERF_ZM), DS_ERF_ZM);
  const float t = fmaf(z * z, DS_ERF_A, -1.0f);
  float p = fmaf(t, DS_ERF_C10, DS_ERF_C9);
  p = fmaf(p, t, DS_ERF_C8);
  p = fmaf(p, t, DS_ERF_C7);
  p = fmaf(p, t, DS_ERF_C6);
  p = fmaf(p, t, DS_ERF_C5);
  p = fmaf(p, t, DS_ERF_C4);
  p = fmaf(p, t, DS_ERF_C3);
  p = fmaf(p, t, DS_ERF_C2);
  p = fmaf(p, t, DS_ERF_C1);
  p = fmaf(p, t, DS_ERF_C0);
  return fmaf(p, z * 0.5f, 0.5f);      // 0.5 + 0.5 * z * P(t)
}
__device__ __forceinline__ float2 gelu_cdf_fast2(float2 x) { return make_float2(gelu_cdf_fast(x.x), gelu_cdf_fast(x.y)); }
__device__ __forceinline__ float gelu_pdf_fast(float x) {
  return 0.39894228040143268f * ex2_approx(x * x * -0.72134752044448170f);   // exp(-x^2/2) / sqrt(2 pi), one MUFU
}
__device__ __forceinline__ void gelu_parts_fast2(float2 x, float2& cdf, float2& pdf) {
  cdf = gelu_cdf_fast2(x);
  pdf = make_float2(gelu_pdf_fast(x.x), gelu_pdf_fast(x.y));
}
__device__ __forceinline__ void gelu_parts_fast(float x, float& cdf, float& pdf) {
  cdf = gelu_cdf_fast(x);
  pdf = gelu_pdf_fast(x);
}
// gelu'(x) = Phi(x) + x phi(x) directly: gelu'(x) - 0.5 is odd, = x * Q(t), t = 2 x^2 / 25 - 1, |x| clamped to 5 (beyond: 1 + 1e-5
// and -1e-5 instead of 1 and 0), Q of degree 12; fp32 Horner error 9.8e-6 against float64.  No MUFU, ~19 issue slots per pair.
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float xc = fminf(fmaxf(x, -5.0f), 5.0f);
  const float t = fmaf(xc * xc, 0.08f, -1.0f);
  float p = fmaf(t, 1.020706059e-02f, -2.475356668e-02f);
  p = fmaf(p, t, 1.795447979e-02f);
  p = fmaf(p, t, -1.767319918e-02f);
  p = fmaf(p, t, 5.105054164e-02f);
  p = fmaf(p, t, -7.745690087e-02f);
  p = fmaf(p, t, 8.084683210e-02f);
  p = fmaf(p, t, -8.186698327e-02f);
  p = fmaf(p, t, 8.015732403e-02f);
  p = fmaf(p, t, -7.208436384e-02f);
  p = fmaf(p, t, 6.658681821e-02f);
  p = fmaf(p, t, -7.509966350e-02f);
  p = fmaf(p, t, 1.421335801e-01f);
  return fmaf(p, xc, 0.5f);
}
__device__ __forceinline__ float2 gelu_grad_fast2(float2 x) { return make_float2(gelu_grad_fast(x.x), gelu_grad_fast(x.y)); }
__device__ __forceinline__ float act_fwd_fast(int act, float v) {
  if (act == ACT_GELU) return v * gelu_cdf_fast(v);
  return act_fwd(act, v);
}
__device__ __forceinline__ float act_bwd_fast(int act, float a) {
  if (act == ACT_GELU) return gelu_grad_fast(a);
  return act_bwd(act, a);
}
// (every non-GELU activation is written out branch-free under ONE uniform test of `act`: going through act_fwd()'s switch
//  per element compiled to an indirect branch per element -- 23 % of the VGG conv1_2 kernel's time in the ReLU epilogue)
__device__ __forceinline__ float2 act_fwd_fast2(int act, float2 u) {
  if (act == ACT_GELU) return __fmul2_rn(u, gelu_cdf_fast2(u));
  if (act == ACT_NONE) return u;
  if (act == ACT_RELU) return make_float2(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f));
  if (act == ACT_LEAKY) return make_float2(fmaxf(u.x, 0.2f * u.x), fmaxf(u.y, 0.2f * u.y));
  return make_float2(act_fwd(act, u.x), act_fwd(act, u.y));
}
// derivative; the argument is the activation OUTPUT for relu/leaky/sigmoid and the PRE-activation for gelu
__device__ __forceinline__ float2 act_bwd_fast2(int act, float2 u) {
  if (act == ACT_GELU) return gelu_grad_fast2(u);
  if (act == ACT_NONE) return make_float2(1.f, 1.f);
  if (act == ACT_RELU) return make_float2(u.x > 0.f ? 1.f : 0.f, u.y > 0.f ? 1.f : 0.f);
  if (act == ACT_LEAKY) return make_float2(u.x > 0.f ? 1.f : 0.2f, u.y > 0.f ? 1.f : 0.2f);
  return make_float2(act_bwd(act, u.x), act_bwd(act, u.y));
}
// v[0..n) <- act(v) / v *= act'(a), two at a time for GELU
template <int N>
__device__ __forceinline__ void act_fwd_fast_vec(int act, float (&v)[N]) {
  if (act == ACT_NONE) return;
  if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const float2 x = make_float2(v[j], v[j + 1]);
      const float2 r = __fmul2_rn(x, gelu_cdf_fast2(x));
      v[j] = r.x; v[j + 1] = r.y;
    }
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_fwd(act, v[j]);
  }
}
template <int N>
__device__ __forceinline__ void act_bwd_fast_mul(int act, float (&v)[N], const float (&a)[N]) {
  if (act == ACT_NONE) return;
  if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const float2 d = gelu_grad_fast2(make_float2(a[j], a[j + 1]));
      v[j] *= d.x; v[j + 1] *= d.y;
    }
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = a[j] > 0.f ? v[j] : 0.f;
  } else if (act == ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = a[j] > 0.f ? v[j] : 0.2f * v[j];
  } else {
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
