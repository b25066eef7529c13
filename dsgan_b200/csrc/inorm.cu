// InstanceNorm2d(affine=False, eps=1e-5, biased variance) fused with activation / residual / concat-slice
// writes; forward and backward.  HBM-bound: warp/thread-private partial sums, one atomic per (block, channel).
// Reference: networks.py:25 and every nn.InstanceNorm2d in MixConvNeXtML.py (Q1 in SURVEY.md §10).
#include "common.cuh"
#include <initializer_list>
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
                               T* __restrict__ dres, int lddr, int acc_dres, long long HW, int C, int act,
                               float* __restrict__ dbias, float* __restrict__ dsum_nc) {
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
    float vsum = 0.f;
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const float xh = (ldf(x + (base + p) * ldx + c) - mean) * rstd;
      float g = ldf(dy + (base + p) * lddy + c);
      if (act) {
        float u = xh;
        if (res) u += ldf(res + (base + p) * ldr + c);
        g *= act_bwd(act, u);
      }
      float v = rstd * (g - mg - xh * mgx);
      vsum += v;
      T* o = dx + (base + p) * lddx + c;
      if (acc_dx) v += ldf(o);
      stf(o, v);
      if (dres) {
        T* r = dres + (base + p) * lddr + c;
        stf(r, acc_dres ? g + ldf(r) : g);
      }
    }
    if (dbias) atomicAdd(dbias + c, vsum);
    if (dsum_nc) atomicAdd(dsum_nc + (long long)n * C + c, vsum);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 fast path: 8 channels (16 bytes) per thread, warp rows of consecutive channel groups (coalesced 512 B), block-level
// reduction through shared-memory atomics, one global atomic per (block, channel).
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(h);
    f[2 * e + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld8(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

struct VLanes { int gl, pl, tg, tp; };
__device__ __forceinline__ VLanes vlanes(int C) {
  VLanes l;
  const int groups = (C + 7) / 8;  // C % 8 != 0 only on a pitch of exactly ceil8(C): the pad lanes ride along
  l.gl = groups < 32 ? groups : 32;
  l.pl = blockDim.x / l.gl;
  l.tg = threadIdx.x % l.gl;
  l.tp = threadIdx.x / l.gl;
  return l;
}

__global__ void __launch_bounds__(256) k_in_stats_v8(const bf16* __restrict__ x, int ldx, long long HW, int C,
                                                      float* __restrict__ stats, int VCHUNK) {
  __shared__ float sacc[2][256];
  const VLanes l = vlanes(C);
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * VCHUNK, p1 = min(p0 + (long long)VCHUNK, HW);
  const bf16* xb = x + (size_t)n * HW * ldx;
  {
    const int cb = blockIdx.z * l.gl * 8;
    for (int i = threadIdx.x; i < 512; i += 256) (&sacc[0][0])[i] = 0.f;
    __syncthreads();
    const int c0 = cb + l.tg * 8;
    if (l.tp < l.pl && c0 < C) {
      float k[8], s[8], ss[8];
      unpack8(ld8(xb + c0), k);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
#pragma unroll 4
      for (long long p = p0 + l.tp; p < p1; p += l.pl) {
        float v[8];
        unpack8(ld8(xb + p * ldx + c0), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[e] - k[e]; s[e] += d; ss[e] = fmaf(d, d, ss[e]); }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(&sacc[0][l.tg * 8 + e], s[e]); atomicAdd(&sacc[1][l.tg * 8 + e], ss[e]); }
      if (blockIdx.x == 0 && l.tp == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (c0 + e < C) stats[((size_t)n * C + c0 + e) * 3] = k[e];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < l.gl * 8; i += 256) {
      if (cb + i < C) {
        atomicAdd(stats + ((size_t)n * C + cb + i) * 3 + 1, sacc[0][i]);
        atomicAdd(stats + ((size_t)n * C + cb + i) * 3 + 2, sacc[1][i]);
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float2 bf2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }
__device__ __forceinline__ uint32_t f2b(float2 v) {
  __nv_bfloat162 h = __float22bfloat162_rn(v);
  return *reinterpret_cast<uint32_t*>(&h);
}
// per-thread normalisation constants of its 8 channels as 4 float2 pairs: xhat = x * rstd + shift, shift = -mean * rstd
__device__ __forceinline__ void norm_consts(const float* stats, int n, int C, int c0, float inv, float2 (&rs)[4], float2 (&sh)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float m0 = 0.f, r0 = 0.f, m1 = 0.f, r1 = 0.f;
    if (c0 + 2 * e < C) mean_rstd(stats + ((size_t)n * C + c0 + 2 * e) * 3, inv, m0, r0);
    if (c0 + 2 * e + 1 < C) mean_rstd(stats + ((size_t)n * C + c0 + 2 * e + 1) * 3, inv, m1, r1);
    rs[e] = make_float2(r0, r1);
    sh[e] = make_float2(-m0 * r0, -m1 * r1);
  }
}

// ACT >= 0: the activation is a compile-time constant (GELU / LeakyReLU / none: every norm of the step), so the other
// activations' code and the per-pair tests of `act` drop out of the loops; ACT < 0: run-time `act_rt`.
// PLAIN: no residual / no second output / no accumulation -- the common case, with those paths compiled out.
template <int ACT, bool PLAIN>
__global__ void __launch_bounds__(256) k_in_apply_v8(const bf16* __restrict__ x, int ldx, const float* __restrict__ stats,
                                                      const bf16* __restrict__ res_rt, int ldr, bf16* __restrict__ y, int ldy,
                                                      long long HW, int C, int act_rt, int VCHUNK) {
  const int act = ACT >= 0 ? ACT : act_rt;
  const bf16* __restrict__ res = PLAIN ? nullptr : res_rt;
  const VLanes l = vlanes(C);
  if (l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * VCHUNK, p1 = min(p0 + (long long)VCHUNK, HW);
  const size_t base = (size_t)n * HW;
  const float inv = 1.0f / (float)HW;
  for (int c0 = blockIdx.z * l.gl * 8 + l.tg * 8; c0 < min(C, (int)(blockIdx.z + 1) * l.gl * 8); c0 += l.gl * 8) {
    float2 rs[4], sh[4];
    norm_consts(stats, n, C, c0, inv, rs, sh);
#pragma unroll 4
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const uint4 xv = ld8(x + (base + p) * ldx + c0);
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (res) rv = ld8(res + (base + p) * ldr + c0);
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 u = __ffma2_rn(bf2(xw[e]), rs[e], sh[e]);
        if (res) u = __fadd2_rn(u, bf2(rw[e]));
        ow[e] = f2b(act_fwd_fast2(act, u));
      }
      *reinterpret_cast<uint4*>(y + (base + p) * ldy + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

template <int ACT, bool PLAIN>
__global__ void __launch_bounds__(256, 4) k_in_bwd_stats_v8(const bf16* __restrict__ x, int ldx,
                                                          const float* __restrict__ stats, const bf16* __restrict__ res_rt,
                                                          int ldr, const bf16* __restrict__ dy, int lddy, long long HW,
                                                          int C, int act_rt, float* __restrict__ bst, int VCHUNK) {
  const int act = ACT >= 0 ? ACT : act_rt;
  const bf16* __restrict__ res = PLAIN ? nullptr : res_rt;
  __shared__ float sacc[2][256];
  const VLanes l = vlanes(C);
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * VCHUNK, p1 = min(p0 + (long long)VCHUNK, HW);
  const size_t base = (size_t)n * HW;
  const float inv = 1.0f / (float)HW;
  {
    const int cb = blockIdx.z * l.gl * 8;
    for (int i = threadIdx.x; i < 512; i += 256) (&sacc[0][0])[i] = 0.f;
    __syncthreads();
    const int c0 = cb + l.tg * 8;
    if (l.tp < l.pl && c0 < C) {
      float2 rs[4], sh[4], sg2[4], sgx2[4];
      norm_consts(stats, n, C, c0, inv, rs, sh);
#pragma unroll
      for (int e = 0; e < 4; ++e) sg2[e] = sgx2[e] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (long long p = p0 + l.tp; p < p1; p += l.pl) {
        const uint4 xv = ld8(x + (base + p) * ldx + c0), gv = ld8(dy + (base + p) * lddy + c0);
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (act && res) rv = ld8(res + (base + p) * ldr + c0);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xh = __ffma2_rn(bf2(xw[e]), rs[e], sh[e]);
          float2 g = bf2(gw[e]);
          if (act) g = __fmul2_rn(g, act_bwd_fast2(act, res ? __fadd2_rn(xh, bf2(rw[e])) : xh));
          sg2[e] = __fadd2_rn(sg2[e], g);
          sgx2[e] = __ffma2_rn(g, xh, sgx2[e]);
        }
      }
      float sg[8], sgx[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { sg[2 * e] = sg2[e].x; sg[2 * e + 1] = sg2[e].y; sgx[2 * e] = sgx2[e].x; sgx[2 * e + 1] = sgx2[e].y; }
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(&sacc[0][l.tg * 8 + e], sg[e]); atomicAdd(&sacc[1][l.tg * 8 + e], sgx[e]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < l.gl * 8; i += 256) {
      if (cb + i < C) {
        atomicAdd(bst + ((size_t)n * C + cb + i) * 2, sacc[0][i]);
        atomicAdd(bst + ((size_t)n * C + cb + i) * 2 + 1, sacc[1][i]);
      }
    }
    __syncthreads();
  }
}

// SUMS: additionally accumulate the per-(n, c) sum of the fp32 dx values BEFORE they are rounded to bf16 (and before any
// fan-in add).  A bias that feeds an InstanceNorm has the gradient sum_p dx[p] = 0 in exact arithmetic; summing the rounded bf16
// tensor afterwards (the old colsum pass) turned that structural zero into rounding noise of the size of a real gradient.
template <bool SUMS, int ACT, bool PLAIN>
__global__ void __launch_bounds__(256, 3) k_in_bwd_apply_v8(const bf16* __restrict__ x, int ldx,
                                                          const float* __restrict__ stats, const bf16* __restrict__ res_rt,
                                                          int ldr, const bf16* __restrict__ dy, int lddy,
                                                          const float* __restrict__ bst, bf16* __restrict__ dx, int lddx,
                                                          int acc_dx_rt, bf16* __restrict__ dres_rt, int lddr, int acc_dres,
                                                          long long HW, int C, int act_rt, int VCHUNK,
                                                          float* __restrict__ dbias, float* __restrict__ dsum_nc) {
  const int act = ACT >= 0 ? ACT : act_rt;
  const bf16* __restrict__ res = PLAIN ? nullptr : res_rt;
  bf16* __restrict__ dres = PLAIN ? nullptr : dres_rt;
  const int acc_dx = PLAIN ? 0 : acc_dx_rt;
  __shared__ float sacc[SUMS ? 256 : 1];
  const VLanes l = vlanes(C);
  if (SUMS) {
    sacc[threadIdx.x] = 0.f;
    __syncthreads();
  }
  if (!SUMS && l.tp >= l.pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * VCHUNK, p1 = min(p0 + (long long)VCHUNK, HW);
  const size_t base = (size_t)n * HW;
  const float inv = 1.0f / (float)HW;
  for (int c0 = blockIdx.z * l.gl * 8 + l.tg * 8; l.tp < l.pl && c0 < min(C, (int)(blockIdx.z + 1) * l.gl * 8); c0 += l.gl * 8) {
    float2 rs[4], sh[4], mg[4], mgx[4], vs[4];
    norm_consts(stats, n, C, c0, inv, rs, sh);
#pragma unroll
    for (int e = 0; e < 4; ++e) vs[e] = make_float2(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* b0 = bst + ((size_t)n * C + c0 + 2 * e) * 2;
      const bool v0 = c0 + 2 * e < C, v1 = c0 + 2 * e + 1 < C;
      mg[e] = make_float2(v0 ? b0[0] * inv : 0.f, v1 ? b0[2] * inv : 0.f);
      mgx[e] = make_float2(v0 ? -b0[1] * inv : 0.f, v1 ? -b0[3] * inv : 0.f);  // negated: t = g - mg + xh * (-mgx)
    }
#pragma unroll 2
    for (long long p = p0 + l.tp; p < p1; p += l.pl) {
      const uint4 xv = ld8(x + (base + p) * ldx + c0), gv = ld8(dy + (base + p) * lddy + c0);
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (act && res) rv = ld8(res + (base + p) * ldr + c0);
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
      float2 g[4], o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 xh = __ffma2_rn(bf2(xw[e]), rs[e], sh[e]);
        g[e] = bf2(gw[e]);
        if (act) g[e] = __fmul2_rn(g[e], act_bwd_fast2(act, res ? __fadd2_rn(xh, bf2(rw[e])) : xh));
        const float2 t = __ffma2_rn(xh, mgx[e], __fadd2_rn(g[e], make_float2(-mg[e].x, -mg[e].y)));
        o[e] = __fmul2_rn(rs[e], t);
        if (SUMS) vs[e] = __fadd2_rn(vs[e], o[e]);
      }
      uint4* dp = reinterpret_cast<uint4*>(dx + (base + p) * lddx + c0);
      if (acc_dx) {
        const uint4 ov = *dp;
        const uint32_t owd[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __fadd2_rn(o[e], bf2(owd[e]));
      }
      *dp = make_uint4(f2b(o[0]), f2b(o[1]), f2b(o[2]), f2b(o[3]));
      if (dres) {
        uint4* rp = reinterpret_cast<uint4*>(dres + (base + p) * lddr + c0);
        if (acc_dres) {
          const uint4 ov = *rp;
          const uint32_t owd[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) g[e] = __fadd2_rn(g[e], bf2(owd[e]));
        }
        *rp = make_uint4(f2b(g[0]), f2b(g[1]), f2b(g[2]), f2b(g[3]));
      }
    }
    if (SUMS) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { atomicAdd(&sacc[l.tg * 8 + 2 * e], vs[e].x); atomicAdd(&sacc[l.tg * 8 + 2 * e + 1], vs[e].y); }
    }
  }
  if (SUMS) {
    __syncthreads();
    const int cb = blockIdx.z * l.gl * 8;
    for (int i = threadIdx.x; i < l.gl * 8; i += 256) {
      if (cb + i < C) {
        if (dbias) atomicAdd(dbias + cb + i, sacc[i]);
        if (dsum_nc) atomicAdd(dsum_nc + (size_t)n * C + cb + i, sacc[i]);
      }
    }
  }
}

// grid for the v8 kernels: (pixel chunks, N, channel blocks of <=256).  All blocks do the same work, so the chunk is chosen
// to make the grid a whole number of waves of the kernel's resident blocks (148 SMs x occupancy): a 1.7-wave grid idles a
// quarter of the machine in its second wave.  Small planes fall back to >= 64-pixel chunks.
template <typename K>
inline int resident_blocks(K kern, int* cache) {
  if (*cache == 0) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
    *cache = sms * (occ > 0 ? occ : 1);
  }
  return *cache;
}
inline dim3 v8grid(long long HW, int N, int C, int* chunk, int slots) {
  const int groups = (C + 7) / 8, gl = groups < 32 ? groups : 32;
  const int cblocks = (C + gl * 8 - 1) / (gl * 8);
  const long long base = (long long)N * cblocks;
  long long bpi = slots / base;            // blocks per (image, channel block) for one full wave
  if (bpi < 1) bpi = 1;
  long long ch = (HW + bpi - 1) / bpi;
  while (ch > 2048) { bpi *= 2; ch = (HW + bpi - 1) / bpi; }   // long chunks: more waves instead
  if (ch < 64) ch = 64;
  if (ch > HW) ch = HW;
  *chunk = (int)ch;
  return dim3((unsigned)((HW + ch - 1) / ch), N, cblocks);
}
inline bool v8ok(int C, std::initializer_list<const void*> ptrs, std::initializer_list<int> lds) {
  const int Cp = (C + 7) / 8 * 8;
  for (const void* p : ptrs) if (p && ((uintptr_t)p % 16)) return false;
  for (int l : lds) if (l % 8 || (C % 8 && l && l != Cp)) return false;  // ragged C: whole-pitch tensors only
  return true;
}
}  // namespace

extern "C" {
int dsgan_inorm_stats(const void* x, int ld_x, int dtype, int N, long long HW, int C, float* stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, sizeof(float) * 3 * N * C, s);
  if (dtype == DT_BF16 && v8ok(C, {x}, {ld_x})) {
    int ch;
    static int slots = 0;
    dim3 g8 = v8grid(HW, N, C, &ch, resident_blocks(k_in_stats_v8, &slots));
    k_in_stats_v8<<<g8, 256, 0, s>>>((const bf16*)x, ld_x, HW, C, stats, ch);
    return DS_LAUNCHED("inorm_stats_v8");
  }
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_stats<T><<<grid, 256, 0, s>>>((const T*)x, ld_x, HW, C, stats)));
  return DS_LAUNCHED("inorm_stats");
}
int dsgan_inorm_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, void* y, int ld_y,
                      int dtype, int N, long long HW, int C, int act, void* stream) {
  if (dtype == DT_BF16 && v8ok(C, {x, res, y}, {ld_x, res ? ld_res : 0, ld_y})) {
    int ch;
    static int slots = 0;
    dim3 g8 = v8grid(HW, N, C, &ch, resident_blocks(k_in_apply_v8<ACT_GELU, false>, &slots));
#define IN_APPLY(A)                                                                                                         \
  do {                                                                                                                      \
    if (res) k_in_apply_v8<A, false><<<g8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld_x, stats, (const bf16*)res,   \
                                                                           ld_res, (bf16*)y, ld_y, HW, C, act, ch);       \
    else k_in_apply_v8<A, true><<<g8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld_x, stats, nullptr, 0, (bf16*)y,    \
                                                                      ld_y, HW, C, act, ch);                                \
  } while (0)
    if (act == ACT_GELU) IN_APPLY(ACT_GELU); else if (act == ACT_LEAKY) IN_APPLY(ACT_LEAKY); else if (act == ACT_NONE) IN_APPLY(ACT_NONE);
    else IN_APPLY(-1);
#undef IN_APPLY
    return DS_LAUNCHED("inorm_apply_v8");
  }
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_apply<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, ld_x, stats, (const T*)res,
                                                                              ld_res, (T*)y, ld_y, HW, C, act)));
  return DS_LAUNCHED("inorm_apply");
}
int dsgan_inorm_bwd_stats(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, int dtype, int N, long long HW, int C, int act, float* bstats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(bstats, 0, sizeof(float) * 2 * N * C, s);
  if (dtype == DT_BF16 && v8ok(C, {x, res, dy}, {ld_x, res ? ld_res : 0, ld_dy})) {
    int ch;
    static int slots = 0;
    dim3 g8 = v8grid(HW, N, C, &ch, resident_blocks(k_in_bwd_stats_v8<ACT_GELU, false>, &slots));
#define IN_BSTATS(A)                                                                                                        \
  do {                                                                                                                      \
    if (res) k_in_bwd_stats_v8<A, false><<<g8, 256, 0, s>>>((const bf16*)x, ld_x, stats, (const bf16*)res, ld_res,          \
                                                            (const bf16*)dy, ld_dy, HW, C, act, bstats, ch);               \
    else k_in_bwd_stats_v8<A, true><<<g8, 256, 0, s>>>((const bf16*)x, ld_x, stats, nullptr, 0, (const bf16*)dy, ld_dy, HW, \
                                                       C, act, bstats, ch);                                                 \
  } while (0)
    if (act == ACT_GELU) IN_BSTATS(ACT_GELU); else if (act == ACT_LEAKY) IN_BSTATS(ACT_LEAKY); else if (act == ACT_NONE) IN_BSTATS(ACT_NONE);
    else IN_BSTATS(-1);
#undef IN_BSTATS
    return DS_LAUNCHED("inorm_bwd_stats_v8");
  }
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_bwd_stats<T><<<grid, 256, 0, s>>>((const T*)x, ld_x, stats, (const T*)res, ld_res,
                                                               (const T*)dy, ld_dy, HW, C, act, bstats)));
  return DS_LAUNCHED("inorm_bwd_stats");
}
int dsgan_inorm_bwd_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, const float* bstats, void* dx, int ld_dx, int acc_dx, void* dres, int ld_dres,
                          int acc_dres, int dtype, int N, long long HW, int C, int act, float* dbias, float* dsum_nc,
                          void* stream) {
  if (dtype == DT_BF16 && v8ok(C, {x, res, dy, dx, dres}, {ld_x, res ? ld_res : 0, ld_dy, ld_dx, dres ? ld_dres : 0})) {
    int ch;
    const bool plain = !res && !dres && !acc_dx;
#define IN_BAPPLY_P(SUMS, A, P, DB, DS)                                                                                         \
  k_in_bwd_apply_v8<SUMS, A, P><<<g8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld_x, stats, (const bf16*)res, ld_res,    \
                                                                      (const bf16*)dy, ld_dy, bstats, (bf16*)dx, ld_dx, acc_dx, \
                                                                      (bf16*)dres, ld_dres, acc_dres, HW, C, act, ch, DB, DS)
#define IN_BAPPLY(SUMS, A, DB, DS)                                                                                              \
  do { if (plain) IN_BAPPLY_P(SUMS, A, true, DB, DS); else IN_BAPPLY_P(SUMS, A, false, DB, DS); } while (0)
#define IN_BAPPLY_ACT(SUMS, DB, DS)                                                                                             \
  if (act == ACT_GELU) IN_BAPPLY(SUMS, ACT_GELU, DB, DS); else if (act == ACT_LEAKY) IN_BAPPLY(SUMS, ACT_LEAKY, DB, DS);         \
  else if (act == ACT_NONE) IN_BAPPLY(SUMS, ACT_NONE, DB, DS); else IN_BAPPLY(SUMS, -1, DB, DS)
    if (dbias || dsum_nc) {
      static int slots = 0;
      dim3 g8 = v8grid(HW, N, C, &ch, resident_blocks(k_in_bwd_apply_v8<true, ACT_GELU, false>, &slots));
      IN_BAPPLY_ACT(true, dbias, dsum_nc);
    } else {
      static int slots = 0;
      dim3 g8 = v8grid(HW, N, C, &ch, resident_blocks(k_in_bwd_apply_v8<false, ACT_GELU, false>, &slots));
      IN_BAPPLY_ACT(false, nullptr, nullptr);
    }
#undef IN_BAPPLY_ACT
#undef IN_BAPPLY
#undef IN_BAPPLY_P
    return DS_LAUNCHED("inorm_bwd_apply_v8");
  }
  dim3 grid(cdiv(HW, CHUNK), N);
  DS_DISPATCH_DT(dtype, (k_in_bwd_apply<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, ld_x, stats, (const T*)res, ld_res, (const T*)dy, ld_dy, bstats, (T*)dx, ld_dx,
                            acc_dx, (T*)dres, ld_dres, acc_dres, HW, C, act, dbias, dsum_nc)));
  return DS_LAUNCHED("inorm_bwd_apply");
}
}
