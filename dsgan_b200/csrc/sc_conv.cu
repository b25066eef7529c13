// Small-channel convolutions on the CUDA cores (packed FFMA2), NHWC bf16.
//
// The tcgen05 implicit GEMM (tc_conv.cu) needs K = 64 input channels per tap and N >= 32 output channels to feed the
// tensor pipe; layers with 1/3/6/12 channels on one side (VGG conv1_1 3->64, the PatchGAN's first 6->64 conv, the
// generator's 64->3 `res` conv and the 3 -> 12 -> 64 MLP of block c1, their input- and weight-gradients) spend their time
// moving zero padding there.  These layers are HBM-bound (one side is a 64-channel 256x256 tensor, the arithmetic is
// < 2 GFMA), so they run here as register-tiled direct convolutions with the same contract as dsgan_tc_conv /
// dsgan_tc_conv_wgrad (same descriptors, same packed bf16 weight slabs, same fused epilogue):
//   k_sc_conv_in  : narrow INPUT (Ci = 1/3/6/12).  A thread owns PT = 4 consecutive grid positions x 8 output channels;
//                   the whole pixel is one or two 16-byte loads; weights in smem as fp32, conflict-free 16-byte reads.
//   k_sc_conv_out : narrow OUTPUT (Co <= 16, Ci a multiple of 8).  LC = Ci/8 lanes share a quad of positions, each lane
//                   owns one 16-byte channel chunk (a warp reads whole 128-byte pixels), partial sums are combined with
//                   warp shuffles, lane j of the group finishes position j.
//   k_sc_wgrad    : a thread owns one (tap, 8-channel group of the wide side) and all channels of the narrow side;
//                   partial sums go through shared-memory atomics to one global atomic per (block, weight).
// Reference op sites: models/vgg.py:16 (conv1_1), networks.py:544 (PatchGAN first conv),
// MixConvNeXtML.py:222-226 (block c1), :459 (`res`), :150-160 (OriginMLKA to32 / shortcut).
#include "common.cuh"
#include "sc_conv.cuh"
#include <string.h>

using namespace dsgan;

namespace {
constexpr int PT = 4;           // grid positions (consecutive in x) per thread
constexpr int SC_THREADS = 256;

struct ScParams {
  int N, Hg, Wg, Hi, Wi, Ho, Wo, Ci, Co, co_pad, ci_pad;
  int ld_in, ldc, ld_pre, ld_aux;
  int is_, os_, nclass, cog, qx, lc;
  int oy0[4], ox0[4], ntaps[4], tap0[4];
  int dy[64], dx[64], slab[64];
  int act, dact, accumulate;
  unsigned items;
  const bf16* in; const bf16* w; bf16* out; const float* bias; bf16* pre; const bf16* aux;
};

__device__ __forceinline__ float2 bf2f(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2bf(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int V>
__device__ __forceinline__ void ld_vec(const bf16* ptr, uint32_t (&w)[V / 2]) {
#pragma unroll
  for (int q = 0; q < V / 8; ++q) {
    const uint4 t = *reinterpret_cast<const uint4*>(ptr + q * 8);
    w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
  }
  if constexpr (V == 4) { const uint2 t = *reinterpret_cast<const uint2*>(ptr); w[0] = t.x; w[1] = t.y; }
}
template <int V>
__device__ __forceinline__ void st_vec(bf16* ptr, const float (&f)[V]) {
#pragma unroll
  for (int q = 0; q < V / 8; ++q)
    *reinterpret_cast<uint4*>(ptr + q * 8) = make_uint4(f2bf(f[8 * q], f[8 * q + 1]), f2bf(f[8 * q + 2], f[8 * q + 3]),
                                                        f2bf(f[8 * q + 4], f[8 * q + 5]), f2bf(f[8 * q + 6], f[8 * q + 7]));
  if constexpr (V == 4) *reinterpret_cast<uint2*>(ptr) = make_uint2(f2bf(f[0], f[1]), f2bf(f[2], f[3]));
}

// f[0..V) holds the raw sums of output channels co0.. of one output pixel: + bias, (+ old), * act'(aux), pre, act, store
template <int V>
__device__ __forceinline__ void epilogue(const ScParams& p, size_t pix, int co0, float (&f)[V], const float (&bv)[V]) {
#pragma unroll
  for (int e = 0; e < V; ++e) f[e] += bv[e];
  bf16* o = p.out + pix * p.ldc + co0;
  if (p.accumulate) {
    uint32_t w[V / 2];
    ld_vec<V>(o, w);
#pragma unroll
    for (int e = 0; e < V / 2; ++e) { const float2 h = bf2f(w[e]); f[2 * e] += h.x; f[2 * e + 1] += h.y; }
  }
  if (p.dact) {
    uint32_t w[V / 2];
    ld_vec<V>(p.aux + pix * p.ld_aux + co0, w);
    float a[V];
#pragma unroll
    for (int e = 0; e < V / 2; ++e) { const float2 h = bf2f(w[e]); a[2 * e] = h.x; a[2 * e + 1] = h.y; }
    act_bwd_fast_mul<V>(p.dact, f, a);
  }
  if (p.pre) st_vec<V>(p.pre + pix * p.ld_pre + co0, f);
  act_fwd_fast_vec<V>(p.act, f);
  st_vec<V>(o, f);
}

__device__ __forceinline__ int tap_of(const ScParams& p, int tt) {   // compact tap index -> descriptor tap index
  int cls = 0;
  while (cls + 1 < p.nclass && tt >= p.tap0[cls + 1]) ++cls;
  return cls * 16 + (tt - p.tap0[cls]);
}

// ---- narrow input: CI = exact channel count (1, 3, 6, 12), COT = 8 output channels per thread -----------------------
template <int CI>
__global__ void __launch_bounds__(SC_THREADS, 3) k_sc_conv_in(const __grid_constant__ ScParams p) {
  extern __shared__ __align__(16) float sw[];
  constexpr int COT = 8, HV = 2, NCH = (CI + 7) / 8;
  const int cog = p.cog;
  // sw[((tt*CI + ci)*HV + h)*cog*4 + g*4 + e] = W[tap tt][co = g*8 + h*4 + e][ci]
  {
    int ttot = 0;
    for (int c = 0; c < p.nclass; ++c) ttot += p.ntaps[c];
    const int total = ttot * CI * cog * COT;
    for (int i = threadIdx.x; i < total; i += SC_THREADS) {
      const int ci = i % CI;
      const int r = i / CI;
      const int co = r % (cog * COT), tt = r / (cog * COT);
      const int t = tap_of(p, tt);
      const float v = co < p.Co ? __bfloat162float(p.w[((size_t)p.slab[t] * p.co_pad + co) * p.ci_pad + ci]) : 0.f;
      sw[((tt * CI + ci) * HV + (co % COT) / 4) * cog * 4 + (co / COT) * 4 + co % 4] = v;
    }
  }
  __syncthreads();
  for (unsigned item = blockIdx.x * SC_THREADS + threadIdx.x; item < p.items; item += gridDim.x * SC_THREADS) {
    const int g = item % cog;
    unsigned r = item / cog;
    const int q = r % p.qx; r /= p.qx;
    const int gy = r % p.Hg; r /= p.Hg;
    const int cls = r % p.nclass;
    const int n = r / p.nclass;
    const int gx0 = q * PT;
    float2 acc[PT][COT / 2];
#pragma unroll
    for (int j = 0; j < PT; ++j)
#pragma unroll
      for (int e = 0; e < COT / 2; ++e) acc[j][e] = make_float2(0.f, 0.f);
    const int nt = p.ntaps[cls], tt0 = p.tap0[cls];
    for (int k = 0; k < nt; ++k) {
      const int t = cls * 16 + k;
      const int iy = gy * p.is_ + p.dy[t];
      if (iy < 0 || iy >= p.Hi) continue;
      const int ixb = gx0 * p.is_ + p.dx[t];
      const bf16* rowp = p.in + (unsigned)((n * p.Hi + iy) * p.Wi) * (unsigned)p.ld_in;   // (< 2^31 elements: host check)
      const float* wt = sw + (tt0 + k) * CI * HV * cog * 4 + g * 4;
      uint4 u[PT][NCH];
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        const int ix = ixb + j * p.is_;
        const bool ok = ix >= 0 && ix < p.Wi;
#pragma unroll
        for (int h = 0; h < NCH; ++h)
          u[j][h] = ok ? __ldg(reinterpret_cast<const uint4*>(rowp + ix * p.ld_in + h * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + (ci * HV) * cog * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (ci * HV + 1) * cog * 4);
        const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const uint4& uu = u[j][ci / 8];
          const uint32_t word = ((ci % 8) / 2 == 0) ? uu.x : ((ci % 8) / 2 == 1) ? uu.y : ((ci % 8) / 2 == 2) ? uu.z : uu.w;
          const float xv = (ci & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
          const float2 x2 = make_float2(xv, xv);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[j][e] = __ffma2_rn(wv[e], x2, acc[j][e]);
        }
      }
    }
    const int co0 = g * COT;
    float bv[COT];
#pragma unroll
    for (int e = 0; e < COT; ++e) bv[e] = (p.bias && co0 + e < p.Co) ? __ldg(p.bias + co0 + e) : 0.f;
    const int oy = gy * p.os_ + p.oy0[cls];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int gx = gx0 + j, ox = gx * p.os_ + p.ox0[cls];
      if (gx >= p.Wg || oy >= p.Ho || ox >= p.Wo) continue;
      float f[COT];
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = acc[j][e].x; f[2 * e + 1] = acc[j][e].y; }
      epilogue<COT>(p, ((size_t)n * p.Ho + oy) * p.Wo + ox, co0, f, bv);
    }
  }
}

// ---- narrow output: V = 4, 8 or 16 output channels (all of them) per thread, LC = Ci/8 lanes per position quad -----------
template <int V>
__global__ void __launch_bounds__(SC_THREADS, 2) k_sc_conv_out(const __grid_constant__ ScParams p) {
  extern __shared__ __align__(16) float sw[];
  constexpr int WS = 8 * V + 4;   // floats per (tap, channel chunk): padded so the LC lanes hit different banks
  const int lc = p.lc;
  // sw[(tt*lc + c8)*WS + ci8*V + co] = W[tap tt][co][c8*8 + ci8]
  {
    int ttot = 0;
    for (int c = 0; c < p.nclass; ++c) ttot += p.ntaps[c];
    const int total = ttot * p.Ci * V;
    for (int i = threadIdx.x; i < total; i += SC_THREADS) {
      const int ci = i % p.Ci;
      const int r = i / p.Ci;
      const int co = r % V, tt = r / V;
      const int t = tap_of(p, tt);
      const float v = co < p.Co ? __bfloat162float(p.w[((size_t)p.slab[t] * p.co_pad + co) * p.ci_pad + ci]) : 0.f;
      sw[(tt * lc + ci / 8) * WS + (ci % 8) * V + co] = v;
    }
  }
  __syncthreads();
  // items = quads * lc; the trip count is warp-uniform (lc divides 32 and p.items is a multiple of lc), so the
  // shuffles below always see whole groups
  const unsigned items_pad = (p.items + 31u) & ~31u;
  for (unsigned item = blockIdx.x * SC_THREADS + threadIdx.x; item < items_pad; item += gridDim.x * SC_THREADS) {
    const bool live = item < p.items;
    const int c8 = item % lc;
    unsigned r = item / lc;
    const int q = r % p.qx; r /= p.qx;
    const int gy = r % p.Hg; r /= p.Hg;
    const int cls = r % p.nclass;
    const int n = live ? r / p.nclass : 0;
    const int gx0 = q * PT;
    float2 acc[PT][V / 2];
#pragma unroll
    for (int j = 0; j < PT; ++j)
#pragma unroll
      for (int e = 0; e < V / 2; ++e) acc[j][e] = make_float2(0.f, 0.f);
    const int nt = live ? p.ntaps[cls] : 0, tt0 = p.tap0[cls];
    for (int k = 0; k < nt; ++k) {
      const int t = cls * 16 + k;
      const int iy = gy * p.is_ + p.dy[t];
      if (iy < 0 || iy >= p.Hi) continue;
      const int ixb = gx0 * p.is_ + p.dx[t];
      const bf16* rowp = p.in + (unsigned)((n * p.Hi + iy) * p.Wi) * (unsigned)p.ld_in + c8 * 8;
      const float* wt = sw + ((tt0 + k) * lc + c8) * WS;
      uint4 u[PT];
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        const int ix = ixb + j * p.is_;
        u[j] = (ix >= 0 && ix < p.Wi) ? __ldg(reinterpret_cast<const uint4*>(rowp + ix * p.ld_in)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        float2 wv[V / 2];
#pragma unroll
        for (int h = 0; h < V / 4; ++h) {
          const float4 w4 = *reinterpret_cast<const float4*>(wt + ci * V + h * 4);
          wv[2 * h] = make_float2(w4.x, w4.y);
          wv[2 * h + 1] = make_float2(w4.z, w4.w);
        }
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const uint32_t word = (ci / 2 == 0) ? u[j].x : (ci / 2 == 1) ? u[j].y : (ci / 2 == 2) ? u[j].z : u[j].w;
          const float xv = (ci & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
          const float2 x2 = make_float2(xv, xv);
#pragma unroll
          for (int e = 0; e < V / 2; ++e) acc[j][e] = __ffma2_rn(wv[e], x2, acc[j][e]);
        }
      }
    }
    // combine the lc channel-chunk lanes of the quad (butterfly: every lane ends with the full sums)
    for (int o = 1; o < lc; o <<= 1) {
#pragma unroll
      for (int j = 0; j < PT; ++j)
#pragma unroll
        for (int e = 0; e < V / 2; ++e) {
          acc[j][e].x += __shfl_xor_sync(0xffffffffu, acc[j][e].x, o);
          acc[j][e].y += __shfl_xor_sync(0xffffffffu, acc[j][e].y, o);
        }
    }
    if (!live) continue;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = (p.bias && e < p.Co) ? __ldg(p.bias + e) : 0.f;
    const int oy = gy * p.os_ + p.oy0[cls];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      if (c8 != (j & (lc - 1))) continue;   // lane j of the group finishes position j (lc = 2: lanes 0/1 take two each)
      const int gx = gx0 + j, ox = gx * p.os_ + p.ox0[cls];
      if (gx >= p.Wg || oy >= p.Ho || ox >= p.Wo) continue;
      float f[V];
#pragma unroll
      for (int e = 0; e < V / 2; ++e) { f[2 * e] = acc[j][e].x; f[2 * e + 1] = acc[j][e].y; }
      epilogue<V>(p, ((size_t)n * p.Ho + oy) * p.Wo + ox, 0, f, bv);
    }
  }
}

template <typename K>
int launch_k(K kern, const ScParams& p, size_t smem, cudaStream_t s, bool* attr) {
  if (!*attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) { set_error("sc_conv smem attr: %s", cudaGetErrorString(e)); return 1; }
    *attr = true;
  }
  long long blocks = ((long long)p.items + SC_THREADS - 1) / SC_THREADS;
  const long long cap = 148LL * 6;       // persistent-ish: the weight staging is amortised over the grid-stride loop
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, SC_THREADS, smem, s>>>(p);
  return DS_LAUNCHED("sc_conv");
}

// -----------------------------------------------------------------------------------------------------------------
// weight gradient: dW[tap_off[t] + m*s_g + n*s_x] += sum_{img,y,x} G[img,y,x,m] * X[img, y*xs + dy[t], x*xs + dx[t], n]
// One side has 1/3/6/12 channels (the "narrow" side, whole pixel in registers), the other is read in 8-channel groups.
struct ScWgParams {
  int N, Hg, Wg, Cg, ld_g, Hx, Wx, Cx, ld_x, xs, ntaps;
  int dy[16], dx[16];
  long long tap_off[16];
  long long s_g, s_x;
  int wide_is_g;       // 1: G is the wide side (roles over G's channel groups), 0: X is
  int wgroups, roles, pl;
  unsigned npos, chunk;     // grid positions in total / per block
  const bf16* G; const bf16* X; float* dW;
};

// NC = channels of the narrow side
template <int NC>
__global__ void __launch_bounds__(SC_THREADS, 2) k_sc_wgrad(const __grid_constant__ ScWgParams p) {
  extern __shared__ __align__(16) float sacc[];   // [roles][NC*8 + 1]
  constexpr int NS8 = (NC + 7) / 8, RS = NC * 8 + 1;
  const int role = threadIdx.x % p.roles, lane = threadIdx.x / p.roles;
  for (int i = threadIdx.x; i < p.roles * RS; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int t = role / p.wgroups, wg = role % p.wgroups;
  const unsigned q0 = blockIdx.x * p.chunk, q1 = min(q0 + p.chunk, p.npos);
  float2 acc[NC][4];
#pragma unroll
  for (int s = 0; s < NC; ++s)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[s][e] = make_float2(0.f, 0.f);
  if (lane < p.pl) {
    const int dy = p.dy[t], dx = p.dx[t];
    constexpr int UB = NC > 6 ? 2 : 4;   // positions in flight per thread
    // (gx, gy, n) of the thread's first position; advanced incrementally by pl positions
    unsigned q = q0 + lane;
    int gx = q % p.Wg, gy = (q / p.Wg) % p.Hg, n = q / (p.Wg * p.Hg);
    while (q < q1) {
      uint4 wu[UB], nu[UB][NS8];
#pragma unroll
      for (int b = 0; b < UB; ++b) {
        const int iy = gy * p.xs + dy, ix = gx * p.xs + dx;
        const bool ok = q < q1 && iy >= 0 && iy < p.Hx && ix >= 0 && ix < p.Wx;
        wu[b] = make_uint4(0, 0, 0, 0);   // a zero wide operand cancels the position
#pragma unroll
        for (int h = 0; h < NS8; ++h) nu[b][h] = make_uint4(0, 0, 0, 0);
        if (ok) {
          const bf16* gp = p.G + q * (unsigned)p.ld_g;
          const bf16* xp = p.X + (unsigned)((n * p.Hx + iy) * p.Wx + ix) * (unsigned)p.ld_x;
          const bf16* wp = p.wide_is_g ? gp : xp;
          const bf16* np_ = p.wide_is_g ? xp : gp;
          wu[b] = __ldg(reinterpret_cast<const uint4*>(wp + wg * 8));
#pragma unroll
          for (int h = 0; h < NS8; ++h) nu[b][h] = __ldg(reinterpret_cast<const uint4*>(np_ + h * 8));
        }
        q += p.pl; gx += p.pl;
        while (gx >= p.Wg) { gx -= p.Wg; if (++gy == p.Hg) { gy = 0; ++n; } }
      }
#pragma unroll
      for (int b = 0; b < UB; ++b) {
        const float2 wv[4] = {bf2f(wu[b].x), bf2f(wu[b].y), bf2f(wu[b].z), bf2f(wu[b].w)};
#pragma unroll
        for (int s = 0; s < NC; ++s) {
          const uint4& uu = nu[b][s / 8];
          const uint32_t word = ((s % 8) / 2 == 0) ? uu.x : ((s % 8) / 2 == 1) ? uu.y : ((s % 8) / 2 == 2) ? uu.z : uu.w;
          const float sv = (s & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
          const float2 s2 = make_float2(sv, sv);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[s][e] = __ffma2_rn(wv[e], s2, acc[s][e]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NC; ++s)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(&sacc[role * RS + s * 8 + 2 * e], acc[s][e].x);
        atomicAdd(&sacc[role * RS + s * 8 + 2 * e + 1], acc[s][e].y);
      }
  }
  __syncthreads();
  const int ncw = p.wide_is_g ? p.Cg : p.Cx;
  for (int i = threadIdx.x; i < p.roles * NC * 8; i += blockDim.x) {
    const int e = i % 8, s = (i / 8) % NC, ro = i / (8 * NC);
    const int tt = ro / p.wgroups, wgc = (ro % p.wgroups) * 8 + e;
    if (wgc >= ncw) continue;
    const long long m = p.wide_is_g ? wgc : s, nn = p.wide_is_g ? s : wgc;
    atomicAdd(p.dW + p.tap_off[tt] + m * p.s_g + nn * p.s_x, sacc[ro * RS + s * 8 + e]);
  }
}
}  // namespace

namespace dsgan {
namespace sc {

bool conv_try(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out, void* pre_out,
              const void* aux, void* stream, int* rc) {
  const bool narrow_in = (d->Ci == 1 || d->Ci == 3 || d->Ci == 6 || d->Ci == 12) && d->Co <= 256;
  const bool narrow_out = !narrow_in && d->Co <= 16 && (d->Ci == 16 || d->Ci == 32 || d->Ci == 64 || d->Ci == 128);
  if (!narrow_in && !narrow_out) return false;
  if (d->Ci == 1 && d->Co > 64) return false;   // (PatchGAN head's input-gradient 1 -> 256: measured faster on the tensor cores)
  // pointwise 64 -> 12 (block c1's pwconv2 input-gradient): eight lanes per pixel spend as long in the shuffle reduction as in
  // the FMAs; measured 0.275 ms here against 0.129 ms as a one-tap tcgen05 implicit GEMM with the narrow epilogue
  if (narrow_out && d->nclass == 1 && d->ntaps[0] == 1 && d->Ci >= 64) return false;
  const long long quads = (long long)d->N * d->nclass * d->Hg * ((d->Wg + PT - 1) / PT);
  const int V = narrow_out ? (d->Co <= 4 ? 4 : (d->Co <= 8 ? 8 : 16)) : 8;   // channels per output vector
  const int cog = narrow_out ? 1 : (d->Co + 7) / 8;
  const int lc = narrow_out ? d->Ci / 8 : 1;
  if (quads * (narrow_out ? lc : cog) >= (1LL << 31)) return false;
  if ((long long)d->N * d->Hi * d->Wi * d->ld_in >= (1LL << 31)) return false;   // 32-bit element offsets in the kernels
  // pixel pitches: inputs are read in 16-byte chunks, outputs written in V-wide vectors (ragged Co only on a
  // whole-pitch tensor, where the extra lanes are padding)
  const int ci8 = (d->Ci + 7) / 8 * 8, cp8 = (d->Co + 7) / 8 * 8;
  if (d->ld_in % 8 || d->ld_in < ci8 || (uintptr_t)in % 16) return false;
  const int valign = V == 4 ? 4 : 8;
  auto out_ok = [&](const void* ptr, int ld) {
    if (!ptr) return true;
    if ((uintptr_t)ptr % (2 * valign) || ld % valign) return false;
    return d->Co % V == 0 || ld == cp8;
  };
  if (!out_ok(out, d->ld_out) || !out_ok(pre_out, d->ld_pre) || !out_ok(aux, d->ld_aux)) return false;
  int ttot = 0;
  for (int c = 0; c < d->nclass; ++c) ttot += d->ntaps[c];
  const size_t smem = narrow_out ? (size_t)ttot * lc * (8 * V + 4) * sizeof(float)
                                 : (size_t)ttot * d->Ci * cog * 8 * sizeof(float);
  if (smem > 96 * 1024) return false;
  ScParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Hi = d->Hi; p.Wi = d->Wi; p.Ho = d->Ho; p.Wo = d->Wo;
  p.Ci = d->Ci; p.Co = d->Co; p.co_pad = d->co_pad; p.ci_pad = d->ci_pad;
  p.ld_in = d->ld_in; p.ldc = d->ld_out; p.ld_pre = d->ld_pre; p.ld_aux = d->ld_aux;
  p.is_ = d->in_stride; p.os_ = d->out_stride; p.nclass = d->nclass; p.cog = cog; p.lc = lc; p.qx = (d->Wg + PT - 1) / PT;
  int t0 = 0;
  for (int c = 0; c < d->nclass; ++c) {
    p.oy0[c] = d->oy0[c]; p.ox0[c] = d->ox0[c]; p.ntaps[c] = d->ntaps[c]; p.tap0[c] = t0; t0 += d->ntaps[c];
    for (int t = c * 16; t < c * 16 + d->ntaps[c]; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.slab[t] = d->slab[t]; }
  }
  p.act = d->act; p.dact = d->dact; p.accumulate = d->accumulate;
  p.items = (unsigned)(quads * (narrow_out ? lc : cog));
  p.in = (const bf16*)in; p.w = (const bf16*)w_slabs; p.out = (bf16*)out; p.bias = bias; p.pre = (bf16*)pre_out;
  p.aux = (const bf16*)aux;
  cudaStream_t s = (cudaStream_t)stream;
  static bool a_in[4] = {false, false, false, false}, a_out[3] = {false, false, false};
  if (narrow_out) {
    if (V == 4) *rc = launch_k(k_sc_conv_out<4>, p, smem, s, &a_out[0]);
    else if (V == 8) *rc = launch_k(k_sc_conv_out<8>, p, smem, s, &a_out[1]);
    else *rc = launch_k(k_sc_conv_out<16>, p, smem, s, &a_out[2]);
    return true;
  }
  switch (d->Ci) {
    case 1: *rc = launch_k(k_sc_conv_in<1>, p, smem, s, &a_in[0]); return true;
    case 3: *rc = launch_k(k_sc_conv_in<3>, p, smem, s, &a_in[1]); return true;
    case 6: *rc = launch_k(k_sc_conv_in<6>, p, smem, s, &a_in[2]); return true;
    case 12: *rc = launch_k(k_sc_conv_in<12>, p, smem, s, &a_in[3]); return true;
  }
  return false;
}

bool wgrad_try(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream, int* rc) {
  auto narrow = [](int c) { return c == 1 || c == 3 || c == 6 || c == 12; };
  const bool ng = narrow(d->Cg), nx = narrow(d->Cx);
  if (!ng && !nx) return false;
  if (d->ntaps == 1) return false;              // pointwise weight gradients: measured faster on the tensor cores
  // the wide side is the one with more channels; it is read in 8-channel groups
  const bool wide_is_g = nx && (!ng || d->Cg >= d->Cx);
  const int cw = wide_is_g ? d->Cg : d->Cx, cn = wide_is_g ? d->Cx : d->Cg;
  const int ldw = wide_is_g ? d->ld_g : d->ld_x, ldn = wide_is_g ? d->ld_x : d->ld_g;
  const int cn8 = (cn + 7) / 8 * 8, cw8 = (cw + 7) / 8 * 8;
  if (ldw % 8 || ldn % 8 || ldw < cw8 || ldn < cn8 || (uintptr_t)G % 16 || (uintptr_t)X % 16) return false;
  const int wgroups = cw8 / 8, roles = wgroups * d->ntaps;
  if (roles > 128) return false;          // (e.g. the 256 -> 1 PatchGAN head: stays on the tensor cores)
  const long long npos = (long long)d->N * d->Hg * d->Wg;
  if (npos * d->ld_g >= (1LL << 31) || (long long)d->N * d->Hx * d->Wx * d->ld_x >= (1LL << 31)) return false;
  ScWgParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Cg = d->Cg; p.ld_g = d->ld_g; p.Hx = d->Hx; p.Wx = d->Wx; p.Cx = d->Cx;
  p.ld_x = d->ld_x; p.xs = d->x_stride; p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.tap_off[t] = d->tap_off[t]; }
  p.s_g = d->s_g; p.s_x = d->s_x; p.wide_is_g = wide_is_g ? 1 : 0; p.wgroups = wgroups; p.roles = roles;
  p.pl = SC_THREADS / roles;
  const int threads = (roles * p.pl + 31) / 32 * 32;
  long long blocks = 148 * 4;
  long long chunk = (npos + blocks - 1) / blocks;
  if (chunk < 8LL * p.pl) chunk = 8LL * p.pl;
  blocks = (npos + chunk - 1) / chunk;
  p.chunk = (unsigned)chunk; p.npos = (unsigned)npos;
  p.G = (const bf16*)G; p.X = (const bf16*)X; p.dW = dW;
  const size_t smem = (size_t)roles * (cn * 8 + 1) * sizeof(float);
  if (smem > 48 * 1024) return false;
  cudaStream_t s = (cudaStream_t)stream;
  switch (cn) {
    case 1: k_sc_wgrad<1><<<(unsigned)blocks, threads, smem, s>>>(p); break;
    case 3: k_sc_wgrad<3><<<(unsigned)blocks, threads, smem, s>>>(p); break;
    case 6: k_sc_wgrad<6><<<(unsigned)blocks, threads, smem, s>>>(p); break;
    default: k_sc_wgrad<12><<<(unsigned)blocks, threads, smem, s>>>(p); break;
  }
  *rc = DS_LAUNCHED("sc_wgrad");
  return true;
}

}  // namespace sc
}  // namespace dsgan
