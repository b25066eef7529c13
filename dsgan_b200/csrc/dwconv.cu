// Depthwise k x k convolution (k = 3,5,7,9; stride 1; pad k/2) — CUDA-core, HBM/L2-bound.
// Reference op sites: Block.dwconv (MixConvNeXtML.py:220) and MidMLKA.X3/X5/X7/X9 (:94-97).
#include "common.cuh"
#include "dwconv_mma.cuh"
#include "../../include/dsgan_b200.h"
#include <stdlib.h>
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

// ---------------------------------------------------------------------------------------------------------------
// bf16 fast path: each thread owns 8 channels (one 16-byte vector) and a strip of TW output pixels along x, sliding the
// K-wide window in registers.  Threads of a warp cover consecutive channel groups -> 16 B x 32 = 512 B coalesced rows.
// Weights of the block's channels are staged in shared memory as [tap][channel] fp32.
constexpr int TW = 8;
constexpr int VEC_CB = 128;  // channels per block (16 groups of 8)

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

// grid: (strip blocks, channel blocks); block: (groups, strips) with groups*strips = 256
template <int K>
__global__ void __launch_bounds__(256) k_dwconv_v8(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                    const float* __restrict__ bias, bf16* __restrict__ y, int ldy, int N,
                                                    int H, int W, int C, int flip, int acc_out) {
  __shared__ float sw[K * K][VEC_CB];
  constexpr int P = K / 2;
  const int c_base = blockIdx.y * VEC_CB;
  const int cb = min(VEC_CB, C - c_base);
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < K * K * cb; i += blockDim.x * blockDim.y) {
    const int tap = i / cb, c = i % cb;
    sw[tap][c] = __ldg(w + (size_t)(c_base + c) * K * K + (flip ? (K * K - 1 - tap) : tap));
  }
  __syncthreads();
  const int cg = threadIdx.x;  // channel group inside the block
  if (cg * 8 >= cb) return;
  const int c0 = c_base + cg * 8;
  const int strips_x = (W + TW - 1) / TW;
  const long long strip = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (strip >= (long long)N * H * strips_x) return;
  const int sx = (int)(strip % strips_x);
  const int yy = (int)((strip / strips_x) % H);
  const int n = (int)(strip / ((long long)strips_x * H));
  const int x0 = sx * TW;
  // packed fp32x2 FMAs (sm_100 FFMA2): two channels per instruction halve the issue count of this issue-bound kernel
  float2 acc[TW][4];
#pragma unroll
  for (int i = 0; i < TW; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = make_float2(0.f, 0.f);
  const bf16* xb = x + (size_t)n * H * W * ldx + c0;
  const bool x_interior = (x0 - P >= 0) && (x0 + TW + P <= W);
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    const int iy = yy + ky - P;
    if (iy < 0 || iy >= H) continue;
    float2 wr[K][4];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 a = *reinterpret_cast<const float4*>(&sw[ky * K + kx][cg * 8]);
      const float4 b = *reinterpret_cast<const float4*>(&sw[ky * K + kx][cg * 8 + 4]);
      wr[kx][0] = make_float2(a.x, a.y); wr[kx][1] = make_float2(a.z, a.w);
      wr[kx][2] = make_float2(b.x, b.y); wr[kx][3] = make_float2(b.z, b.w);
    }
    const bf16* row = xb + (size_t)iy * W * ldx;
    uint4 rv[TW + K - 1];  // all loads of the row are issued before any is consumed (memory-level parallelism)
    if (x_interior) {      // no per-load bounds predicates away from the left/right image border
      const bf16* r0 = row + (size_t)(x0 - P) * ldx;
#pragma unroll
      for (int xi = 0; xi < TW + K - 1; ++xi) rv[xi] = __ldg(reinterpret_cast<const uint4*>(r0 + (size_t)xi * ldx));
    } else {
#pragma unroll
      for (int xi = 0; xi < TW + K - 1; ++xi) {
        const int ix = x0 + xi - P;
        rv[xi] = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const uint4*>(row + (size_t)ix * ldx)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int xi = 0; xi < TW + K - 1; ++xi) {
      const uint32_t w4[4] = {rv[xi].x, rv[xi].y, rv[xi].z, rv[xi].w};
      float2 v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ox = xi - kx;
        if (ox >= 0 && ox < TW) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[ox][e] = __ffma2_rn(v[e], wr[kx][e], acc[ox][e]);
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bv[e] = bias ? __ldg(bias + c0 + e) : 0.f;
  bf16* yb = y + ((size_t)(n * H + yy) * W) * ldy + c0;
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    const int ox = x0 + i;
    if (ox >= W) break;
    float o[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { o[2 * e] = acc[i][e].x + bv[2 * e]; o[2 * e + 1] = acc[i][e].y + bv[2 * e + 1]; }
    uint4* dst = reinterpret_cast<uint4*>(yb + (size_t)ox * ldy);
    if (acc_out) {
      float old[8];
      unpack8(*dst, old);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += old[e];
    }
    *dst = pack8(o);
  }
}

// weight gradient: grid (strip chunks, channel blocks, K) — blockIdx.z = ky; each thread accumulates acc[kx][8]
template <int K>
__global__ void __launch_bounds__(256) k_dwconv_wgrad_v8(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                          int lddy, float* __restrict__ dw, float* __restrict__ db, int N,
                                                          int H, int W, int C, long long strips_per_block) {
  __shared__ float sacc[K + 1][VEC_CB];
  constexpr int P = K / 2;
  const int ky = blockIdx.z;
  const int c_base = blockIdx.y * VEC_CB;
  const int cb = min(VEC_CB, C - c_base);
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K + 1) * VEC_CB; i += blockDim.x * blockDim.y) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x;
  const bool active = cg * 8 < cb;
  const int c0 = c_base + cg * 8;
  const int strips_x = (W + TW - 1) / TW;
  const long long total = (long long)N * H * strips_x;
  const long long s_begin = (long long)blockIdx.x * strips_per_block;
  const long long s_end = min(s_begin + strips_per_block, total);
  float acc[K][8], accb[8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) accb[e] = 0.f;
  if (active) {
    for (long long strip = s_begin + threadIdx.y; strip < s_end; strip += blockDim.y) {
      const int sx = (int)(strip % strips_x);
      const int yy = (int)((strip / strips_x) % H);
      const int n = (int)(strip / ((long long)strips_x * H));
      const int x0 = sx * TW;
      const int iy = yy + ky - P;
      const bool row_ok = iy >= 0 && iy < H;
      if (!row_ok && ky != 0) continue;
      float g[TW][8];
      const bf16* gb = dy + ((size_t)(n * H + yy) * W) * lddy + c0;
#pragma unroll
      for (int i = 0; i < TW; ++i) {
        if (x0 + i < W) unpack8(__ldg(reinterpret_cast<const uint4*>(gb + (size_t)(x0 + i) * lddy)), g[i]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[i][e] = 0.f;
        }
      }
      if (ky == 0) {
#pragma unroll
        for (int i = 0; i < TW; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) accb[e] += g[i][e];
      }
      if (!row_ok) continue;
      const bf16* row = x + ((size_t)(n * H + iy) * W) * ldx + c0;
      uint4 rv[TW + K - 1];
#pragma unroll
      for (int xi = 0; xi < TW + K - 1; ++xi) {
        const int ix = x0 + xi - P;
        rv[xi] = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const uint4*>(row + (size_t)ix * ldx)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int xi = 0; xi < TW + K - 1; ++xi) {
        float v[8];
        unpack8(rv[xi], v);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ox = xi - kx;
          if (ox >= 0 && ox < TW) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[kx][e] = fmaf(v[e], g[ox][e], acc[kx][e]);
          }
        }
      }
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&sacc[kx][cg * 8 + e], acc[kx][e]);
    if (ky == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&sacc[K][cg * 8 + e], accb[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K + 1) * cb; i += blockDim.x * blockDim.y) {
    const int kx = i / cb, c = i % cb;
    if (kx < K) atomicAdd(dw + (size_t)(c_base + c) * K * K + ky * K + kx, sacc[kx][c]);
    else if (db && ky == 0) atomicAdd(db + c_base + c, sacc[K][c]);
  }
}


// weight gradient, single pass over the data: each thread owns TWO channels (one 32-bit load; a warp still covers
// 64 contiguous channels = 128 B) and ALL K*K taps (2*K*K fp32 accumulators), so dy and x are read once instead of K times.
template <int K>
__global__ void __launch_bounds__(256) k_dwconv_wgrad_v2(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                          int lddy, float* __restrict__ dw, float* __restrict__ db, int N,
                                                          int H, int W, int C, long long strips_per_block) {
  constexpr int CB = 128;  // channels per block (64 pairs)
  __shared__ float sacc[K * K + 1][CB];
  constexpr int P = K / 2;
  const int c_base = blockIdx.y * CB;
  const int cb = min(CB, C - c_base);
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K * K + 1) * CB; i += blockDim.x * blockDim.y) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  const int cp = threadIdx.x;
  const bool active = cp * 2 < cb;
  const int c0 = c_base + cp * 2;
  const int strips_x = (W + TW - 1) / TW;
  const long long total = (long long)N * H * strips_x;
  const long long s_begin = (long long)blockIdx.x * strips_per_block;
  const long long s_end = min(s_begin + strips_per_block, total);
  float2 acc[K * K];
  float accb[2] = {0.f, 0.f};
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);
  if (active) {
    for (long long strip = s_begin + threadIdx.y; strip < s_end; strip += blockDim.y) {
      const int sx = (int)(strip % strips_x);
      const int yy = (int)((strip / strips_x) % H);
      const int n = (int)(strip / ((long long)strips_x * H));
      const int x0 = sx * TW;
      const bool x_interior = (x0 - P >= 0) && (x0 + TW + P <= W);
      float2 g[TW];
      const bf16* gb = dy + ((size_t)(n * H + yy) * W) * lddy + c0;
#pragma unroll
      for (int i = 0; i < TW; ++i) {
        g[i] = (x0 + i < W) ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gb + (size_t)(x0 + i) * lddy))
                            : make_float2(0.f, 0.f);
        accb[0] += g[i].x; accb[1] += g[i].y;
      }
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = yy + ky - P;
        if (iy < 0 || iy >= H) continue;   // uniform across the warp's strips only at image borders
        const bf16* row = x + ((size_t)(n * H + iy) * W) * ldx + c0;
        uint32_t rv[TW + K - 1];
        if (x_interior) {
          const bf16* r0 = row + (size_t)(x0 - P) * ldx;
#pragma unroll
          for (int xi = 0; xi < TW + K - 1; ++xi) rv[xi] = __ldg(reinterpret_cast<const uint32_t*>(r0 + (size_t)xi * ldx));
        } else {
#pragma unroll
          for (int xi = 0; xi < TW + K - 1; ++xi) {
            const int ix = x0 + xi - P;
            rv[xi] = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const uint32_t*>(row + (size_t)ix * ldx)) : 0u;
          }
        }
#pragma unroll
        for (int xi = 0; xi < TW + K - 1; ++xi) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[xi]));
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const int ox = xi - kx;
            if (ox >= 0 && ox < TW) acc[ky * K + kx] = __ffma2_rn(v, g[ox], acc[ky * K + kx]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      atomicAdd(&sacc[t][cp * 2], acc[t].x);
      atomicAdd(&sacc[t][cp * 2 + 1], acc[t].y);
    }
    atomicAdd(&sacc[K * K][cp * 2], accb[0]);
    atomicAdd(&sacc[K * K][cp * 2 + 1], accb[1]);
  }
  __syncthreads();
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K * K + 1) * cb; i += blockDim.x * blockDim.y) {
    const int t = i / cb, c = i % cb;
    if (t < K * K) atomicAdd(dw + (size_t)(c_base + c) * K * K + t, sacc[t][c]);
    else if (db) atomicAdd(db + c_base + c, sacc[K * K][c]);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Shared-memory tiled bf16 kernels (the production path): a block stages an (8+K-1) x (32+K-1) x CB input tile with
// 16-byte cp.async (zero-filled outside the image), then every thread owns ONE channel pair (bf16x2 -> float2, FFMA2)
// and slides a K-wide register window along x, so each MAC pair costs ~1.4 instructions instead of ~5.
constexpr int T_TH = 8, T_TW = 32;
// acc += v * w on a channel pair.  Measured on B200 (scripts/micro/ffma2_rate.cu): with three DISTINCT 64-bit operands per
// instruction (the sliding-window pattern: a fresh input, a per-tap weight and a per-column accumulator) packed FFMA2 sustains
// 78 FMA/clk/SM, two scalar FFMAs 104 (both 125 when the multiplicands are shared) -- so the window sweep uses scalar FFMAs.
// (A persistent, double-buffered variant of k_dwconv_t that overlaps the staging of tile i+1 with the sweep of tile i was
// measured at the same 0.46 ms per 16x256x256x128 pass and was not kept.)
__device__ __forceinline__ float2 fma_pair(float2 v, float2 w, float2 acc) {
  return make_float2(fmaf(v.x, w.x, acc.x), fmaf(v.y, w.y, acc.y));
}

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// stage rows [y0, y0+rows) x cols [x0, x0+cols) x channels [c_base, c_base+CB) of image n into smem [rows][cols][CB]
template <int CB>
__device__ __forceinline__ void stage_tile(bf16* sm, const bf16* __restrict__ src, int ld, int n, int H, int W, int y0, int x0,
                                           int rows, int cols, int c_base, int C, int tid, int nthreads) {
  constexpr int OCT = CB / 8;
  const int total = rows * cols * OCT;
  for (int i = tid; i < total; i += nthreads) {
    const int o = i % OCT, px = (i / OCT) % cols, py = i / (OCT * cols);
    const int gy = y0 + py, gx = x0 + px, c = c_base + o * 8;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && c < C;
    const bf16* g = ok ? src + (((size_t)n * H + gy) * W + gx) * ld + c : src;
    cp_async16_zfill(sm + ((size_t)py * cols + px) * CB + o * 8, g, ok);
  }
}

// forward / input-gradient.  PB = channel pairs per block (CB = 2*PB channels); each thread: one pair, one tile row,
// XW = PB consecutive output columns.  grid: (tiles_x * tiles_y * N, channel blocks)
template <int K, int PB>
__global__ void __launch_bounds__(256) k_dwconv_t(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                   const float* __restrict__ bias, bf16* __restrict__ y, int ldy, int N, int H,
                                                   int W, int C, int flip, int acc_out) {
  constexpr int CB = 2 * PB, XW = PB, P = K / 2, IR = T_TH + K - 1, IC = T_TW + K - 1;
  extern __shared__ __align__(16) unsigned char dsm[];
  bf16* sx = reinterpret_cast<bf16*>(dsm);                       // [IR][IC][CB]
  float* sw = reinterpret_cast<float*>(dsm + (size_t)IR * IC * CB * 2);  // [K*K][CB]
  const int tid = threadIdx.x;
  const int tiles_x = (W + T_TW - 1) / T_TW, tiles_y = (H + T_TH - 1) / T_TH;
  const int t = blockIdx.x;
  const int n = t / (tiles_x * tiles_y), ty = (t / tiles_x) % tiles_y, tx = t % tiles_x;
  const int y0 = ty * T_TH, x0 = tx * T_TW, c_base = blockIdx.y * CB;
  stage_tile<CB>(sx, x, ldx, n, H, W, y0 - P, x0 - P, IR, IC, c_base, C, tid, 256);
  for (int i = tid; i < K * K * CB; i += 256) {
    const int tap = i / CB, c = i % CB;
    sw[i] = (c_base + c < C) ? __ldg(w + (size_t)(c_base + c) * K * K + (flip ? K * K - 1 - tap : tap)) : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  const int pair = tid % PB, lane = tid / PB;         // lane in [0, 256/PB)
  const int row = lane % T_TH, cgp = lane / T_TH;      // output row of the tile, column group
  const int xs = cgp * XW;                             // first output column of this thread inside the tile
  if (c_base + pair * 2 >= C) return;
  float2 acc[XW];
#pragma unroll
  for (int i = 0; i < XW; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    float2 wr[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) wr[kx] = *reinterpret_cast<const float2*>(&sw[(ky * K + kx) * CB + pair * 2]);
    const bf16* rowp = sx + ((size_t)(row + ky) * IC + xs) * CB + pair * 2;
#pragma unroll
    for (int xi = 0; xi < XW + K - 1; ++xi) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rowp + (size_t)xi * CB));
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ox = xi - kx;
        if (ox >= 0 && ox < XW) acc[ox] = fma_pair(v, wr[kx], acc[ox]);
      }
    }
  }
  const int gy = y0 + row;
  if (gy >= H) return;
  const int c0 = c_base + pair * 2;
  const float b0 = bias ? __ldg(bias + c0) : 0.f, b1 = (bias && c0 + 1 < C) ? __ldg(bias + c0 + 1) : 0.f;
  bf16* yb = y + (((size_t)n * H + gy) * W) * ldy + c0;
  if (acc_out) {
    // accumulate (gradient fan-in): fetch the old values in batches of 8 independent loads before touching them
#pragma unroll
    for (int i0 = 0; i0 < XW; i0 += 8) {
      uint32_t oldw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gx = x0 + xs + i0 + j;
        oldw[j] = (i0 + j < XW && gx < W) ? *reinterpret_cast<const uint32_t*>(yb + (size_t)gx * ldy) : 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (i0 + j < XW) {
          const float2 old = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&oldw[j]));
          acc[i0 + j].x += old.x; acc[i0 + j].y += old.y;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < XW; ++i) {
    const int gx = x0 + xs + i;
    if (gx >= W) break;
    *reinterpret_cast<__nv_bfloat162*>(yb + (size_t)gx * ldy) = __floats2bfloat162_rn(acc[i].x + b0, acc[i].y + b1);
  }
}

// weight gradient.  Thread = (channel pair, ky, row lane); slides along x with a K-wide window of dy.
// grid: (tile chunks, channel blocks); each block loops over `tiles_per_block` tiles before one reduction.
template <int K, int PB>
__global__ void __launch_bounds__(320) k_dwconv_wgrad_t(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                         int lddy, float* __restrict__ dw, float* __restrict__ db, int N,
                                                         int H, int W, int C, int tiles_per_block, int RL) {
  constexpr int CB = 2 * PB, P = K / 2, IR = T_TH + K - 1, IC = T_TW + K - 1;
  extern __shared__ __align__(16) unsigned char dsm[];
  bf16* sx = reinterpret_cast<bf16*>(dsm);                                  // [IR][IC][CB]
  bf16* sg = sx + (size_t)IR * IC * CB;                                     // [T_TH][T_TW][CB]
  float* sacc = reinterpret_cast<float*>(sg + (size_t)T_TH * T_TW * CB);    // [K*K+1][CB]
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < (K * K + 1) * CB; i += nthr) sacc[i] = 0.f;
  const int tiles_x = (W + T_TW - 1) / T_TW, tiles_y = (H + T_TH - 1) / T_TH;
  const int total_tiles = N * tiles_x * tiles_y;
  const int c_base = blockIdx.y * CB;
  const int pair = tid % PB, ky = (tid / PB) % K, rl = tid / (PB * K);
  const bool active = rl < RL && (c_base + pair * 2 < C);
  float2 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  const int t_begin = blockIdx.x * tiles_per_block, t_end = min(t_begin + tiles_per_block, total_tiles);
  for (int t = t_begin; t < t_end; ++t) {
    const int n = t / (tiles_x * tiles_y), ty = (t / tiles_x) % tiles_y, tx = t % tiles_x;
    const int y0 = ty * T_TH, x0 = tx * T_TW;
    __syncthreads();  // previous tile fully consumed
    stage_tile<CB>(sx, x, ldx, n, H, W, y0 - P, x0 - P, IR, IC, c_base, C, tid, nthr);
    stage_tile<CB>(sg, dy, lddy, n, H, W, y0, x0, T_TH, T_TW, c_base, C, tid, nthr);
    cp_async_wait_all();
    __syncthreads();
    if (active) {
      for (int r = rl; r < T_TH; r += RL) {
        const bf16* xr = sx + ((size_t)(r + ky) * IC) * CB + pair * 2;
        const bf16* gr = sg + ((size_t)r * T_TW) * CB + pair * 2;
        float2 gw[K];  // dy window: gw[j] = dy[xi - j]
#pragma unroll
        for (int j = 0; j < K; ++j) gw[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int xi = 0; xi < IC; ++xi) {
#pragma unroll
          for (int j = K - 1; j > 0; --j) gw[j] = gw[j - 1];
          gw[0] = (xi < T_TW) ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gr + (size_t)xi * CB))
                              : make_float2(0.f, 0.f);
          if (ky == 0 && xi < T_TW) { accb.x += gw[0].x; accb.y += gw[0].y; }
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + (size_t)xi * CB));
          // input column xi pairs with output column ox = xi - kx  ->  tap kx uses dy[xi - kx] = gw[kx]
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc[kx] = fma_pair(v, gw[kx], acc[kx]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      atomicAdd(&sacc[(ky * K + kx) * CB + pair * 2], acc[kx].x);
      atomicAdd(&sacc[(ky * K + kx) * CB + pair * 2 + 1], acc[kx].y);
    }
    if (ky == 0) {
      atomicAdd(&sacc[K * K * CB + pair * 2], accb.x);
      atomicAdd(&sacc[K * K * CB + pair * 2 + 1], accb.y);
    }
  }
  __syncthreads();
  const int cb = min(CB, C - c_base);
  for (int i = tid; i < (K * K + 1) * CB; i += nthr) {
    const int tap = i / CB, c = i % CB;
    if (c >= cb) continue;
    if (tap < K * K) atomicAdd(dw + (size_t)(c_base + c) * K * K + tap, sacc[i]);
    else if (db) atomicAdd(db + c_base + c, sacc[i]);
  }
}

template <int K, int PB>
int launch_dw_t(const bf16* x, int ldx, const float* w, const float* bias, bf16* y, int ldy, int N, int H, int W, int C,
                int flip, int acc, cudaStream_t s) {
  constexpr int CB = 2 * PB;
  const size_t smem = (size_t)(T_TH + K - 1) * (T_TW + K - 1) * CB * 2 + (size_t)K * K * CB * 4;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_dwconv_t<K, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid((unsigned)(N * cdiv(H, T_TH) * cdiv(W, T_TW)), (unsigned)cdiv(C, CB));
  k_dwconv_t<K, PB><<<grid, 256, smem, s>>>(x, ldx, w, bias, y, ldy, N, H, W, C, flip, acc);
  return DS_LAUNCHED("dwconv_t");
}
template <int K, int PB>
int launch_dwg_t(const bf16* x, int ldx, const bf16* dy, int lddy, float* dw, float* db, int N, int H, int W, int C,
                 cudaStream_t s) {
  constexpr int CB = 2 * PB;
  const size_t smem = ((size_t)(T_TH + K - 1) * (T_TW + K - 1) + (size_t)T_TH * T_TW) * CB * 2 + (size_t)(K * K + 1) * CB * 4;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_dwconv_wgrad_t<K, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  int RL = 320 / (PB * K);
  if (RL < 1) RL = 1;
  if (RL > T_TH) RL = T_TH;
  const int threads = (PB * K * RL + 31) / 32 * 32;
  const int total_tiles = N * cdiv(H, T_TH) * cdiv(W, T_TW), cblocks = cdiv(C, CB);
  int want = (3 * 148 + cblocks - 1) / cblocks;
  int tpb = (total_tiles + want - 1) / want;
  if (tpb < 1) tpb = 1;
  dim3 grid((unsigned)cdiv(total_tiles, tpb), (unsigned)cblocks);
  k_dwconv_wgrad_t<K, PB><<<grid, threads, smem, s>>>(x, ldx, dy, lddy, dw, db, N, H, W, C, tpb, RL);
  return DS_LAUNCHED("dwconv_wgrad_t");
}

inline bool vec_ok(const void* a, int lda, const void* b, int ldb, int C) {
  return C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}
// C < 8 on an 8-element pixel pitch (the 3-channel tensors of block c1): the tiled kernels stage whole 16-byte pixels and
// guard weights / outputs per channel, so the pad lanes are only ever copied, never mixed into real channels
inline bool small_ok(const void* a, int lda, const void* b, int ldb, int C) {
  return C < 8 && lda % 8 == 0 && ldb % 8 == 0 && lda >= 8 && ldb >= 8 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}
// DSGAN_DW_MMA=0 keeps the CUDA-core kernels (A/B measurements and the kernel tests that compare the two paths)
inline bool dw_mma_enabled() {
  const char* e = getenv("DSGAN_DW_MMA");
  return !(e && e[0] == '0');
}
}  // namespace

extern "C" {
int dsgan_dwconv_fwd(const void* x, int ld_x, const float* w, const float* bias, void* y, int ld_y, int dtype, int N,
                     int H, int W, int C, int k, int flip, int accumulate, void* stream) {
  const long long total = (long long)N * H * W * C;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DT_BF16 && dw_mma_enabled()) {
    int rc = 0;
    if (dwm::fwd_try((const bf16*)x, ld_x, w, bias, (bf16*)y, ld_y, N, H, W, C, k, flip, accumulate, s, &rc)) return rc;
  }
  if (dtype == DT_BF16 && ((vec_ok(x, ld_x, y, ld_y, C) && (C >= 64 || C == 32 || C == 16 || C == 8)) ||
                           small_ok(x, ld_x, y, ld_y, C))) {
    const bf16* xp = (const bf16*)x; bf16* yp = (bf16*)y;
#define DWT(KK, PB) return launch_dw_t<KK, PB>(xp, ld_x, w, bias, yp, ld_y, N, H, W, C, flip, accumulate, s);
#define DWT_K(PB) switch (k) { case 3: DWT(3, PB) case 5: DWT(5, PB) case 7: DWT(7, PB) case 9: DWT(9, PB) default: break; }
    if (C >= 64) { DWT_K(32) } else if (C == 32) { DWT_K(16) } else if (C == 16) { DWT_K(8) } else { DWT_K(4) }
#undef DWT_K
#undef DWT
    set_error("dwconv: unsupported k=%d", k);
    return 1;
  }
  if (dtype == DT_BF16 && vec_ok(x, ld_x, y, ld_y, C)) {
    const int groups = C >= VEC_CB ? VEC_CB / 8 : C / 8;
    dim3 block(groups, 256 / groups);
    const long long strips = (long long)N * H * ((W + TW - 1) / TW);
    dim3 grid(cdiv(strips, block.y), cdiv(C, VEC_CB));
#define DWV_CASE(KK)                                                                                             \
  case KK:                                                                                                       \
    k_dwconv_v8<KK><<<grid, block, 0, s>>>((const bf16*)x, ld_x, w, bias, (bf16*)y, ld_y, N, H, W, C, flip, accumulate); \
    break;
    switch (k) {
      DWV_CASE(3) DWV_CASE(5) DWV_CASE(7) DWV_CASE(9)
      default: set_error("dwconv: unsupported k=%d", k); return 1;
    }
#undef DWV_CASE
    return DS_LAUNCHED("dwconv_fwd_v8");
  }
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
  if (dtype == DT_BF16 && dw_mma_enabled()) {
    int rc = 0;
    if (dwm::wgrad_try((const bf16*)x, ld_x, (const bf16*)dy, ld_dy, dw, db, N, H, W, C, k, (cudaStream_t)stream, &rc)) return rc;
  }
  if (dtype == DT_BF16 && ((vec_ok(x, ld_x, dy, ld_dy, C) && (C >= 64 || C == 32 || C == 16 || C == 8)) ||
                           small_ok(x, ld_x, dy, ld_dy, C))) {
    const bf16* xp = (const bf16*)x; const bf16* gp = (const bf16*)dy;
    cudaStream_t s = (cudaStream_t)stream;
#define DWG(KK, PB) return launch_dwg_t<KK, PB>(xp, ld_x, gp, ld_dy, dw, db, N, H, W, C, s);
#define DWG_K(PB) switch (k) { case 3: DWG(3, PB) case 5: DWG(5, PB) case 7: DWG(7, PB) case 9: DWG(9, PB) default: break; }
    if (C >= 64) { DWG_K(32) } else if (C == 32) { DWG_K(16) } else if (C == 16) { DWG_K(8) } else { DWG_K(4) }
#undef DWG_K
#undef DWG
    set_error("dwconv_wgrad: unsupported k=%d", k);
    return 1;
  }
  if (dtype == DT_BF16 && C % 2 == 0 && ld_x % 2 == 0 && ld_dy % 2 == 0 && ((uintptr_t)x % 4 == 0) && ((uintptr_t)dy % 4 == 0)) {
    const int pairs = C >= 128 ? 64 : C / 2;
    dim3 block(pairs, 256 / pairs);
    const long long strips = (long long)N * H * ((W + TW - 1) / TW);
    const long long cblocks = cdiv(C, 128);
    long long want = (4LL * 148 + cblocks - 1) / cblocks;
    long long per_block = (strips + want - 1) / want;
    if (per_block < (long long)block.y) per_block = block.y;
    dim3 grid((unsigned)((strips + per_block - 1) / per_block), (unsigned)cblocks);
    cudaStream_t s = (cudaStream_t)stream;
#define DWG2_CASE(KK)                                                                                            \
  case KK:                                                                                                       \
    k_dwconv_wgrad_v2<KK><<<grid, block, 0, s>>>((const bf16*)x, ld_x, (const bf16*)dy, ld_dy, dw, db, N, H, W, C, per_block); \
    break;
    switch (k) {
      DWG2_CASE(3) DWG2_CASE(5) DWG2_CASE(7) DWG2_CASE(9)
      default: set_error("dwconv_wgrad: unsupported k=%d", k); return 1;
    }
#undef DWG2_CASE
    return DS_LAUNCHED("dwconv_wgrad_v2");
  }
  if (dtype == DT_BF16 && vec_ok(x, ld_x, dy, ld_dy, C)) {
    const int groups = C >= VEC_CB ? VEC_CB / 8 : C / 8;
    dim3 block(groups, 256 / groups);
    const long long strips = (long long)N * H * ((W + TW - 1) / TW);
    // ~4 waves of blocks over (strip chunks x channel blocks x k), at least one strip per thread row
    long long want = (4LL * 148 + (long long)cdiv(C, VEC_CB) * k - 1) / ((long long)cdiv(C, VEC_CB) * k);
    long long per_block = (strips + want - 1) / want;
    if (per_block < (long long)block.y) per_block = block.y;
    long long blocks_x = (strips + per_block - 1) / per_block;
    dim3 grid((unsigned)blocks_x, cdiv(C, VEC_CB), k);
    cudaStream_t s = (cudaStream_t)stream;
#define DWGV_CASE(KK)                                                                                            \
  case KK:                                                                                                       \
    k_dwconv_wgrad_v8<KK><<<grid, block, 0, s>>>((const bf16*)x, ld_x, (const bf16*)dy, ld_dy, dw, db, N, H, W, C, per_block); \
    break;
    switch (k) {
      DWGV_CASE(3) DWGV_CASE(5) DWGV_CASE(7) DWGV_CASE(9)
      default: set_error("dwconv_wgrad: unsupported k=%d", k); return 1;
    }
#undef DWGV_CASE
    return DS_LAUNCHED("dwconv_wgrad_v8");
  }
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

int dsgan_dwconv_multi_fwd(const void* x, int ld_x, void* y, int ld_y, int dtype, int N, int H, int W,
                           const dsgan_dw_branch* br, int nbr, int flip, int accumulate, void* stream) {
  DS_REQUIRE(br && nbr >= 1 && nbr <= 4, "dwconv_multi_fwd: 1..4 branches");
  const int es = dtype == DT_BF16 ? 2 : 4;
  if (dtype == DT_BF16 && dw_mma_enabled()) {
    int rc = 0;
    if (dwm::multi_fwd_try((const bf16*)x, ld_x, (bf16*)y, ld_y, N, H, W, br, nbr, flip, accumulate, (cudaStream_t)stream, &rc))
      return rc;
  }
  for (int i = 0; i < nbr; ++i) {   // shapes the one-launch kernel does not take: one call per branch
    const int rc = dsgan_dwconv_fwd((const char*)x + (size_t)br[i].c0 * es, ld_x, br[i].w, br[i].bias,
                                    (char*)y + (size_t)br[i].c0 * es, ld_y, dtype, N, H, W, br[i].c, br[i].k, flip, accumulate,
                                    stream);
    if (rc) return rc;
  }
  return 0;
}
int dsgan_dwconv_multi_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, int dtype, int N, int H, int W,
                             const dsgan_dw_branch* br, int nbr, void* stream) {
  DS_REQUIRE(br && nbr >= 1 && nbr <= 4, "dwconv_multi_wgrad: 1..4 branches");
  const int es = dtype == DT_BF16 ? 2 : 4;
  if (dtype == DT_BF16 && dw_mma_enabled()) {
    int rc = 0;
    if (dwm::multi_wgrad_try((const bf16*)x, ld_x, (const bf16*)dy, ld_dy, N, H, W, br, nbr, (cudaStream_t)stream, &rc)) return rc;
  }
  for (int i = 0; i < nbr; ++i) {
    const int rc = dsgan_dwconv_wgrad((const char*)x + (size_t)br[i].c0 * es, ld_x, (const char*)dy + (size_t)br[i].c0 * es,
                                      ld_dy, br[i].dw, br[i].db, dtype, N, H, W, br[i].c, br[i].k, stream);
    if (rc) return rc;
  }
  return 0;
}
}
