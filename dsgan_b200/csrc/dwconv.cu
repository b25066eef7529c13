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
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < K * K * cb; i += 256) {
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
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K + 1) * VEC_CB; i += 256) (&sacc[0][0])[i] = 0.f;
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
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K + 1) * cb; i += 256) {
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
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K * K + 1) * CB; i += 256) (&sacc[0][0])[i] = 0.f;
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
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (K * K + 1) * cb; i += 256) {
    const int t = i / cb, c = i % cb;
    if (t < K * K) atomicAdd(dw + (size_t)(c_base + c) * K * K + t, sacc[t][c]);
    else if (db) atomicAdd(db + c_base + c, sacc[K * K][c]);
  }
}

inline bool vec_ok(const void* a, int lda, const void* b, int ldb, int C) {
  return C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}
}  // namespace

extern "C" {
int dsgan_dwconv_fwd(const void* x, int ld_x, const float* w, const float* bias, void* y, int ld_y, int dtype, int N,
                     int H, int W, int C, int k, int flip, int accumulate, void* stream) {
  const long long total = (long long)N * H * W * C;
  cudaStream_t s = (cudaStream_t)stream;
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
}
