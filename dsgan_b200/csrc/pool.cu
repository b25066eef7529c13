// MaxPool2d(k) forward/backward and the CA channel-attention gate (global avg+max pooling, shared
// fc1->PReLU->fc2, sigmoid) forward/backward.  Reference: MixConvNeXtML.py:5-22,68-74,333-354; vgg.py pools.
#include "common.cuh"
#include <string.h>
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
template <typename T>
__global__ void k_maxpool_fwd(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int H, int W, int C,
                              int k, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Ho = H / k, Wo = W / k;
  const int c = (int)(i % C);
  const long long op = i / C;
  const int ox = (int)(op % Wo);
  const long long r = op / Wo;
  const int oy = (int)(r % Ho);
  const long long n = r / Ho;
  const T* xb = x + ((n * H + (long long)oy * k) * W + (long long)ox * k) * ldx + c;
  float m = -INFINITY;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) m = fmaxf(m, ldf(xb + ((long long)ky * W + kx) * ldx));
  stf(y + op * ldy + c, m);
}

template <typename T>
__global__ void k_maxpool_bwd(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy,
                              T* __restrict__ dx, int lddx, int H, int W, int C, int k, int acc, int relu_mask,
                              long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Ho = H / k, Wo = W / k;
  const int c = (int)(i % C);
  const long long op = i / C;
  const int ox = (int)(op % Wo);
  const long long r = op / Wo;
  const int oy = (int)(r % Ho);
  const long long n = r / Ho;
  const long long win = (n * H + (long long)oy * k) * W + (long long)ox * k;
  const T* xb = x + win * ldx + c;
  float m = -INFINITY;
  int am = 0;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const float v = ldf(xb + ((long long)ky * W + kx) * ldx);
      if (v > m) { m = v; am = ky * k + kx; }  // strict: first maximum in scan order (Q5)
    }
  const float g = ldf(dy + op * lddy + c);
  T* db = dx + win * lddx + c;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      T* o = db + ((long long)ky * W + kx) * lddx;
      float v = (ky * k + kx == am) ? g : 0.f;
      if (acc) v += ldf(o);
      if (relu_mask && !(ldf(xb + ((long long)ky * W + kx) * ldx) > 0.f)) v = 0.f;
      stf(o, v);
    }
}


// bf16 fast paths of MaxPool2d: one thread per (output pixel, 8-channel group)
__device__ __forceinline__ void unpack8p(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(h);
    f[2 * e + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8p(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// K = window (2/4/8/16): the window is scanned in row batches of UB = min(K, 8) independent 16-byte loads
template <int K>
__global__ void __launch_bounds__(256) k_maxpool_fwd_v8(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy,
                                                         int H, int W, int C, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  constexpr int UB = K < 8 ? K : 8;
  const int G = C / 8, Ho = H / K, Wo = W / K;
  const int c0 = (int)(i % G) * 8;
  const long long op = i / G;
  const int ox = (int)(op % Wo);
  const long long r = op / Wo;
  const int oy = (int)(r % Ho);
  const long long n = r / Ho;
  const bf16* xb = x + ((n * H + (long long)oy * K) * W + (long long)ox * K) * ldx + c0;
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll(K <= 4 ? K : 1)
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int kb = 0; kb < K; kb += UB) {
      uint4 u[UB];
#pragma unroll
      for (int j = 0; j < UB; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)ky * W + kb + j) * ldx));
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        float v[8];
        unpack8p(u[j], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    }
  *reinterpret_cast<uint4*>(y + op * ldy + c0) = pack8p(m);
}

template <int K>
__global__ void __launch_bounds__(256) k_maxpool_bwd_v8(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                         int lddy, bf16* __restrict__ dx, int lddx, int H, int W, int C,
                                                         int acc, int relu_mask, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  constexpr int UB = K < 8 ? K : 8;
  const int G = C / 8, Ho = H / K, Wo = W / K;
  const int c0 = (int)(i % G) * 8;
  const long long op = i / G;
  const int ox = (int)(op % Wo);
  const long long r = op / Wo;
  const int oy = (int)(r % Ho);
  const long long n = r / Ho;
  const long long win = (n * H + (long long)oy * K) * W + (long long)ox * K;
  const bf16* xb = x + win * ldx + c0;
  bf16* db = dx + win * lddx + c0;
  const uint4 gu = __ldg(reinterpret_cast<const uint4*>(dy + op * lddy + c0));
  float m[8];
  int am[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; am[e] = 0; }
  if constexpr (K == 2) {
    // whole window (and the accumulate operand) in registers: 4 + 4 independent loads, then 4 stores
    uint4 xu[4], ou[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) xu[t] = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)(t >> 1) * W + (t & 1)) * ldx));
    if (acc) {
#pragma unroll
      for (int t = 0; t < 4; ++t) ou[t] = *reinterpret_cast<const uint4*>(db + ((long long)(t >> 1) * W + (t & 1)) * lddx);
    }
    float xv[4][8], g[8];
    unpack8p(gu, g);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      unpack8p(xu[t], xv[t]);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (xv[t][e] > m[e]) { m[e] = xv[t][e]; am[e] = t; }  // strict: first maximum in scan order (Q5)
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float v[8], old[8];
      if (acc) unpack8p(ou[t], old);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = (am[e] == t) ? g[e] : 0.f;
        if (acc) v[e] += old[e];
        if (relu_mask && !(xv[t][e] > 0.f)) v[e] = 0.f;
      }
      *reinterpret_cast<uint4*>(db + ((long long)(t >> 1) * W + (t & 1)) * lddx) = pack8p(v);
    }
    return;
  }
#pragma unroll(K <= 4 ? K : 1)
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int kb = 0; kb < K; kb += UB) {
      uint4 u[UB];
#pragma unroll
      for (int j = 0; j < UB; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)ky * W + kb + j) * ldx));
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        float v[8];
        unpack8p(u[j], v);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v[e] > m[e]) { m[e] = v[e]; am[e] = ky * K + kb + j; }
      }
    }
  float g[8];
  unpack8p(gu, g);
#pragma unroll(K <= 4 ? K : 1)
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int kb = 0; kb < K; kb += UB) {
      uint4 ou[UB], xu[UB];
      if (acc) {
#pragma unroll
        for (int j = 0; j < UB; ++j) ou[j] = *reinterpret_cast<const uint4*>(db + ((long long)ky * W + kb + j) * lddx);
      }
      if (relu_mask) {
#pragma unroll
        for (int j = 0; j < UB; ++j) xu[j] = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)ky * W + kb + j) * ldx));
      }
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        float v[8], old[8], xv[8];
        if (acc) unpack8p(ou[j], old);
        if (relu_mask) unpack8p(xu[j], xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = (am[e] == ky * K + kb + j) ? g[e] : 0.f;
          if (acc) v[e] += old[e];
          if (relu_mask && !(xv[e] > 0.f)) v[e] = 0.f;
        }
        *reinterpret_cast<uint4*>(db + ((long long)ky * W + kb + j) * lddx) = pack8p(v);
      }
    }
}

__device__ __forceinline__ unsigned long long pack_max(float v, unsigned p) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - p);
}
__device__ __forceinline__ void unpack_max(unsigned long long k, float& v, int& p) {
  unsigned u = (unsigned)(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  v = __uint_as_float(u);
  p = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
}

// grid (pixel chunks, N)
template <typename T>
__global__ void k_ca_pool(const T* __restrict__ x, long long HW, int C, float* __restrict__ sum,
                          unsigned long long* __restrict__ packed, int chunk) {
  const int cl = C < 64 ? C : 64, pl = blockDim.x / cl;
  const int tc = threadIdx.x % cl, tp = threadIdx.x / cl;
  if (tp >= pl) return;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(p0 + (long long)chunk, HW);
  for (int c = tc; c < C; c += cl) {
    float s = 0.f, m = -INFINITY;
    unsigned am = 0;
    for (long long p = p0 + tp; p < p1; p += pl) {
      const float v = ldf(x + ((long long)n * HW + p) * C + c);
      s += v;
      if (v > m) { m = v; am = (unsigned)p; }
    }
    if (p0 + tp < p1) {
      atomicAdd(sum + (long long)n * C + c, s);
      atomicMax(packed + (long long)n * C + c, pack_max(m, am));
    }
  }
}

// bf16 fast path: 8 channels (16 bytes) per thread, 4 independent loads in flight, block-level reduction in shared memory
// (sum: float atomics, max: the packed (value, first index) 64-bit atomicMax), one global atomic pair per (block, channel).
// grid (pixel chunks, N, channel blocks of <= 256)
__global__ void __launch_bounds__(256) k_ca_pool_v8(const bf16* __restrict__ x, long long HW, int C, float* __restrict__ sum,
                                                     unsigned long long* __restrict__ packed, int chunk) {
  __shared__ float ssum[256];
  __shared__ unsigned long long smax[256];
  const int groups = C / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;
  const int tg = threadIdx.x % gl, tp = threadIdx.x / gl;
  const int n = blockIdx.y, cb = blockIdx.z * gl * 8, c0 = cb + tg * 8;
  ssum[threadIdx.x] = 0.f;
  smax[threadIdx.x] = 0ull;
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(p0 + (long long)chunk, HW);
  if (tp < pl && c0 < C && p0 + tp < p1) {
    float s[8], m[8];
    unsigned am[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; m[e] = -INFINITY; am[e] = 0; }
    const bf16* xb = x + (size_t)n * HW * C + c0;
    long long p = p0 + tp;
    for (; p + 3 * pl < p1; p += 4 * pl) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(xb + (p + (long long)j * pl) * C));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
        unpack8p(u[j], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[e] += v[e];
          if (v[e] > m[e]) { m[e] = v[e]; am[e] = (unsigned)(p + (long long)j * pl); }   // strict: first maximum in scan order
        }
      }
    }
    for (; p < p1; p += pl) {
      float v[8];
      unpack8p(__ldg(reinterpret_cast<const uint4*>(xb + p * C)), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s[e] += v[e];
        if (v[e] > m[e]) { m[e] = v[e]; am[e] = (unsigned)p; }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&ssum[tg * 8 + e], s[e]);
      atomicMax(&smax[tg * 8 + e], pack_max(m[e], am[e]));
    }
  }
  __syncthreads();
  if (threadIdx.x < gl * 8 && cb + threadIdx.x < C && smax[threadIdx.x] != 0ull) {
    atomicAdd(sum + (long long)n * C + cb + threadIdx.x, ssum[threadIdx.x]);
    atomicMax(packed + (long long)n * C + cb + threadIdx.x, smax[threadIdx.x]);
  }
}

// one block per sample.  smem: avg[C], mx[C], ha[Cr], hm[Cr]
__global__ void k_ca_mlp(const unsigned long long* __restrict__ packed, long long HW,
                         int C, const float* __restrict__ fc1, const float* __restrict__ slope,
                         const float* __restrict__ fc2, float* avg /* in: plane sums, out: means */,
                         float* __restrict__ mx,
                         int* __restrict__ argmax, float* __restrict__ s) {
  extern __shared__ float sm[];
  const int Cr = C / 8, n = blockIdx.x;
  float* a = sm; float* m = sm + C; float* ha = m + C; float* hm = ha + Cr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mv; int mp;
    unpack_max(packed[(long long)n * C + c], mv, mp);
    const float av = avg[(long long)n * C + c] / (float)HW;
    a[c] = av; m[c] = mv;
    avg[(long long)n * C + c] = av; mx[(long long)n * C + c] = mv; argmax[(long long)n * C + c] = mp;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float sl = slope[0];
  for (int j = w; j < Cr; j += nw) {
    float pa = 0.f, pm = 0.f;
    for (int c = lane; c < C; c += 32) { const float f = fc1[j * C + c]; pa = fmaf(f, a[c], pa); pm = fmaf(f, m[c], pm); }
    pa = warp_sum(pa); pm = warp_sum(pm);
    if (lane == 0) { ha[j] = pa; hm[j] = pm; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float z = 0.f;
    for (int j = 0; j < Cr; ++j) {
      const float pa = ha[j] > 0.f ? ha[j] : sl * ha[j], pm = hm[j] > 0.f ? hm[j] : sl * hm[j];
      z = fmaf(fc2[c * Cr + j], pa + pm, z);
    }
    s[(long long)n * C + c] = 1.0f / (1.0f + __expf(-z));
  }
}

__global__ void k_ca_bwd(const float* __restrict__ ds, const float* __restrict__ s, const float* __restrict__ avg,
                         const float* __restrict__ mx, int C, const float* __restrict__ fc1,
                         const float* __restrict__ slope, const float* __restrict__ fc2, float* __restrict__ dfc1,
                         float* __restrict__ dslope, float* __restrict__ dfc2, float* __restrict__ davg,
                         float* __restrict__ dmax) {
  extern __shared__ float sm[];
  const int Cr = C / 8, n = blockIdx.x;
  float* a = sm; float* m = a + C; float* dz = m + C; float* ha = dz + C; float* hm = ha + Cr;
  float* dha = hm + Cr; float* dhm = dha + Cr; float* red = dhm + Cr;  // red[32]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long o = (long long)n * C + c;
    a[c] = avg[o]; m[c] = mx[o];
    const float sv = s[o];
    dz[c] = ds[o] * sv * (1.f - sv);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float sl = slope[0];
  float dsl = 0.f;
  for (int j = w; j < Cr; j += nw) {
    float pa = 0.f, pm = 0.f, dp = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float f = fc1[j * C + c];
      pa = fmaf(f, a[c], pa); pm = fmaf(f, m[c], pm);
      dp = fmaf(fc2[c * Cr + j], dz[c], dp);
    }
    pa = warp_sum(pa); pm = warp_sum(pm); dp = warp_sum(dp);
    if (lane == 0) {
      ha[j] = pa; hm[j] = pm;
      dha[j] = dp * (pa > 0.f ? 1.f : sl);
      dhm[j] = dp * (pm > 0.f ? 1.f : sl);
      dsl += dp * ((pa > 0.f ? 0.f : pa) + (pm > 0.f ? 0.f : pm));
    }
  }
  dsl = block_sum(dsl, red);
  if (threadIdx.x == 0) atomicAdd(dslope, dsl);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float da = 0.f, dm = 0.f;
    for (int j = 0; j < Cr; ++j) {
      const float pa = ha[j] > 0.f ? ha[j] : sl * ha[j], pm = hm[j] > 0.f ? hm[j] : sl * hm[j];
      atomicAdd(dfc2 + c * Cr + j, dz[c] * (pa + pm));
      atomicAdd(dfc1 + j * C + c, dha[j] * a[c] + dhm[j] * m[c]);
      const float f = fc1[j * C + c];
      da = fmaf(f, dha[j], da); dm = fmaf(f, dhm[j], dm);
    }
    davg[(long long)n * C + c] = da; dmax[(long long)n * C + c] = dm;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-scale max-pooling of ONE tensor: MaxPool2d(2), (4), (8), (16) of the same encoder skip (downSkip*: MixConvNeXtML.py:
// 328-426, plus the encoder's own downSample :68-74 for k = 2) in one pass.  The separate kernels read R1 four times and
// read-modify-write its gradient four times; here x is read once and dx written once.
// A block owns 256 pixels = TB tiles of T x T (T = 2^NLEV) and up to 128 channels; level l is built from level l-1 in shared
// memory (max is associative).  Backward carries (value, position) pairs: MaxPool2d routes the gradient to the FIRST maximum
// in row-major scan order of each window, which inside a tile is the smallest code y*T + x among the maxima -- so merging the
// four children with "greater value, or equal value and smaller code" reproduces the scan-order rule exactly.
constexpr int MP_CH = 128;          // channels per block
constexpr int MP_G = MP_CH / 8;

struct MultiPoolParams {
  const bf16* x; int ldx;
  bf16* y[4];                      // outputs k = 2, 4, 8, 16 (dense, pitch C)
  const bf16* dy[4];               // backward: gradients of those outputs (null = no gradient flowed)
  bf16* dx; int lddx; int acc;
  int N, H, W, C;
  int tiles_x, tiles_y, total_tiles;
};

// 2x2 block b (0..63) of the block's pixel set -> tile slot, block coordinates inside the tile
template <int NLEV>
__device__ __forceinline__ void mp_locate(int b, int& slot, int& by, int& bx) {
  constexpr int T = 1 << NLEV, NB = (T / 2) * (T / 2);
  slot = b / NB;
  const int r = b % NB;
  by = r / (T / 2); bx = r % (T / 2);
}

template <int NLEV, bool BWD>
__global__ void __launch_bounds__(256) k_multipool(const MultiPoolParams p) {
  constexpr int T = 1 << NLEV, TB = 256 / (T * T);
  extern __shared__ __align__(16) unsigned char mp_smem[];
  // level arrays: windows per block at level l (1-based) = 256 >> (2 l); values bf16x8 per channel group, codes u8x8
  uint4* sval[4]; uint2* spos[4];
  {
    unsigned char* q = mp_smem;
    for (int l = 0; l < NLEV; ++l) { sval[l] = reinterpret_cast<uint4*>(q); q += (size_t)(64 >> (2 * l)) * MP_G * 16; }
    for (int l = 0; l < NLEV; ++l) { spos[l] = reinterpret_cast<uint2*>(q); q += (size_t)(64 >> (2 * l)) * MP_G * 8; }
  }
  const int c_base = blockIdx.y * MP_CH;
  const int G = min(MP_G, (p.C - c_base) / 8);
  const int tile0 = blockIdx.x * TB;
  // ---- level 1: 2x2 blocks straight from global memory ----
  for (int item = threadIdx.x; item < 64 * G; item += 256) {
    const int g = item % G, b = item / G;
    int slot, by, bx;
    mp_locate<NLEV>(b, slot, by, bx);
    const int tile = tile0 + slot;
    if (tile >= p.total_tiles) continue;
    const int n = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int y0 = ty * T + 2 * by, x0 = tx * T + 2 * bx, c0 = c_base + g * 8;
    const bf16* xb = p.x + (((size_t)n * p.H + y0) * p.W + x0) * p.ldx + c0;
    uint4 u[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) u[t] = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)(t >> 1) * p.W + (t & 1)) * p.ldx));
    float m[8];
    unsigned code[8];
    unpack8p(u[0], m);
#pragma unroll
    for (int e = 0; e < 8; ++e) code[e] = (2 * by) * T + 2 * bx;
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      float v[8];
      unpack8p(u[t], v);
      const unsigned cd = (2 * by + (t >> 1)) * T + 2 * bx + (t & 1);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (v[e] > m[e]) { m[e] = v[e]; code[e] = cd; }      // strict: the first maximum in scan order wins
    }
    const uint4 mv = pack8p(m);
    sval[0][b * MP_G + g] = mv;
    if (BWD) {
      spos[0][b * MP_G + g] = make_uint2(code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24),
                                         code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24));
    } else {
      const int Ho = p.H / 2, Wo = p.W / 2;
      *reinterpret_cast<uint4*>(p.y[0] + (((size_t)n * Ho + y0 / 2) * Wo + x0 / 2) * p.C + c0) = mv;
    }
  }
  __syncthreads();
  // ---- levels 2 .. NLEV from shared memory ----
#pragma unroll
  for (int l = 1; l < NLEV; ++l) {
    const int wpt = T >> (l + 1);                 // windows per tile side at this level
    const int cpt = T >> l;                       // children per tile side
    const int nwin = TB * wpt * wpt;
    for (int item = threadIdx.x; item < nwin * G; item += 256) {
      const int g = item % G, w = item / G;
      const int slot = w / (wpt * wpt), wy = (w % (wpt * wpt)) / wpt, wx = w % wpt;
      const int tile = tile0 + slot;
      if (tile >= p.total_tiles) continue;
      float m[8];
      unsigned code[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int child = slot * cpt * cpt + (2 * wy + (t >> 1)) * cpt + 2 * wx + (t & 1);
        float v[8];
        unpack8p(sval[l - 1][child * MP_G + g], v);
        unsigned cd[8];
        if (BWD) {
          const uint2 pc = spos[l - 1][child * MP_G + g];
#pragma unroll
          for (int e = 0; e < 4; ++e) { cd[e] = (pc.x >> (8 * e)) & 255u; cd[4 + e] = (pc.y >> (8 * e)) & 255u; }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (t == 0) { m[e] = v[e]; if (BWD) code[e] = cd[e]; }
          else if (v[e] > m[e] || (BWD && v[e] == m[e] && cd[e] < code[e])) { m[e] = v[e]; if (BWD) code[e] = cd[e]; }
        }
      }
      const uint4 mv = pack8p(m);
      sval[l][w * MP_G + g] = mv;
      if (BWD) {
        spos[l][w * MP_G + g] = make_uint2(code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24),
                                           code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24));
      } else {
        const int n = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
        const int k = 2 << l, Ho = p.H / k, Wo = p.W / k;
        *reinterpret_cast<uint4*>(p.y[l] + (((size_t)n * Ho + ty * wpt + wy) * Wo + tx * wpt + wx) * p.C + c_base + g * 8) = mv;
      }
    }
    __syncthreads();
  }
  if (!BWD) return;
  // ---- backward: every pixel collects the gradient of each level whose window maximum it is ----
  for (int item = threadIdx.x; item < 64 * G; item += 256) {
    const int g = item % G, b = item / G;
    int slot, by, bx;
    mp_locate<NLEV>(b, slot, by, bx);
    const int tile = tile0 + slot;
    if (tile >= p.total_tiles) continue;
    const int n = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int c0 = c_base + g * 8;
    float gr[4][8];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) gr[t][e] = 0.f;
#pragma unroll
    for (int l = 0; l < NLEV; ++l) {
      if (p.dy[l] == nullptr) continue;
      const int wpt = T >> (l + 1), k = 2 << l;
      const int wy = by >> l, wx = bx >> l;                      // this 2x2 block's window at level l
      const int w = slot * wpt * wpt + wy * wpt + wx;
      const uint2 pc = spos[l][w * MP_G + g];
      const int Ho = p.H / k, Wo = p.W / k;
      float dv[8];
      unpack8p(__ldg(reinterpret_cast<const uint4*>(p.dy[l] + (((size_t)n * Ho + ty * wpt + wy) * Wo + tx * wpt + wx) * p.C + c0)), dv);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned cd = (2 * by + (t >> 1)) * T + 2 * bx + (t & 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned pe = ((e < 4 ? pc.x : pc.y) >> (8 * (e & 3))) & 255u;
          if (pe == cd) gr[t][e] += dv[e];
        }
      }
    }
    bf16* db = p.dx + (((size_t)n * p.H + ty * T + 2 * by) * p.W + tx * T + 2 * bx) * p.lddx + c0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint4* o = reinterpret_cast<uint4*>(db + ((size_t)(t >> 1) * p.W + (t & 1)) * p.lddx);
      if (p.acc) {
        float old[8];
        unpack8p(*o, old);
#pragma unroll
        for (int e = 0; e < 8; ++e) gr[t][e] += old[e];
      }
      *o = pack8p(gr[t]);
    }
  }
}

template <int NLEV, bool BWD>
int launch_multipool(const MultiPoolParams& p, cudaStream_t s) {
  constexpr int T = 1 << NLEV, TB = 256 / (T * T);
  size_t smem = 0;
  for (int l = 0; l < NLEV; ++l) smem += (size_t)(64 >> (2 * l)) * MP_G * 24;
  dim3 grid((unsigned)((p.total_tiles + TB - 1) / TB), (unsigned)((p.C + MP_CH - 1) / MP_CH));
  k_multipool<NLEV, BWD><<<grid, 256, smem, s>>>(p);
  return DS_LAUNCHED(BWD ? "multipool_bwd" : "multipool_fwd");
}

// Bias gradients of a MidMLKA (MixConvNeXtML.py:109-117) from per-plane fp32 statistics instead of a sum over the rounded
// bf16 gradient tensor.  o = conv1x1(cat(dw_k(x))) + b_conv, then o * CA(o), then InstanceNorm:
//   d b_conv[c]  = sum_n ( s[n,c] * S[n,c] + davg[n,c] + dmax[n,c] ),  S = sum_p of the norm's dx (fp32, ~0: dsgan_inorm_bwd_apply)
//   d b_dw[c']   = sum_c W_conv[c,c'] * d b_conv[c]        (a depthwise bias shifts o by W_conv . b_dw)
// C / 32 blocks of 256 threads; C <= 4096.
__global__ void k_mid_bias_grads(const float* __restrict__ S, const float* __restrict__ s, const float* __restrict__ davg,
                                 const float* __restrict__ dmax, int N, int C, const float* __restrict__ w,
                                 float* __restrict__ d_conv_bias, float* __restrict__ dq0, float* __restrict__ dq1,
                                 float* __restrict__ dq2, float* __restrict__ dq3) {
  extern __shared__ float db[];
  // every block rebuilds the (tiny) conv-bias gradient; block b then owns the depthwise-bias entries [32 b, 32 b + 32)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int n = 0; n < N; ++n) {
      const long long o = (long long)n * C + c;
      a += s[o] * S[o] + davg[o] + dmax[o];
    }
    db[c] = a;
    if (blockIdx.x == 0) atomicAdd(d_conv_bias + c, a);
  }
  __syncthreads();
  const int q = C / 4;
  const int cp = blockIdx.x * 32 + (threadIdx.x & 31), part = threadIdx.x >> 5, nparts = blockDim.x >> 5;
  float a = 0.f;
  if (cp < C)
    for (int c = part; c < C; c += nparts) a = fmaf(w[(long long)c * C + cp], db[c], a);   // coalesced along cp
  __shared__ float red[8][33];
  red[part][threadIdx.x & 31] = a;
  __syncthreads();
  if (part == 0 && cp < C) {
    for (int i = 1; i < nparts; ++i) a += red[i][threadIdx.x & 31];
    float* dst = cp < q ? dq0 : (cp < 2 * q ? dq1 : (cp < 3 * q ? dq2 : dq3));
    atomicAdd(dst + (cp % q), a);
  }
}
}  // namespace

extern "C" {
int dsgan_mid_bias_grads(const float* dsum_nc, const float* s, const float* davg, const float* dmax, int N, int C,
                         const float* w_conv, float* d_conv_bias, float* d_b3, float* d_b5, float* d_b7, float* d_b9,
                         void* stream) {
  DS_REQUIRE(C % 4 == 0 && C <= 4096, "mid_bias_grads: bad C=%d", C);
  k_mid_bias_grads<<<(C + 31) / 32, 256, sizeof(float) * C, (cudaStream_t)stream>>>(dsum_nc, s, davg, dmax, N, C, w_conv,
                                                                                  d_conv_bias, d_b3, d_b5, d_b7, d_b9);
  return DS_LAUNCHED("mid_bias_grads");
}
int dsgan_maxpool_fwd(const void* x, int ld_x, void* y, int ld_y, int dtype, int N, int H, int W, int C, int k,
                      void* stream) {
  DS_REQUIRE(k >= 1 && H % k == 0 && W % k == 0, "maxpool: H,W (%d,%d) must be multiples of k=%d", H, W, k);
  const long long total = (long long)N * (H / k) * (W / k) * C;
  if (dtype == DT_BF16 && (k == 2 || k == 4 || k == 8 || k == 16) && C % 8 == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 &&
      (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) {
#define DS_MP_FWD(K) k_maxpool_fwd_v8<K><<<cdiv(total / 8, 256), 256, 0, (cudaStream_t)stream>>>( \
    (const bf16*)x, ld_x, (bf16*)y, ld_y, H, W, C, total / 8)
    if (k == 2) DS_MP_FWD(2); else if (k == 4) DS_MP_FWD(4); else if (k == 8) DS_MP_FWD(8); else DS_MP_FWD(16);
#undef DS_MP_FWD
    return DS_LAUNCHED("maxpool_fwd_v8");
  }
  DS_DISPATCH_DT(dtype, (k_maxpool_fwd<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, ld_x, (T*)y,
                                                                                             ld_y, H, W, C, k, total)));
  return DS_LAUNCHED("maxpool_fwd");
}
int dsgan_maxpool_bwd(const void* x, int ld_x, const void* dy, int ld_dy, void* dx, int ld_dx, int dtype, int N, int H,
                      int W, int C, int k, int accumulate, int relu_mask, void* stream) {
  DS_REQUIRE(k >= 1 && H % k == 0 && W % k == 0, "maxpool: H,W (%d,%d) must be multiples of k=%d", H, W, k);
  const long long total = (long long)N * (H / k) * (W / k) * C;
  if (dtype == DT_BF16 && (k == 2 || k == 4 || k == 8 || k == 16) && C % 8 == 0 && ld_x % 8 == 0 && ld_dy % 8 == 0 &&
      ld_dx % 8 == 0 && (uintptr_t)x % 16 == 0 &&
      (uintptr_t)dy % 16 == 0 && (uintptr_t)dx % 16 == 0) {
#define DS_MP_BWD(K) k_maxpool_bwd_v8<K><<<cdiv(total / 8, 256), 256, 0, (cudaStream_t)stream>>>( \
    (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, (bf16*)dx, ld_dx, H, W, C, accumulate, relu_mask, total / 8)
    if (k == 2) DS_MP_BWD(2); else if (k == 4) DS_MP_BWD(4); else if (k == 8) DS_MP_BWD(8); else DS_MP_BWD(16);
#undef DS_MP_BWD
    return DS_LAUNCHED("maxpool_bwd_v8");
  }
  DS_DISPATCH_DT(dtype, (k_maxpool_bwd<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, ld_x, (const T*)dy, ld_dy, (T*)dx, ld_dx, H, W, C, k, accumulate, relu_mask,
                            total)));
  return DS_LAUNCHED("maxpool_bwd");
}
int dsgan_multipool_supported(int dtype, int H, int W, int C, int nlev, int ld_x) {
  if (dtype != DT_BF16 || nlev < 2 || nlev > 4 || C % 8 || ld_x % 8) return 0;
  const int T = 1 << nlev;
  return (H % T == 0 && W % T == 0) ? 1 : 0;
}
static int multipool_fill(MultiPoolParams* p, const void* x, int ld_x, int N, int H, int W, int C, int nlev) {
  memset(p, 0, sizeof(*p));
  const int T = 1 << nlev;
  p->x = (const bf16*)x; p->ldx = ld_x; p->N = N; p->H = H; p->W = W; p->C = C;
  p->tiles_x = W / T; p->tiles_y = H / T; p->total_tiles = N * p->tiles_x * p->tiles_y;
  return 0;
}
int dsgan_multipool_fwd(const void* x, int ld_x, void* y2, void* y4, void* y8, void* y16, int dtype, int N, int H, int W, int C,
                        int nlev, void* stream) {
  DS_REQUIRE(dsgan_multipool_supported(dtype, H, W, C, nlev, ld_x), "multipool_fwd: unsupported shape %dx%dx%d nlev=%d", H, W, C, nlev);
  DS_REQUIRE(((uintptr_t)x % 16 == 0) && y2 && y4 && (nlev < 3 || y8) && (nlev < 4 || y16), "multipool_fwd: bad pointers");
  MultiPoolParams p;
  multipool_fill(&p, x, ld_x, N, H, W, C, nlev);
  p.y[0] = (bf16*)y2; p.y[1] = (bf16*)y4; p.y[2] = (bf16*)y8; p.y[3] = (bf16*)y16;
  cudaStream_t s = (cudaStream_t)stream;
  if (nlev == 2) return launch_multipool<2, false>(p, s);
  if (nlev == 3) return launch_multipool<3, false>(p, s);
  return launch_multipool<4, false>(p, s);
}
int dsgan_multipool_bwd(const void* x, int ld_x, const void* dy2, const void* dy4, const void* dy8, const void* dy16, void* dx,
                        int ld_dx, int accumulate, int dtype, int N, int H, int W, int C, int nlev, void* stream) {
  DS_REQUIRE(dsgan_multipool_supported(dtype, H, W, C, nlev, ld_x) && ld_dx % 8 == 0, "multipool_bwd: unsupported shape");
  DS_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)dx % 16 == 0), "multipool_bwd: unaligned");
  MultiPoolParams p;
  multipool_fill(&p, x, ld_x, N, H, W, C, nlev);
  p.dy[0] = (const bf16*)dy2; p.dy[1] = (const bf16*)dy4; p.dy[2] = (const bf16*)dy8; p.dy[3] = (const bf16*)dy16;
  p.dx = (bf16*)dx; p.lddx = ld_dx; p.acc = accumulate;
  cudaStream_t s = (cudaStream_t)stream;
  if (nlev == 2) return launch_multipool<2, true>(p, s);
  if (nlev == 3) return launch_multipool<3, true>(p, s);
  return launch_multipool<4, true>(p, s);
}
int dsgan_ca_fwd(const void* x, int dtype, int N, long long HW, int C, const float* fc1, const float* slope,
                 const float* fc2, float* avg, float* mx, int* argmax, float* s, void* workspace, void* stream) {
  DS_REQUIRE(C % 8 == 0 && C <= 2048, "ca_fwd: C=%d unsupported", C);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* packed = (unsigned long long*)workspace;
  cudaMemsetAsync(avg, 0, sizeof(float) * N * C, st);
  cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * N * C, st);
  const int chunk = 512;
  dim3 grid(cdiv(HW, chunk), N);
  if (dtype == DT_BF16 && (uintptr_t)x % 16 == 0) {
    const int groups = C / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;
    int ch = 2048;                                     // pixels per block: shrink until the launch covers the machine
    const int cblocks = (C + gl * 8 - 1) / (gl * 8);
    while (ch > 4 * pl && ((HW + ch - 1) / ch) * N * cblocks < 296) ch >>= 1;
    dim3 g8(cdiv(HW, ch), N, cblocks);
    k_ca_pool_v8<<<g8, 256, 0, st>>>((const bf16*)x, HW, C, avg, packed, ch);
  } else {
    DS_DISPATCH_DT(dtype, (k_ca_pool<T><<<grid, 256, 0, st>>>((const T*)x, HW, C, avg, packed, chunk)));
  }
  if (DS_LAUNCHED("ca_pool")) return 1;
  const size_t smem = sizeof(float) * (2 * C + 2 * (C / 8));
  k_ca_mlp<<<N, 256, smem, st>>>(packed, HW, C, fc1, slope, fc2, avg, mx, argmax, s);
  return DS_LAUNCHED("ca_mlp");
}
int dsgan_ca_bwd(const float* ds, const float* s, const float* avg, const float* mx, int N, int C, const float* fc1,
                 const float* slope, const float* fc2, float* dfc1, float* dslope, float* dfc2, float* davg,
                 float* dmax, void* stream) {
  const size_t smem = sizeof(float) * (3 * C + 4 * (C / 8) + 32);
  k_ca_bwd<<<N, 256, smem, (cudaStream_t)stream>>>(ds, s, avg, mx, C, fc1, slope, fc2, dfc1, dslope, dfc2, davg, dmax);
  return DS_LAUNCHED("ca_bwd");
}
}
