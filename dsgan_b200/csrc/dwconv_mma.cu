// Depthwise k x k convolution (ConvNeXt Block 7x7, MidMLKA 3/5/7/9 branches) on the warp-level tensor-core path.
//
// Reference: networks.py Block.dwconv (nn.Conv2d(dim, dim, 7, padding=3, groups=dim)) and the MidMLKA LKA/X branches
// (groups=dim, k in {3,5,7,9}); forward, input gradient (flipped taps) and weight gradient.
//
// Why tensor cores for a depthwise filter: 49 multiply-adds per 4 bytes of traffic make the 7x7 filter FMA-issue bound on the
// CUDA cores (dwconv.cu k_dwconv_t: 0.45 ms for 16x256x256x128, 18 % of the HBM roofline).  Per channel the filter is a banded
// (Toeplitz) matrix product along x:
//     out_c[y, x0+n] = sum_ky  sum_k  in_c[y+ky, x0+k] * T_ky[k, n],   T_ky[k, n] = w_c[ky, k-n]  (0 <= k-n < K, else 0)
// i.e. for 8 output columns a [16 rows x 16 cols] x [16 x 8] product per ky: one mma.m16n8k16 (2.3x padded work for K=7, on
// a pipe 8x wider than FFMA).  The Toeplitz fragments live in registers for the life of the CTA; the data operand comes from a
// channel-planar bf16 tile in shared memory through ldmatrix, whose per-lane row addresses make the ky row shift free (the
// reason this is mma.sync + ldmatrix and not tcgen05: a UMMA shared-memory descriptor cannot start at an arbitrary row of a
// swizzled tile, and the shift would have to be paid as 7 shifted copies of the tile or 7 cross-lane TMEM reductions).
// Weights enter as bf16 like those of every other tensor-core layer of the bf16 mode (fp32 accumulate); DSGAN_DW_SPLIT=1 applies
// them as hi + lo (two MMAs on the same data fragment, 16 mantissa bits).  The step-level parity numbers are the same either
// way (fake_B 2.07e-2, G gradients 1.2e-2 at 16x256x256) and the split costs 14 % of the forward kernel, so it is off.
//
// The weight gradient uses the transposed product: for 16 rows of dy and 8 columns, P_ky[m, n] = sum_y in[y+ky, x0+m] *
// dy[y, x0+n] accumulated over ALL tiles into one 16x8 fragment per ky; dw[ky, kx] is the sum of its kx-th diagonal.
//
// Layout: NHWC bf16 in HBM; a CTA owns 16 channels (32 B of every pixel = one DRAM sector) of a tile, warp w owns channels
// 2w, 2w+1.  Tiles are staged global -> registers -> planar shared ([16 ch][rows][cols], row pitch = cols*2 B chosen so that
// 8 consecutive rows hit 8 different 16-byte bank groups), results are written back in place into the dead rows of the planes
// and leave as 16-byte NHWC vectors.  Algorithmic traffic: read x once, write y once (wgrad: read x and dy once).
#include "dwconv_mma.cuh"
#include <stdlib.h>

namespace dsgan {
namespace dwm {
namespace {

constexpr int CB = 16;  // channels per CTA
#ifndef DW_STAGE_U
#define DW_STAGE_U 3   // pixel pairs per thread in flight while staging (6 independent 32-byte loads; 4 measured the same)
#endif

__device__ __forceinline__ void ldsm_x4(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t a, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t a, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
// D += A(16x16, row) * B(16x8, col), bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ void ld_px32(const bf16* p, bool wide, bool two, uint32_t (&w)[8]) {
  if (wide) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]),
                 "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
  } else {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = two ? __ldg(reinterpret_cast<const uint4*>(p) + 1) : make_uint4(0u, 0u, 0u, 0u);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  }
}
__device__ __forceinline__ void st_px32(bf16* p, bool wide, bool two, const uint32_t (&w)[8]) {
  if (wide) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  } else {
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    if (two) *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// rows [yorg, yorg+ROWS) x cols [xorg, xorg+COLS) x channels [c_base, c_base+16) of image n -> planes [16][ROWS][PITCHB bytes]
// (zero outside the image / beyond C).  A thread takes two x-adjacent pixels (one 32-byte access each: the CTA's 16 channels
// are exactly one sector of the pixel) and writes one 32-bit word per channel plane; lanes walk x, so every plane store of a
// warp is one conflict-free wavefront.  wide: 32-byte aligned pixels and all 16 channels inside the tensor.
template <int ROWS, int COLS, int PITCHB>
__device__ __forceinline__ void stage_planar(unsigned char* sm, const bf16* __restrict__ src, int ld, int n, int H, int W,
                                             int yorg, int xorg, int c_base, int C, int tid, bool wide) {
  static_assert(COLS % 2 == 0, "pixel pairs");
  constexpr int PAIRS = COLS / 2, ITEMS = ROWS * PAIRS, PLANE = ROWS * PITCHB, U = DW_STAGE_U;
  const bool two = c_base + 8 < C;
  for (int i0 = tid; i0 < ITEMS; i0 += 256 * U) {
    uint32_t v[U][2][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      const int pp = i % PAIRS, py = i / PAIRS, gy = yorg + py;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int gx = xorg + 2 * pp + q;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[u][q][e] = 0u;
        if (i < ITEMS && gy >= 0 && gy < H && gx >= 0 && gx < W)
          ld_px32(src + (((size_t)n * H + gy) * W + gx) * ld + c_base, wide, two, v[u][q]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      if (i < ITEMS) {
        const int pp = i % PAIRS, py = i / PAIRS;
        unsigned char* b = sm + py * PITCHB + pp * 4;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          *reinterpret_cast<uint32_t*>(b + (2 * e) * PLANE) = prmt(v[u][0][e], v[u][1][e], 0x5410u);
          *reinterpret_cast<uint32_t*>(b + (2 * e + 1) * PLANE) = prmt(v[u][0][e], v[u][1][e], 0x7632u);
        }
      }
    }
  }
}

template <int K, int NT, int NS>
struct Geo {
  static constexpr int TX = 8 * NT, TY = 16 * NS, P = K / 2, IR = TY + K - 1, ICP = TX + 8, PITCHB = ICP * 2,
                       PLANE = IR * PITCHB, SMEM_X = CB * PLANE;
  // wgrad: dy planes [16][TY][TX (+8 pad)]
  static constexpr int GPLANE = TY * PITCHB, SMEM_G = CB * GPLANE;
};

// forward / input gradient.  grid: (tile groups, channel blocks); a CTA walks tiles blockIdx.x, +gridDim.x, ...
template <int K, int NT, int NS, bool SPLIT>
__device__ __forceinline__ void dw_mma_body(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                            const float* __restrict__ bias, bf16* __restrict__ y, int ldy, int N, int H,
                                            int W, int C, int flip, int acc_out, int tiles_x, int tiles_y,
                                            int total_tiles, int c_base, unsigned char* dsm) {
  using G = Geo<K, NT, NS>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  // Toeplitz B fragments of this warp's two channels: b0 = T[2t, 2t+1][g], b1 = T[2t+8, 2t+9][g], T[k][n] = w[ky][k-n]
  uint32_t bh[2][K][2], bl[2][K][2];
  float bias_c[2];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = c_base + warp * 2 + cc;
    bias_c[cc] = (bias && c < C) ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      float wv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 2 * t + (e & 1) + (e >> 1) * 8, d = k - g;
        const int tap = flip ? (K - 1 - ky) * K + (K - 1 - d) : ky * K + d;
        wv[e] = (d >= 0 && d < K && c < C) ? __ldg(w + (size_t)c * K * K + tap) : 0.f;
      }
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __bfloat162float(__float2bfloat16_rn(wv[e]));
        lo[e] = wv[e] - hi[e];
      }
      bh[cc][ky][0] = pack_bf2(hi[0], hi[1]);
      bh[cc][ky][1] = pack_bf2(hi[2], hi[3]);
      bl[cc][ky][0] = pack_bf2(lo[0], lo[1]);
      bl[cc][ky][1] = pack_bf2(lo[2], lo[3]);
    }
  }
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(dsm);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lchunk = lane >> 4;
  const bool full16 = c_base + CB <= C;
  const bool wide_x = full16 && ldx % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 32 == 0);
  const bool wide_y = full16 && ldy % 16 == 0 && (reinterpret_cast<uintptr_t>(y) % 32 == 0);

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
    const int y0 = ty * G::TY, x0 = tx * G::TX;
    __syncthreads();  // the previous tile's results have left the planes
    stage_planar<G::IR, G::ICP, G::PITCHB>(dsm, x, ldx, n, H, W, y0 - G::P, x0 - G::P, c_base, C, tid, wide_x);
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int pl = warp * 2 + cc;
      const uint32_t plane = sm_base + pl * G::PLANE;
      unsigned char* plane_g = dsm + pl * G::PLANE;
#pragma unroll 1
      for (int s = 0; s < NS; ++s) {
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = bias_c[cc];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const uint32_t a = plane + (s * 16 + ky + lrow) * G::PITCHB + lchunk * 16;
          uint32_t ch[NT + 1][2];
#pragma unroll
          for (int j2 = 0; j2 < NT / 2; ++j2)
            ldsm_x4(a + j2 * 32, ch[2 * j2][0], ch[2 * j2][1], ch[2 * j2 + 1][0], ch[2 * j2 + 1][1]);
          ldsm_x2(a + NT * 16, ch[NT][0], ch[NT][1]);  // x2 takes the row addresses of lanes 0..15 (lchunk = 0 there)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(acc[j], ch[j][0], ch[j][1], ch[j + 1][0], ch[j + 1][1], bh[cc][ky][0], bh[cc][ky][1]);
          if (SPLIT) {
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(acc[j], ch[j][0], ch[j][1], ch[j + 1][0], ch[j + 1][1], bl[cc][ky][0], bl[cc][ky][1]);
          }
        }
        // rows s*16 .. s*16+15 of this plane are dead now (later strips read rows >= (s+1)*16): results go there, in place
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          *reinterpret_cast<uint32_t*>(plane_g + (s * 16 + g) * G::PITCHB + (8 * j + 2 * t) * 2) = pack_bf2(acc[j][0], acc[j][1]);
          *reinterpret_cast<uint32_t*>(plane_g + (s * 16 + g + 8) * G::PITCHB + (8 * j + 2 * t) * 2) = pack_bf2(acc[j][2], acc[j][3]);
        }
      }
    }
    __syncthreads();
    // planes -> NHWC: a thread gathers two x-adjacent pixels (one 32-bit word per plane) and stores 32 bytes per pixel
    {
      constexpr int PAIRS = G::TX / 2, ITEMS = G::TY * PAIRS;
      const bool two = c_base + 8 < C;
#pragma unroll 2
      for (int i = tid; i < ITEMS; i += 256) {
        const int pp = i % PAIRS, py = i / PAIRS, gy = y0 + py, gx = x0 + 2 * pp;
        if (gy >= H || gx >= W) continue;
        const unsigned char* b = dsm + py * G::PITCHB + pp * 4;
        uint32_t r[16], o[2][8];
#pragma unroll
        for (int c = 0; c < 16; ++c) r[c] = *reinterpret_cast<const uint32_t*>(b + c * G::PLANE);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[0][e] = prmt(r[2 * e], r[2 * e + 1], 0x5410u);
          o[1][e] = prmt(r[2 * e], r[2 * e + 1], 0x7632u);
        }
        bf16* yp = y + (((size_t)n * H + gy) * W + gx) * ldy + c_base;
        const bool second = gx + 1 < W;
        if (acc_out) {
          uint32_t old[2][8];
          ld_px32(yp, wide_y, two, old[0]);
          if (second) ld_px32(yp + ldy, wide_y, two, old[1]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q == 1 && !second) break;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 a = unpack_bf2(o[q][e]), c2 = unpack_bf2(old[q][e]);
              o[q][e] = pack_bf2(a.x + c2.x, a.y + c2.y);
            }
          }
        }
        st_px32(yp, wide_y, two, o[0]);
        if (second) st_px32(yp + ldy, wide_y, two, o[1]);
      }
    }
  }
}

// weight (and bias) gradient.  Same tiling; the K fragments P_ky of a warp's two channels stay in registers across all tiles
// of the CTA and are reduced once: diagonal kx of P_ky -> shared -> one atomicAdd per (channel, tap) per CTA.
template <int K, int NT, int NS>
__device__ __forceinline__ void dw_mma_wgrad_body(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                  int lddy, float* __restrict__ dw, float* __restrict__ db, int N,
                                                  int H, int W, int C, int tiles_x, int tiles_y, int total_tiles,
                                                  int c_base, unsigned char* dsm) {
  using G = Geo<K, NT, NS>;
  unsigned char* sg = dsm + G::SMEM_X;
  float* sacc = reinterpret_cast<float*>(dsm + G::SMEM_X + G::SMEM_G);  // [16][K*K+1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < CB * (K * K + 1); i += 256) sacc[i] = 0.f;
  float acc[2][K][4], accb[2][4];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
    for (int ky = 0; ky < K; ++ky) acc[cc][ky][0] = acc[cc][ky][1] = acc[cc][ky][2] = acc[cc][ky][3] = 0.f;
    accb[cc][0] = accb[cc][1] = accb[cc][2] = accb[cc][3] = 0.f;
  }
  const uint32_t sx_base = (uint32_t)__cvta_generic_to_shared(dsm), sg_base = (uint32_t)__cvta_generic_to_shared(sg);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lchunk = lane >> 4;
  const uint32_t ones = 0x3F803F80u;  // bf16 (1, 1)
  const bool full16 = c_base + CB <= C;
  const bool wide_x = full16 && ldx % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 32 == 0);
  const bool wide_g = full16 && lddy % 16 == 0 && (reinterpret_cast<uintptr_t>(dy) % 32 == 0);

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
    const int y0 = ty * G::TY, x0 = tx * G::TX;
    __syncthreads();
    stage_planar<G::IR, G::ICP, G::PITCHB>(dsm, x, ldx, n, H, W, y0 - G::P, x0 - G::P, c_base, C, tid, wide_x);
    stage_planar<G::TY, G::TX, G::PITCHB>(sg, dy, lddy, n, H, W, y0, x0, c_base, C, tid, wide_g);
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int pl = warp * 2 + cc;
      const uint32_t xplane = sx_base + pl * G::PLANE, gplane = sg_base + pl * G::GPLANE;
#pragma unroll 1
      for (int s = 0; s < NS; ++s) {
        // B fragments: dy rows s*16 .. +15 (k), 8 columns per tile (n); stored [row][col] -> .trans
        uint32_t gb[NT][2];
        {
          const uint32_t a = gplane + (s * 16 + lrow) * G::PITCHB + lchunk * 16;
#pragma unroll
          for (int j2 = 0; j2 < NT / 2; ++j2) ldsm_x4_t(a + j2 * 32, gb[2 * j2][0], gb[2 * j2][1], gb[2 * j2 + 1][0], gb[2 * j2 + 1][1]);
        }
        if (db) {
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(accb[cc], ones, ones, ones, ones, gb[j][0], gb[j][1]);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          // A[m = window column][k = row]: memory is [row][col] -> .trans; chunk j = (rows 0-7, rows 8-15) of columns 8j..8j+7
          const uint32_t a = xplane + (s * 16 + ky + lrow) * G::PITCHB + lchunk * 16;
          uint32_t ch[NT + 1][2];
#pragma unroll
          for (int j2 = 0; j2 < NT / 2; ++j2)
            ldsm_x4_t(a + j2 * 32, ch[2 * j2][0], ch[2 * j2][1], ch[2 * j2 + 1][0], ch[2 * j2 + 1][1]);
          ldsm_x2_t(a + NT * 16, ch[NT][0], ch[NT][1]);
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(acc[cc][ky], ch[j][0], ch[j + 1][0], ch[j][1], ch[j + 1][1], gb[j][0], gb[j][1]);
        }
      }
    }
  }
  // diagonal sums: fragment element e of thread (g, t) is P[m = g + 8*(e>>1)][n = 2t + (e&1)], tap kx = m - n
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    float* sa = sacc + (warp * 2 + cc) * (K * K + 1);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kx = g + 8 * (e >> 1) - 2 * t - (e & 1);
        if (kx >= 0 && kx < K) atomicAdd(sa + ky * K + kx, acc[cc][ky][e]);
      }
    }
    if (db && g == 0) atomicAdd(sa + K * K, accb[cc][0] + accb[cc][1]);  // every row of ones*dy holds the column sums
  }
  __syncthreads();
  for (int i = tid; i < CB * (K * K + 1); i += 256) {
    const int c = c_base + i / (K * K + 1), tap = i % (K * K + 1);
    if (c >= C) continue;
    if (tap < K * K) atomicAdd(dw + (size_t)c * K * K + tap, sacc[i]);
    else if (db) atomicAdd(db + c, sacc[i]);
  }
}


template <int K, int NT, int NS, bool SPLIT>
__global__ void __launch_bounds__(256, 2) k_dw_mma(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                    const float* __restrict__ bias, bf16* __restrict__ y, int ldy, int N, int H,
                                                    int W, int C, int flip, int acc_out, int tiles_x, int tiles_y,
                                                    int total_tiles) {
  extern __shared__ __align__(128) unsigned char dsm[];
  dw_mma_body<K, NT, NS, SPLIT>(x, ldx, w, bias, y, ldy, N, H, W, C, flip, acc_out, tiles_x, tiles_y, total_tiles,
                                blockIdx.y * CB, dsm);
}
template <int K, int NT, int NS>
__global__ void __launch_bounds__(256, 2) k_dw_mma_wgrad(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                          int lddy, float* __restrict__ dw, float* __restrict__ db, int N,
                                                          int H, int W, int C, int tiles_x, int tiles_y, int total_tiles) {
  extern __shared__ __align__(128) unsigned char dsm[];
  dw_mma_wgrad_body<K, NT, NS>(x, ldx, dy, lddy, dw, db, N, H, W, C, tiles_x, tiles_y, total_tiles, blockIdx.y * CB, dsm);
}

// Several depthwise convolutions of different kernel sizes over disjoint channel slices of ONE tensor (MidMLKA.X3/X5/X7/X9,
// MixConvNeXtML.py:94-97,110) in one launch: blockIdx.y walks the channel blocks of all branches.  The per-branch launches
// were one tile per CTA and latency bound (~30 us each for 17 MB of traffic); four at once fill the machine.
struct DwBranch {
  const float* w; const float* bias; float* dw; float* db;
  int k, c0, c, blk0;   // kernel size, channel slice [c0, c0+c), first channel block of the branch
};
struct DwMulti { DwBranch b[4]; int n; };

template <int NT, int NS>
__global__ void __launch_bounds__(256, 2) k_dw_mma_multi(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy,
                                                          int N, int H, int W, const DwMulti m, int flip, int acc_out,
                                                          int tiles_x, int tiles_y, int total_tiles) {
  extern __shared__ __align__(128) unsigned char dsm[];
  int bi = 0;
  while (bi + 1 < m.n && (int)blockIdx.y >= m.b[bi + 1].blk0) ++bi;
  const DwBranch b = m.b[bi];
  const int c_base = ((int)blockIdx.y - b.blk0) * CB;
#define DWM_BODY(KK) dw_mma_body<KK, NT, NS, false>(x + b.c0, ldx, b.w, b.bias, y + b.c0, ldy, N, H, W, b.c, flip, acc_out, \
                                                   tiles_x, tiles_y, total_tiles, c_base, dsm)
  switch (b.k) {
    case 3: DWM_BODY(3); break;
    case 5: DWM_BODY(5); break;
    case 7: DWM_BODY(7); break;
    default: DWM_BODY(9); break;
  }
#undef DWM_BODY
}
template <int NT, int NS>
__global__ void __launch_bounds__(256, 2) k_dw_mma_wgrad_multi(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy,
                                                                int lddy, int N, int H, int W, const DwMulti m, int tiles_x,
                                                                int tiles_y, int total_tiles) {
  extern __shared__ __align__(128) unsigned char dsm[];
  int bi = 0;
  while (bi + 1 < m.n && (int)blockIdx.y >= m.b[bi + 1].blk0) ++bi;
  const DwBranch b = m.b[bi];
  const int c_base = ((int)blockIdx.y - b.blk0) * CB;
#define DWM_BODY(KK) dw_mma_wgrad_body<KK, NT, NS>(x + b.c0, ldx, dy + b.c0, lddy, b.dw, b.db, N, H, W, b.c, tiles_x, tiles_y, \
                                                   total_tiles, c_base, dsm)
  switch (b.k) {
    case 3: DWM_BODY(3); break;
    case 5: DWM_BODY(5); break;
    case 7: DWM_BODY(7); break;
    default: DWM_BODY(9); break;
  }
#undef DWM_BODY
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}
// tile groups per channel block: fill the 2 x SMs CTA slots, every CTA the same number of tiles where possible
inline int tile_groups(int total_tiles, int cblocks) {
  int gmax = (2 * sm_count()) / cblocks;
  if (gmax < 1) gmax = 1;
  if (gmax >= total_tiles) return total_tiles;
  const int per = (total_tiles + gmax - 1) / gmax;
  return (total_tiles + per - 1) / per;
}

template <int K, int NT, int NS>
int launch_fwd(const bf16* x, int ldx, const float* w, const float* bias, bf16* y, int ldy, int N, int H, int W, int C,
               int flip, int acc, cudaStream_t s) {
  using G = Geo<K, NT, NS>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_dw_mma<K, NT, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_X);
    cudaFuncSetAttribute(k_dw_mma<K, NT, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_X);
    attr = true;
  }
  const int tiles_x = cdiv(W, G::TX), tiles_y = cdiv(H, G::TY), total = N * tiles_x * tiles_y, cblocks = cdiv(C, CB);
  dim3 grid((unsigned)tile_groups(total, cblocks), (unsigned)cblocks);
  static int split = -1;   // DSGAN_DW_SPLIT=1: weights as bf16 hi + lo (two MMAs per fragment, 16 mantissa bits); default: one bf16 MMA
  if (split < 0) { const char* e = getenv("DSGAN_DW_SPLIT"); split = (e && e[0] == '1') ? 1 : 0; }
  if (split) k_dw_mma<K, NT, NS, true><<<grid, 256, G::SMEM_X, s>>>(x, ldx, w, bias, y, ldy, N, H, W, C, flip, acc, tiles_x, tiles_y, total);
  else k_dw_mma<K, NT, NS, false><<<grid, 256, G::SMEM_X, s>>>(x, ldx, w, bias, y, ldy, N, H, W, C, flip, acc, tiles_x, tiles_y, total);
  return DS_LAUNCHED("dwconv_mma");
}
template <int K, int NT, int NS>
int launch_wgrad(const bf16* x, int ldx, const bf16* dy, int lddy, float* dw, float* db, int N, int H, int W, int C,
                 cudaStream_t s) {
  using G = Geo<K, NT, NS>;
  constexpr int smem = G::SMEM_X + G::SMEM_G + CB * (K * K + 1) * 4;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_dw_mma_wgrad<K, NT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  const int tiles_x = cdiv(W, G::TX), tiles_y = cdiv(H, G::TY), total = N * tiles_x * tiles_y, cblocks = cdiv(C, CB);
  dim3 grid((unsigned)tile_groups(total, cblocks), (unsigned)cblocks);
  k_dw_mma_wgrad<K, NT, NS><<<grid, 256, smem, s>>>(x, ldx, dy, lddy, dw, db, N, H, W, C, tiles_x, tiles_y, total);
  return DS_LAUNCHED("dwconv_mma_wgrad");
}


template <int NT, int NS>
int launch_multi_fwd(const bf16* x, int ldx, bf16* y, int ldy, int N, int H, int W, const DwMulti& m, int cblocks, int flip,
                     int acc, cudaStream_t s) {
  using G = Geo<9, NT, NS>;   // shared memory for the largest kernel size
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_dw_mma_multi<NT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_X); attr = true; }
  const int tiles_x = cdiv(W, G::TX), tiles_y = cdiv(H, G::TY), total = N * tiles_x * tiles_y;
  // Branches differ 3x in cost (k = 3 .. 9): tile groups are sized as for ONE branch, so that the grid is several waves of short
  // CTAs and the block scheduler balances the branches (a one-wave persistent grid finished with the k = 9 CTAs alone).
  dim3 grid((unsigned)tile_groups(total, (cblocks + m.n - 1) / m.n), (unsigned)cblocks);
  k_dw_mma_multi<NT, NS><<<grid, 256, G::SMEM_X, s>>>(x, ldx, y, ldy, N, H, W, m, flip, acc, tiles_x, tiles_y, total);
  return DS_LAUNCHED("dwconv_mma_multi");
}
template <int NT, int NS>
int launch_multi_wgrad(const bf16* x, int ldx, const bf16* dy, int lddy, int N, int H, int W, const DwMulti& m, int cblocks,
                       cudaStream_t s) {
  using G = Geo<9, NT, NS>;
  constexpr int smem = G::SMEM_X + G::SMEM_G + CB * (9 * 9 + 1) * 4;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_dw_mma_wgrad_multi<NT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  const int tiles_x = cdiv(W, G::TX), tiles_y = cdiv(H, G::TY), total = N * tiles_x * tiles_y;
  dim3 grid((unsigned)tile_groups(total, (cblocks + m.n - 1) / m.n), (unsigned)cblocks);
  k_dw_mma_wgrad_multi<NT, NS><<<grid, 256, smem, s>>>(x, ldx, dy, lddy, N, H, W, m, tiles_x, tiles_y, total);
  return DS_LAUNCHED("dwconv_mma_wgrad_multi");
}

// -> branch table; false if a branch is not eligible (kernel size, slice width / alignment)
inline bool make_multi(const dsgan_dw_branch* br, int nbr, DwMulti* m, int* cblocks) {
  if (nbr < 1 || nbr > 4) return false;
  int blk = 0;
  for (int i = 0; i < nbr; ++i) {
    const dsgan_dw_branch& b = br[i];
    if (!(b.k == 3 || b.k == 5 || b.k == 7 || b.k == 9) || b.c < 16 || b.c % 16 || b.c0 % 16) return false;
    m->b[i].w = b.w; m->b[i].bias = b.bias; m->b[i].dw = b.dw; m->b[i].db = b.db;
    m->b[i].k = b.k; m->b[i].c0 = b.c0; m->b[i].c = b.c; m->b[i].blk0 = blk;
    blk += b.c / CB;
  }
  m->n = nbr;
  *cblocks = blk;
  return true;
}

inline bool shape_ok(const void* a, int lda, const void* b, int ldb, int H, int W, int C) {
  return C >= 16 && C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) &&
         H >= 16 && W >= 32;
}
}  // namespace

bool fwd_try(const bf16* x, int ldx, const float* w, const float* bias, bf16* y, int ldy, int N, int H, int W, int C, int k,
             int flip, int accumulate, cudaStream_t s, int* rc) {
  if (!shape_ok(x, ldx, y, ldy, H, W, C)) return false;
#define DWM_F(KK, NT, NS) *rc = launch_fwd<KK, NT, NS>(x, ldx, w, bias, y, ldy, N, H, W, C, flip, accumulate, s); return true;
#define DWM_FK(KK)                                    \
  case KK:                                            \
    if (W >= 64 && H >= 32) { DWM_F(KK, 8, 2) }       \
    else if (H >= 32) { DWM_F(KK, 4, 2) }             \
    else { DWM_F(KK, 4, 1) }
  switch (k) {
    DWM_FK(3) DWM_FK(5) DWM_FK(7) DWM_FK(9)
    default: return false;
  }
#undef DWM_FK
#undef DWM_F
}

bool wgrad_try(const bf16* x, int ldx, const bf16* dy, int lddy, float* dw, float* db, int N, int H, int W, int C, int k,
               cudaStream_t s, int* rc) {
  if (!shape_ok(x, ldx, dy, lddy, H, W, C)) return false;
#define DWM_W(KK, NT, NS) *rc = launch_wgrad<KK, NT, NS>(x, ldx, dy, lddy, dw, db, N, H, W, C, s); return true;
#define DWM_WK(KK)                                    \
  case KK:                                            \
    if (H >= 32) { DWM_W(KK, 4, 2) }                  \
    else { DWM_W(KK, 4, 1) }
  switch (k) {
    DWM_WK(3) DWM_WK(5) DWM_WK(7) DWM_WK(9)
    default: return false;
  }
#undef DWM_WK
#undef DWM_W
}

bool multi_fwd_try(const bf16* x, int ldx, bf16* y, int ldy, int N, int H, int W, const dsgan_dw_branch* br, int nbr, int flip,
                   int accumulate, cudaStream_t s, int* rc) {
  DwMulti m;
  int cblocks = 0;
  if (!make_multi(br, nbr, &m, &cblocks) || !shape_ok(x, ldx, y, ldy, H, W, 16)) return false;
  if (W >= 64 && H >= 32) *rc = launch_multi_fwd<8, 2>(x, ldx, y, ldy, N, H, W, m, cblocks, flip, accumulate, s);
  else if (H >= 32) *rc = launch_multi_fwd<4, 2>(x, ldx, y, ldy, N, H, W, m, cblocks, flip, accumulate, s);
  else *rc = launch_multi_fwd<4, 1>(x, ldx, y, ldy, N, H, W, m, cblocks, flip, accumulate, s);
  return true;
}
bool multi_wgrad_try(const bf16* x, int ldx, const bf16* dy, int lddy, int N, int H, int W, const dsgan_dw_branch* br, int nbr,
                     cudaStream_t s, int* rc) {
  DwMulti m;
  int cblocks = 0;
  if (!make_multi(br, nbr, &m, &cblocks) || !shape_ok(x, ldx, dy, lddy, H, W, 16)) return false;
  if (H >= 32) *rc = launch_multi_wgrad<4, 2>(x, ldx, dy, lddy, N, H, W, m, cblocks, s);
  else *rc = launch_multi_wgrad<4, 1>(x, ldx, dy, lddy, N, H, W, m, cblocks, s);
  return true;
}

}  // namespace dwm
}  // namespace dsgan
