// tcgen05 / TMEM / TMA implicit-GEMM convolution over NHWC bf16 activations.
//
// One output tile = an 8x16 patch of "grid positions" of one image (128 GEMM rows) x BN output channels.  For every
// filter tap the A operand is ONE 4-D TMA box {64 channels, 16, 8, 1} of the input, fetched at the tap's spatial offset —
// out-of-image coordinates are zero-filled by TMA, which implements the padding — and the B operand is a K-major slab
// [tap][Cout][Cin] of the packed bf16 weights.  Grid position (y,x) reads input pixel (y*is + dy_t, x*is + dx_t) and writes
// output pixel (y*os + oy0, x*os + ox0), which covers
//   * nn.Conv2d stride 1 (VGG 3x3, PatchGAN 4x4 s1) and its input-gradient (flipped taps, transposed slabs)    is=1 os=1
//   * nn.Conv2d stride 2 (PatchGAN 4x4 s2) and the input-gradient of nn.ConvTranspose2d(s2)                    is=2 os=1
//   * nn.ConvTranspose2d(k3,s2,p1,op1) forward and the input-gradient of a stride-2 conv, one launch per output
//     parity class (1/2/2/4 taps)                                                                              is=1 os=2
// Reference op sites: models/vgg.py:16-25, networks.py:544-569, MixConvNeXtML.py:53,150.
// Same warp-specialised persistent pipeline as tc_gemm.cu (TMA warp / single-thread MMA issuer / 4 epilogue warps,
// 4-stage smem ring, double-buffered TMEM accumulator).
#include "tc_common.cuh"
#include "../../include/dsgan_b200.h"
#include <stdlib.h>
#include "sc_conv.cuh"
#include "nm_conv.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <string.h>

using namespace dsgan;
using namespace dsgan::tc;

namespace {
constexpr int TH = 8, TW = 16, BM = TH * TW, BK = 64, EPI_WARPS = 16, NUM_THREADS = 64 + 32 * EPI_WARPS, MAX_TAPS = 16;

struct ConvParams {
  int N, Hg, Wg;            // images, grid extent (positions per image)
  int Ci, Co, co_pad;
  int is_, os_;             // input stride, output stride
  int Ho, Wo;               // output extent (pixels)
  int nclass;               // output parity classes handled by this launch (1, or 4 for stride-2 transposed ops)
  int oy0[4], ox0[4], ntaps[4];
  int dy[4 * MAX_TAPS], dx[4 * MAX_TAPS], slab[4 * MAX_TAPS];
  int tiles_y, tiles_x, n_tiles;
  void* C; int ldc;
  const float* bias;
  void* pre; int ld_pre;
  const void* aux; int ld_aux;
  int act, dact, accumulate;
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);  // keep ~190 KB in flight
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;   // barriers (<= 33 x 8 B + TMEM slot) + base-alignment slack
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// HALO variant (3x3 stride-1 layers with 64 input channels, VGG conv1_2 and its input-gradient): the generic kernel fetches
// the input tile once per tap (nine 16 KB boxes per 128x64 output tile: ncu shows the layer L2->SM bound at 13 % tensor
// activity).  Here a tile is 16 rows x 8 columns of pixels, ONE (16+2) x (8+2) halo is staged per tile (a single 4-D TMA box,
// stored densely) and every tap reads it through the smem descriptor at a different start offset: 8 consecutive pixels of an
// image row are one 8x128-byte swizzle group, SBO = one halo row (1280 B).  All tap weights stay resident in smem.
struct HaloLayout {
  static constexpr int ROWS = 18, COLS = 10;                             // halo rows x pixels, dense: [18][10][64 ch]
  static constexpr int ROW_BYTES = COLS * 128;                           // 1280 B: SBO of the A descriptor (the 128-byte swizzle
                                                                         // is a function of the absolute smem address, so
                                                                         // neither the start nor SBO need 1024-byte alignment)
  static constexpr int TILE_BYTES = ROWS * ROW_BYTES;                    // 23040 B delivered by ONE 4-D TMA box per tile
  static constexpr int STAGE_BYTES = 23552;                              // next multiple of 1024
  static constexpr int STAGES = 4;
  static constexpr int W_BYTES = 9 * 64 * 128;                           // nine 64x64 bf16 slabs
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BAR_OFF = A_OFF + STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

template <int BN, bool HALO = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_tc_conv(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using SL = SmemLayout<BN>;
  constexpr int NSTAGE = HALO ? HaloLayout::STAGES : SL::STAGES;
  constexpr int TILE_H = HALO ? 16 : TH, TILE_W = HALO ? 8 : TW;
  // Epilogue teams: a 128 x BN accumulator is 4 lane quarters x BN/32 column chunks = 4*CHUNKS warp tasks.  With BN = 32 / 64
  // that is 4 / 8 of the 16 epilogue warps, and a tile's epilogue is latency- (not throughput-) bound, so the warps are split
  // into TEAMS that drain different tiles concurrently, and the TMEM ring is deepened to two accumulators per team.
  constexpr int CHUNKS = BN / 32;
  constexpr int TEAMS = CHUNKS >= 4 ? 1 : 4 / CHUNKS;                 // BN 32 -> 4, 64 -> 2, >= 128 -> 1
  constexpr int TEAM_WARPS = EPI_WARPS / TEAMS;                       // warps arriving on an accumulator's "empty" barrier
  constexpr int NACC = 2 * TEAMS;                                     // accumulators in TMEM (NACC * BN <= 512 columns)
  constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (HALO ? HaloLayout::BAR_OFF : SL::BAR_OFF));
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + NACC;
  uint64_t* wfull = tempty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int num_tiles = p.N * tiles_per_img * p.n_tiles * p.nclass;
  const int kc = (p.Ci + BK - 1) / BK;  // channel blocks per tap (TMA zero-fills channels >= Ci)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TEAM_WARPS); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (HALO) {   // all tap weights once, then one halo per tile
        mbar_expect_tx(wfull, (uint32_t)p.ntaps[0] * BN * 128);
        for (int t = 0; t < p.ntaps[0]; ++t) tma_load_2d(smem + t * BN * 128, &tmB, wfull, 0, p.slab[t] * p.co_pad);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int img = tile / tiles_per_img, t2 = tile % tiles_per_img;
          const int y0 = (t2 / p.tiles_x) * TILE_H, x0 = (t2 % p.tiles_x) * TILE_W;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + HaloLayout::A_OFF + stage * HaloLayout::STAGE_BYTES;
          mbar_expect_tx(&full[stage], HaloLayout::TILE_BYTES);
          tma_load_4d(sa, &tmA, &full[stage], 0, x0 - 1, y0 - 1, img);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
      for (int tile = blockIdx.x; !HALO && tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        const int cls = (tile / p.n_tiles) % p.nclass;
        const int sp = tile / (p.n_tiles * p.nclass);
        const int img = sp / tiles_per_img, t2 = sp % tiles_per_img;
        const int y0 = (t2 / p.tiles_x) * TH, x0 = (t2 % p.tiles_x) * TW;
        for (int t = cls * MAX_TAPS; t < cls * MAX_TAPS + p.ntaps[cls]; ++t) {
          const int cy = y0 * p.is_ + p.dy[t], cx = x0 * p.is_ + p.dx[t];
          for (int c = 0; c < kc; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * SL::STAGE_BYTES;
            mbar_expect_tx(&full[stage], SL::STAGE_BYTES);
            tma_load_4d(sa, &tmA, &full[stage], c * BK, cx, cy, img);
            tma_load_2d(sa + SL::A_BYTES, &tmB, &full[stage], c * BK, p.slab[t] * p.co_pad + n_blk * BN);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_bf16(BM, BN, false, false);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (HALO) {
        // tap (dy, dx): the 16x8 pixel window starts at halo row 1+dy, pixel 1+dx.  The nine start offsets are tile-invariant:
        // keep them (in the descriptor's 16-byte units) in registers so that the issuing thread does nothing per tile but
        // wait, add and issue 36 MMAs (it, not the tensor pipe, was the bottleneck of this layer).
        uint32_t aoff[9];
#pragma unroll
        for (int t = 0; t < 9; ++t)
          aoff[t] = (uint32_t)((1 + p.dy[t]) * HaloLayout::ROW_BYTES + (1 + p.dx[t]) * 128) >> 4;
        mbar_wait(wfull, 0);
        tc_fence_after();
        const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem), 16, 1024);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc0 =
              smem_desc_sw128(smem_u32(smem + HaloLayout::A_OFF + stage * HaloLayout::STAGE_BYTES), 16, HaloLayout::ROW_BYTES);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d_tmem, adesc0 + (uint64_t)(aoff[t] + k * 2), bdesc0 + (uint64_t)(t * (BN * 128 >> 4) + k * 2), IDESC,
                        (t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          umma_commit(&tfull[acc]);
          if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
      }
      for (int tile = blockIdx.x; !HALO && tile < num_tiles; tile += gridDim.x) {
        const int nk = p.ntaps[(tile / p.n_tiles) % p.nclass] * kc;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t adesc = smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = smem_desc_sw128(sa + SL::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                          // 0..3
    const int team = CHUNKS >= 4 ? 0 : part / CHUNKS;           // which tiles this warp drains: team, team + TEAMS, ...
    const int c_first = CHUNKS >= 4 ? part : part % CHUNKS, c_step = CHUNKS >= 4 ? 4 : CHUNKS;
    int it = team;                                              // index of the tile in this CTA's sequence
    for (int tile = blockIdx.x + team * gridDim.x; tile < num_tiles; tile += TEAMS * gridDim.x, it += TEAMS) {
      const int acc = it % NACC;
      const uint32_t acc_phase = (uint32_t)(it / NACC) & 1u;
      const int n_blk = tile % p.n_tiles;
      const int cls = (tile / p.n_tiles) % p.nclass;
      const int sp = tile / (p.n_tiles * p.nclass);
      const int img = sp / tiles_per_img, t2 = sp % tiles_per_img;
      const int r = quarter * 32 + lane;
      const int gy = (t2 / p.tiles_x) * TILE_H + r / TILE_W, gx = (t2 % p.tiles_x) * TILE_W + r % TILE_W;
      const int oy = gy * p.os_ + p.oy0[cls], ox = gx * p.os_ + p.ox0[cls];
      const bool row_ok = gy < p.Hg && gx < p.Wg && oy < p.Ho && ox < p.Wo;
      const size_t pix = ((size_t)img * p.Ho + oy) * p.Wo + ox;
      const bool wide = (p.ldc % 16 == 0) && ((uintptr_t)p.C % 32 == 0) &&
                        (!p.pre || (p.ld_pre % 16 == 0 && (uintptr_t)p.pre % 32 == 0)) &&
                        (!p.aux || (p.ld_aux % 16 == 0 && (uintptr_t)p.aux % 32 == 0));
      const bool vec_ok = (p.ldc % 8 == 0) && (!p.aux || p.ld_aux % 8 == 0) && (!p.pre || p.ld_pre % 8 == 0) &&
                          (((uintptr_t)p.C | (uintptr_t)p.aux | (uintptr_t)p.pre) % 16 == 0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_first; c < CHUNKS; c += c_step) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 + 32 <= p.Co && vec_ok) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = __ldg(bp + q);
              f[q * 4] += b4.x; f[q * 4 + 1] += b4.y; f[q * 4 + 2] += b4.z; f[q * 4 + 3] += b4.w;
            }
          }
          bf16* o = reinterpret_cast<bf16*>(p.C) + pix * p.ldc + col0;
          if (p.accumulate) {
            float old[32];
            if (wide) load_row32(o, old);
            else {
              const uint4* op = reinterpret_cast<const uint4*>(o);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = op[q];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  old[q * 8 + e * 2] = __uint_as_float(w[e] << 16);
                  old[q * 8 + e * 2 + 1] = __uint_as_float(w[e] & 0xffff0000u);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += old[j];
          }
          if (p.dact) {
            const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + pix * p.ld_aux + col0);
            float a[32];
            if (wide) load_row32(ap, a);
            else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = __ldg(ap + q);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  a[q * 8 + e * 2] = __uint_as_float(w[e] << 16);
                  a[q * 8 + e * 2 + 1] = __uint_as_float(w[e] & 0xffff0000u);
                }
              }
            }
            act_bwd_fast_mul<32>(p.dact, f, a);
          }
          if (p.pre) {
            uint4* pp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.pre) + pix * p.ld_pre + col0);
            if (wide) store_row32(pp, f);
            else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                pp[q] = make_uint4(bf16x2_bits(f[q * 8], f[q * 8 + 1]), bf16x2_bits(f[q * 8 + 2], f[q * 8 + 3]),
                                   bf16x2_bits(f[q * 8 + 4], f[q * 8 + 5]), bf16x2_bits(f[q * 8 + 6], f[q * 8 + 7]));
            }
          }
          act_fwd_fast_vec<32>(p.act, f);
          if (wide) store_row32(o, f);
          else {
            uint4* op = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              op[q] = make_uint4(bf16x2_bits(f[q * 8], f[q * 8 + 1]), bf16x2_bits(f[q * 8 + 2], f[q * 8 + 3]),
                                 bf16x2_bits(f[q * 8 + 4], f[q * 8 + 5]), bf16x2_bits(f[q * 8 + 6], f[q * 8 + 7]));
          }
        }
        if (row_ok && col0 < p.Co && !(col0 + 32 <= p.Co && vec_ok)) {
          // narrow / unaligned output (Co = 1, 3, 6, 12 ...): guarded scalar epilogue
          const int nv = min(32, p.Co - col0);
          bf16* o = reinterpret_cast<bf16*>(p.C) + pix * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nv) {
              float f = __uint_as_float(v[j]);
              if (p.bias) f += __ldg(p.bias + col0 + j);
              if (p.accumulate) f += __bfloat162float(o[j]);
              if (p.dact) f *= act_bwd_fast(p.dact, __bfloat162float(reinterpret_cast<const bf16*>(p.aux)[pix * p.ld_aux + col0 + j]));
              if (p.pre) reinterpret_cast<bf16*>(p.pre)[pix * p.ld_pre + col0 + j] = __float2bfloat16_rn(f);
              o[j] = __float2bfloat16_rn(act_fwd_fast(p.act, f));
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the same convolutions on tensor cores:
//   dW[slab t][m][n] += sum over grid positions (img, y, x):  G[img, y, x, m] * X[img, y*xs + dy_t, x*xs + dx_t, n]
// with G the tensor indexed on the (small) grid and X the tensor read at the tap offset (stride xs).  Both operands are
// pixel-major in memory, i.e. MN-major UMMA operands whose K dimension is the pixel index: one pipeline stage holds an
// 8x16 pixel patch (K = 128) of G (2 boxes of 64 channels) and of X (BN/64 boxes).  A tile is (tap, m block, n block, split
// of the patch range); partial sums are reduced with fp32 atomics straight into the parameter-gradient tensor, addressed
// with the master weight's own strides (OIHW or IOHW).
constexpr int WG_STAGES = 3, WG_BK = TH * TW;  // 128 pixels per stage

struct WgradParams {
  int N, Hg, Wg;
  int M, Nn;               // channels of G (rows of dW) and of X (cols of dW)
  int xs;                  // stride of X w.r.t. the grid
  int ntaps;
  int dy[MAX_TAPS], dx[MAX_TAPS];
  long long tap_off[MAX_TAPS];   // element offset of tap t inside the master weight gradient
  long long s_m, s_n;            // element strides of dW for the G-channel and X-channel index
  int tiles_y, tiles_x, m_tiles, n_tiles, splits, patches_per_split;
  float* dW;
};

template <int BN>
struct WgSmem {
  static constexpr int A_BYTES = 2 * WG_BK * 128;          // 2 atoms of 64 channels x 128 pixels
  static constexpr int B_BYTES = (BN / 64) * WG_BK * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = WG_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_tc_conv_wgrad(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using SL = WgSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* tfull = empty + WG_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int patches_per_img = p.tiles_y * p.tiles_x;
  const int total_patches = p.N * patches_per_img;
  const int num_tiles = p.ntaps * p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int r = tile;
        const int n_blk = r % p.n_tiles; r /= p.n_tiles;
        const int m_blk = r % p.m_tiles; r /= p.m_tiles;
        const int t = r % p.ntaps;
        const int split = r / p.ntaps;
        const int q0 = split * p.patches_per_split, q1 = min(q0 + p.patches_per_split, total_patches);
        for (int q = q0; q < q1; ++q) {
          const int img = q / patches_per_img, t2 = q % patches_per_img;
          const int y0 = (t2 / p.tiles_x) * TH, x0 = (t2 % p.tiles_x) * TW;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          mbar_expect_tx(&full[stage], SL::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_4d(sa + j * (WG_BK * 128), &tmG, &full[stage], m_blk * BM + j * 64, x0, y0, img);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sb + j * (WG_BK * 128), &tmX, &full[stage], n_blk * BN + j * 64, x0 * p.xs + p.dx[t],
                        y0 * p.xs + p.dy[t], img);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_bf16(BM, BN, true, true);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_tiles * p.m_tiles * p.ntaps);
        const int q0 = split * p.patches_per_split, q1 = min(q0 + p.patches_per_split, total_patches);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int q = q0; q < q1; ++q) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint64_t adesc = smem_desc_sw128(sa, WG_BK * 128, 1024);
          const uint64_t bdesc = smem_desc_sw128(sa + SL::A_BYTES, WG_BK * 128, 1024);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)((k * 2048) >> 4), bdesc + (uint64_t)((k * 2048) >> 4), IDESC,
                      (q > q0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2, nparts = EPI_WARPS / 4;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int r = tile;
      const int n_blk = r % p.n_tiles; r /= p.n_tiles;
      const int m_blk = r % p.m_tiles; r /= p.m_tiles;
      const int t = r % p.ntaps;
      const int row = m_blk * BM + quarter * 32 + lane;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = part; c < BN / 32; c += nparts) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row < p.M) {
          float* o = p.dW + p.tap_off[t] + (long long)row * p.s_m;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.Nn) atomicAdd(o + (long long)(col0 + j) * p.s_n, __uint_as_float(v[j]));
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ---- host ------------------------------------------------------------------------------------------------
struct Key4 {
  const void* base; uint64_t c, w, h, n, ld; uint32_t es;
  bool operator<(const Key4& o) const {
    return std::tie(base, c, w, h, n, ld, es) < std::tie(o.base, o.c, o.w, o.h, o.n, o.ld, o.es);
  }
};
std::map<Key4, CUtensorMap> g_maps4;
struct Key2 {
  const void* base; uint64_t d0, d1; uint32_t b1;
  bool operator<(const Key2& o) const { return std::tie(base, d0, d1, b1) < std::tie(o.base, o.d0, o.d1, o.b1); }
};
std::map<Key2, CUtensorMap> g_maps2;
std::mutex g_mu;

// halo of the HALO variant: box {64 channels, 10 pixels, 18 rows, 1 image}
int map_input_halo(CUtensorMap* out, const void* base, int N, int H, int W, int C, int ld) {
  Key4 k{base, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)ld, 0x7a10u};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps4.find(k);
  if (it != g_maps4.end()) { *out = it->second; return 0; }
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  uint32_t box[4] = {64, (uint32_t)HaloLayout::COLS, (uint32_t)HaloLayout::ROWS, 1};
  if (encode_tmap_bf16(out, base, 4, dims, strides, box, nullptr)) return 1;
  if (g_maps4.size() > 4096) g_maps4.clear();
  g_maps4[k] = *out;
  return 0;
}
int map_input(CUtensorMap* out, const void* base, int N, int H, int W, int C, int ld, int es) {
  Key4 k{base, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)ld, (uint32_t)es};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps4.find(k);
  if (it != g_maps4.end()) { *out = it->second; return 0; }
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  uint32_t box[4] = {64, (uint32_t)(TW * es), (uint32_t)(TH * es), 1};
  uint32_t estr[4] = {1, (uint32_t)es, (uint32_t)es, 1};
  if (encode_tmap_bf16(out, base, 4, dims, strides, box, estr)) return 1;
  if (g_maps4.size() > 4096) g_maps4.clear();
  g_maps4[k] = *out;
  return 0;
}
int map_weight(CUtensorMap* out, const void* base, int Ci, int rows, int bn) {
  Key2 k{base, (uint64_t)Ci, (uint64_t)rows, (uint32_t)bn};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps2.find(k);
  if (it != g_maps2.end()) { *out = it->second; return 0; }
  uint64_t dims[2] = {(uint64_t)Ci, (uint64_t)rows}, strides[1] = {(uint64_t)Ci * 2};
  uint32_t box[2] = {64, (uint32_t)bn};
  if (encode_tmap_bf16(out, base, 2, dims, strides, box, nullptr)) return 1;
  if (g_maps2.size() > 4096) g_maps2.clear();
  g_maps2[k] = *out;
  return 0;
}

int g_sms = 0;
int sms() {
  if (!g_sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, d); }
  return g_sms;
}

template <int BN>
int launch(const CUtensorMap& a, const CUtensorMap& b, const ConvParams& p, cudaStream_t s) {
  static bool attr = false;
  constexpr int smem = SmemLayout<BN>::TOTAL;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("tc_conv smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr = true;
  }
  const long long tiles = (long long)p.N * p.tiles_y * p.tiles_x * p.n_tiles * p.nclass;
  const int grid = tiles < sms() ? (int)tiles : sms();
  k_tc_conv<BN><<<grid, NUM_THREADS, smem, s>>>(a, b, p);
  return DS_LAUNCHED("tc_conv");
}

template <int BN>
int launch_halo(const CUtensorMap& a, const CUtensorMap& b, const ConvParams& p, cudaStream_t s) {
  static bool attr = false;
  constexpr int smem = HaloLayout::TOTAL;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("tc_conv(halo) smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr = true;
  }
  const long long tiles = (long long)p.N * p.tiles_y * p.tiles_x;
  const int grid = tiles < sms() ? (int)tiles : sms();
  k_tc_conv<BN, true><<<grid, NUM_THREADS, smem, s>>>(a, b, p);
  return DS_LAUNCHED("tc_conv_halo");
}

// dst[slab][o][i] (bf16) = src[o*so + i*si + tap_offset(slab)]  — packs OIHW / IOHW fp32 weights into K-major slabs
__global__ void k_pack_slabs(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I, int Op, int Ip,
                             int kh, int kw, long long so, long long si, long long sky, long long skx, int flip) {
  const long long total = (long long)kh * kw * Op * Ip;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ip);
    const int o = (int)((idx / Ip) % Op);
    const int slab = (int)(idx / ((long long)Ip * Op));
    int ky = slab / kw, kx = slab % kw;
    if (flip) { ky = kh - 1 - ky; kx = kw - 1 - kx; }
    const float v = (o < O && i < I) ? src[o * so + i * si + ky * sky + kx * skx] : 0.f;  // zero padding rows/cols
    dst[idx] = __float2bfloat16_rn(v);
  }
}
// all conv-weight slabs of a network in ONE launch.  jobs[] lives in device memory; every block owns PACK_CHUNK consecutive
// output elements of one job (jobs[j].block0 = first block of job j, ascending), so big and small layers get blocks in
// proportion to their size.
constexpr int PACK_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_pack_slabs_batched(const dsgan_pack_job* __restrict__ jobs, int njobs) {
  __shared__ int sj;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {                       // last job whose block0 <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    sj = lo;
  }
  __syncthreads();
  const dsgan_pack_job j = jobs[sj];
  const float* src = reinterpret_cast<const float*>(j.src);
  bf16* dst = reinterpret_cast<bf16*>(j.dst);
  const long long total = (long long)j.kh * j.kw * j.O_pad * j.I_pad;
  const long long base = (long long)((int)blockIdx.x - j.block0) * PACK_CHUNK;
#pragma unroll
  for (int u = 0; u < PACK_CHUNK / 256; ++u) {
    const long long idx = base + u * 256 + threadIdx.x;
    if (idx >= total) break;
    const int i = (int)(idx % j.I_pad);
    const int o = (int)((idx / j.I_pad) % j.O_pad);
    const int slab = (int)(idx / ((long long)j.I_pad * j.O_pad));
    const int ky = slab / j.kw, kx = slab % j.kw;
    const float v = (o < j.O && i < j.I) ? __ldg(src + o * j.s_o + i * j.s_i + ky * j.s_ky + kx * j.s_kx) : 0.f;
    dst[idx] = __float2bfloat16_rn(v);
  }
}
}  // namespace

extern "C" {
int dsgan_pack_conv_weights(const dsgan_pack_job* jobs_dev, int njobs, int total_blocks, void* stream) {
  DS_REQUIRE(jobs_dev && njobs >= 1 && total_blocks >= 1, "pack_conv_weights: bad job table");
  k_pack_slabs_batched<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, njobs);
  return DS_LAUNCHED("pack_conv_weights");
}
int dsgan_pack_conv_weight(const float* src, void* dst, int O, int I, int O_pad, int I_pad, int kh, int kw,
                           long long s_o, long long s_i, long long s_ky, long long s_kx, int flip, void* stream) {
  DS_REQUIRE(O_pad >= O && I_pad >= I && I_pad % 64 == 0, "pack_conv_weight: bad padding %d>=%d %d>=%d", O_pad, O, I_pad, I);
  const long long total = (long long)kh * kw * O_pad * I_pad;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_pack_slabs<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, O, I, O_pad, I_pad, kh, kw, s_o, s_i, s_ky,
                                                                 s_kx, flip);
  return DS_LAUNCHED("pack_conv_weight");
}

int dsgan_tc_conv_supported(int Ci, int Co, int ld_in, int ld_out) {
  // any channel counts: TMA zero-fills input channels >= Ci (the pixel pitch must be 16-byte aligned), the packed weight
  // slabs are zero-padded, and narrow / unaligned outputs take the scalar epilogue
  return Ci >= 1 && Co >= 1 && (ld_in % 8 == 0) && ld_out >= Co;
}

int dsgan_tc_conv(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out,
                  void* pre_out, const void* aux, void* stream) {
  DS_REQUIRE(d && in && w_slabs && out, "tc_conv: null argument");
  DS_REQUIRE(dsgan_tc_conv_supported(d->Ci, d->Co, d->ld_in, d->ld_out), "tc_conv: unsupported Ci=%d Co=%d", d->Ci, d->Co);
  DS_REQUIRE(d->nclass >= 1 && d->nclass <= 4, "tc_conv: nclass=%d", d->nclass);
  for (int c = 0; c < d->nclass; ++c) DS_REQUIRE(d->ntaps[c] >= 1 && d->ntaps[c] <= MAX_TAPS, "tc_conv: ntaps=%d", d->ntaps[c]);
  DS_REQUIRE(d->in_stride == 1 || d->in_stride == 2, "tc_conv: in_stride=%d", d->in_stride);
  DS_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)w_slabs % 16 == 0), "tc_conv: unaligned");
  DS_REQUIRE(d->ci_pad % 64 == 0 && d->ci_pad >= d->Ci && d->co_pad >= d->Co, "tc_conv: bad slab padding");
  DS_REQUIRE(!d->dact || aux, "tc_conv: dact needs aux");
  // 3x3 / stride 1 / 64 input channels / <= 64 output channels (VGG conv1_2 and its input-gradient, the generator's 64 -> 3
  // `res` conv, VGG conv1_1's input-gradient): halo-staged variant (taps read one staged halo at different descriptor offsets)
  bool halo = d->Ci == 64 && d->ci_pad == 64 && ((d->Co == 64 && d->co_pad == 64) || (d->Co <= 32 && d->co_pad == 32)) &&
              d->nclass == 1 && d->in_stride == 1 && d->out_stride == 1 && d->ntaps[0] == 9 && d->oy0[0] == 0 &&
              d->ox0[0] == 0;
  for (int t = 0; halo && t < d->ntaps[0]; ++t) halo = d->dy[t] >= -1 && d->dy[t] <= 1 && d->dx[t] >= -1 && d->dx[t] <= 1;
  if (halo && d->Co <= 16) {   // 64 -> 3: an N = 8 mma.sync tile wastes less than the BN = 32 tcgen05 tile (nm_conv.cu)
    const char* e = getenv("DSGAN_NM_NOUT");
    int rc = 0;
    if (!(e && e[0] == '0') && nm::conv_try(d, in, w_slabs, bias, out, pre_out, aux, stream, &rc)) return rc;
  }
  if (halo) {
    CUtensorMap ta, tb;
    if (map_input_halo(&ta, in, d->N, d->Hi, d->Wi, d->Ci, d->ld_in)) return 1;
    if (map_weight(&tb, w_slabs, d->ci_pad, d->nslabs * d->co_pad, d->co_pad)) return 1;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Ci = d->Ci; p.Co = d->Co; p.co_pad = d->co_pad;
    p.is_ = 1; p.os_ = 1; p.Ho = d->Ho; p.Wo = d->Wo; p.nclass = 1; p.ntaps[0] = d->ntaps[0];
    for (int t = 0; t < d->ntaps[0]; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.slab[t] = d->slab[t]; }
    p.tiles_y = (d->Hg + 15) / 16; p.tiles_x = (d->Wg + 7) / 8; p.n_tiles = 1;
    p.C = out; p.ldc = d->ld_out; p.bias = bias; p.pre = pre_out; p.ld_pre = d->ld_pre; p.aux = aux; p.ld_aux = d->ld_aux;
    p.act = d->act; p.dact = d->dact; p.accumulate = d->accumulate;
    return d->co_pad == 64 ? launch_halo<64>(ta, tb, p, (cudaStream_t)stream) : launch_halo<32>(ta, tb, p, (cudaStream_t)stream);
  }
  {
    int rc = 0;   // 1/3/6/12-channel layers: (tap, padded channel) implicit GEMM on mma.sync (nm_conv.cu), else the CUDA-core
    if (nm::conv_try(d, in, w_slabs, bias, out, pre_out, aux, stream, &rc)) return rc;   // direct convolution (sc_conv.cu)
    if (sc::conv_try(d, in, w_slabs, bias, out, pre_out, aux, stream, &rc)) return rc;
  }
  const int BN = d->Co >= 256 ? 256 : (d->Co >= 128 ? 128 : (d->Co >= 64 ? 64 : 32));
  CUtensorMap ta, tb;
  if (map_input(&ta, in, d->N, d->Hi, d->Wi, d->Ci, d->ld_in, d->in_stride)) return 1;
  if (map_weight(&tb, w_slabs, d->ci_pad, d->nslabs * d->co_pad, BN)) return 1;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Ci = d->Ci; p.Co = d->Co; p.co_pad = d->co_pad;
  p.is_ = d->in_stride; p.os_ = d->out_stride; p.Ho = d->Ho; p.Wo = d->Wo; p.nclass = d->nclass;
  for (int c = 0; c < d->nclass; ++c) {
    p.oy0[c] = d->oy0[c]; p.ox0[c] = d->ox0[c]; p.ntaps[c] = d->ntaps[c];
    for (int t = c * MAX_TAPS; t < c * MAX_TAPS + d->ntaps[c]; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.slab[t] = d->slab[t]; }
  }
  p.tiles_y = (d->Hg + TH - 1) / TH; p.tiles_x = (d->Wg + TW - 1) / TW; p.n_tiles = (d->Co + BN - 1) / BN;
  p.C = out; p.ldc = d->ld_out; p.bias = bias; p.pre = pre_out; p.ld_pre = d->ld_pre; p.aux = aux; p.ld_aux = d->ld_aux;
  p.act = d->act; p.dact = d->dact; p.accumulate = d->accumulate;
  cudaStream_t s = (cudaStream_t)stream;
  if (BN == 256) return launch<256>(ta, tb, p, s);
  if (BN == 128) return launch<128>(ta, tb, p, s);
  if (BN == 64) return launch<64>(ta, tb, p, s);
  return launch<32>(ta, tb, p, s);
}
int dsgan_tc_conv_wgrad_supported(int Cg, int Cx, int ld_g, int ld_x) {
  return Cg >= 1 && Cx >= 1 && ld_g % 8 == 0 && ld_x % 8 == 0;  // TMA zero-fills channels beyond Cg / Cx
}

int dsgan_tc_conv_wgrad(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream) {
  DS_REQUIRE(d && G && X && dW, "tc_conv_wgrad: null argument");
  DS_REQUIRE(dsgan_tc_conv_wgrad_supported(d->Cg, d->Cx, d->ld_g, d->ld_x), "tc_conv_wgrad: unsupported Cg=%d Cx=%d", d->Cg, d->Cx);
  DS_REQUIRE(d->ntaps >= 1 && d->ntaps <= MAX_TAPS && (d->x_stride == 1 || d->x_stride == 2), "tc_conv_wgrad: bad taps/stride");
  DS_REQUIRE(((uintptr_t)G % 16 == 0) && ((uintptr_t)X % 16 == 0), "tc_conv_wgrad: unaligned");
  {
    int rc = 0;
    if (nm::wgrad_try(d, G, X, dW, stream, &rc)) return rc;
    if (sc::wgrad_try(d, G, X, dW, stream, &rc)) return rc;
  }
  const int BN = d->Cx >= 128 ? 128 : 64;
  CUtensorMap tg, tx;
  if (map_input(&tg, G, d->N, d->Hg, d->Wg, d->Cg, d->ld_g, 1)) return 1;
  if (map_input(&tx, X, d->N, d->Hx, d->Wx, d->Cx, d->ld_x, d->x_stride)) return 1;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.M = d->Cg; p.Nn = d->Cx; p.xs = d->x_stride; p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.tap_off[t] = d->tap_off[t]; }
  p.s_m = d->s_g; p.s_n = d->s_x;
  p.tiles_y = (d->Hg + TH - 1) / TH; p.tiles_x = (d->Wg + TW - 1) / TW;
  p.m_tiles = (d->Cg + BM - 1) / BM; p.n_tiles = (d->Cx + BN - 1) / BN;
  const int total_patches = d->N * p.tiles_y * p.tiles_x;
  const int base = d->ntaps * p.m_tiles * p.n_tiles;
  // one wave of CTAs, at least 4 patches (8 k-blocks) per split: every split ends in a BM x BN fp32 atomic reduction
  int splits = (sms() + base - 1) / base;
  if (splits > total_patches / 4) splits = total_patches / 4;
  if (splits < 1) splits = 1;
  p.patches_per_split = (total_patches + splits - 1) / splits;
  p.splits = (total_patches + p.patches_per_split - 1) / p.patches_per_split;
  p.dW = dW;
  cudaStream_t s = (cudaStream_t)stream;
  const long long tiles = (long long)base * p.splits;
  const int grid = tiles < sms() ? (int)tiles : sms();
  if (BN == 128) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_tc_conv_wgrad<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<128>::TOTAL); attr = true; }
    k_tc_conv_wgrad<128><<<grid, NUM_THREADS, WgSmem<128>::TOTAL, s>>>(tg, tx, p);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_tc_conv_wgrad<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<64>::TOTAL); attr = true; }
    k_tc_conv_wgrad<64><<<grid, NUM_THREADS, WgSmem<64>::TOTAL, s>>>(tg, tx, p);
  }
  return DS_LAUNCHED("tc_conv_wgrad");
}
}
