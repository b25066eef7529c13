// Fused ConvNeXt Block MLP (MixConvNeXtML.py:230-243): y = shortcut(x) + pwconv2( GELU( pwconv1(t) + b1 ) ) + b2 with the
// 4C-wide hidden tensor kept ON CHIP: it is produced 128 columns at a time in tensor memory, passes through bias + exact
// GELU in registers, is written as bf16 into shared memory in the K-major / 128-byte-swizzled layout a tcgen05 A operand
// needs, and is consumed by the second GEMM straight from there.  The 1x1 shortcut is a third GEMM into the same output
// accumulator, so the block output is rounded to bf16 exactly once.
//
// One persistent CTA per SM, one 128-row tile of pixels at a time:
//   warp 0      : TMA producer  - T tile (resident for the tile) + a FIFO of 16 KB weight / shortcut-input boxes
//   warp 1      : MMA issuer    - G1(j): H_j = T . W1_j^T   (128 x 128, K = C_in)      -> TMEM H[j & 1]
//                                 SC   : Y   = X . Ws^T     (128 x N_out, K = C_in)    -> TMEM Y
//                                 G2(j): Y  += Hs_j . W2_j^T (128 x N_out, K = 128)     <- smem Hs[j & 1]
//   warps 2..17 : epilogue      - H[j & 1] -> +b1 -> GELU -> bf16 -> Hs[j & 1]; at the end of the tile Y -> +b2 -> bf16 -> HBM
// Issue order per tile: G1(0) G1(1) SC { G1(j+2) G2(j) }: the tensor pipe works on chunk j+1 / j+2 while the epilogue warps
// are busy with chunk j.  HBM traffic per pixel: read t and x, write y -- the hidden never leaves the SM.
#include "tc_common.cuh"
#include "../../include/dsgan_b200.h"
#include <string.h>

using namespace dsgan;
using namespace dsgan::tc;

namespace {

constexpr int BM = 128;
constexpr int BOX = 16384;                 // one TMA box: <= 128 rows x 64 bf16 (128 B, SWIZZLE_128B)
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_LIMIT = 232448;         // 227 KB

template <int CIN, int NOUT>
struct FwdCfg {
  static constexpr int KB1 = CIN / 64;                 // k-blocks of pwconv1 / shortcut
  static constexpr int HID = 4 * CIN;
  static constexpr int NCH = HID / 128;                // hidden chunks of 128 columns (even)
  static constexpr int NB = NOUT < 128 ? NOUT : 128;   // rows of one W2 / Ws box
  static constexpr int NH = NOUT / NB;                 // output halves
  static constexpr int NY = (256 + 2 * NOUT <= 512) ? 2 : 1;
  static constexpr int T_OFF = 0;
  static constexpr int HS_OFF = T_OFF + KB1 * BOX;     // 4 slabs: [buffer][64-column half]
  static constexpr int RING_OFF = HS_OFF + 4 * BOX;
  static constexpr int FIXED = 1024 /*align slack*/ + 1024 /*barriers*/ + 4 * (HID + NOUT);
  static constexpr int RING_MAX = (SMEM_LIMIT - RING_OFF - FIXED) / BOX;
  static constexpr int RING = RING_MAX > 8 ? 8 : RING_MAX;
  static constexpr int BAR_OFF = RING_OFF + RING * BOX;
  static constexpr int BIAS_OFF = BAR_OFF + 1024;
  static constexpr int TOTAL = BIAS_OFF + 4 * (HID + NOUT) + 1024;
  static_assert(RING >= 3, "weight ring too small");
  static_assert(NCH % 2 == 0 && NCH >= 2, "hidden must be a multiple of 256");
  static_assert(TOTAL <= SMEM_LIMIT, "shared memory budget");
};

struct FwdParams {
  int M, m_tiles;
  const float* b1;
  const float* b2;
  void* Y;
  int ld_y;
  int has_sc;
};

// ring cursor shared by the producer and the MMA issuer (same deterministic schedule on both sides)
struct Ring {
  int stage; uint32_t phase;
  __device__ __forceinline__ void advance(int n) { if (++stage == n) { stage = 0; phase ^= 1; } }
};

template <int CIN, int NOUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_mlp_fwd(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmX,
          const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
          const __grid_constant__ CUtensorMap tmWs, const FwdParams p) {
  using C = FwdCfg<CIN, NOUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + C::RING;
  uint64_t* t_full = empty + C::RING;
  uint64_t* t_empty = t_full + 1;
  uint64_t* h_full = t_empty + 1;      // [2]
  uint64_t* h_empty = h_full + 2;      // [2]
  uint64_t* hs_full = h_empty + 2;     // [2]
  uint64_t* hs_empty = hs_full + 2;    // [2]
  uint64_t* y_full = hs_empty + 2;     // [NY]
  uint64_t* y_empty = y_full + 2;      // [NY]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 2);
  float* sb1 = reinterpret_cast<float*>(smem + C::BIAS_OFF);
  float* sb2 = sb1 + C::HID;

  // Role map: the two single-thread roles (TMA producer, MMA issuer) sit on the HIGHEST warp ids.  The warp scheduler
  // favours high warp ids among eligible warps (B300_MICROARCH.md), and with the roles on warps 0/1 the sixteen always-eligible
  // epilogue warps starved the one thread that feeds them.
  const int wr = threadIdx.x >> 5, lane = threadIdx.x & 31;      // physical warp
  const int warp = (wr + 2) % (NUM_THREADS / 32);                // logical: 0 producer, 1 MMA issuer, 2.. epilogue

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmWs);
    for (int i = 0; i < C::RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(t_full, 1); mbar_init(t_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], EPI_WARPS);
      mbar_init(&hs_full[i], EPI_WARPS); mbar_init(&hs_empty[i], 1);
      mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < C::HID; i += NUM_THREADS) sb1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < NOUT; i += NUM_THREADS) sb2[i] = p.b2 ? __ldg(p.b2 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_sc = p.has_sc != 0;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      Ring r{0, 0};
      auto push = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
        mbar_wait(&empty[r.stage], r.phase ^ 1);
        mbar_expect_tx(&full[r.stage], bytes);
        tma_load_2d(smem + C::RING_OFF + r.stage * BOX, map, &full[r.stage], c0, c1);
        r.advance(C::RING);
      };
      auto push_w1 = [&](int j) {
        for (int kb = 0; kb < C::KB1; ++kb) push(&tmW1, kb * 64, j * 128, BOX);
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int m0 = tile * BM;
        mbar_wait(t_empty, (uint32_t)(it & 1) ^ 1);
        mbar_expect_tx(t_full, C::KB1 * BOX);
        for (int kb = 0; kb < C::KB1; ++kb) tma_load_2d(smem + C::T_OFF + kb * BOX, &tmT, t_full, kb * 64, m0);
        push_w1(0);
        push_w1(1);
        if (has_sc) {
          for (int kb = 0; kb < C::KB1; ++kb) {
            push(&tmX, kb * 64, m0, BOX);
            for (int h = 0; h < C::NH; ++h) push(&tmWs, kb * 64, h * 128, C::NB * 128);
          }
        }
        for (int j = 0; j < C::NCH; ++j) {
          if (j + 2 < C::NCH) push_w1(j + 2);
          for (int s = 0; s < 2; ++s)
            for (int h = 0; h < C::NH; ++h) push(&tmW2, j * 128 + s * 64, h * 128, C::NB * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC_H = idesc_bf16(BM, 128, false, false);
      constexpr uint32_t IDESC_Y = idesc_bf16(BM, C::NB, false, false);
      Ring r{0, 0};
      const uint32_t ring_base = smem_u32(smem + C::RING_OFF);
      const uint32_t t_base = smem_u32(smem + C::T_OFF);
      const uint32_t hs_base = smem_u32(smem + C::HS_OFF);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const uint32_t y_tmem = tmem_base + 256 + (uint32_t)(it % C::NY) * NOUT;
        auto g1 = [&](int j) {   // H[j & 1] = T . W1_j^T
          const int buf = j & 1;
          const uint32_t u = (uint32_t)(it * (C::NCH / 2) + (j >> 1));
          mbar_wait(&h_empty[buf], (u & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + buf * 128;
          for (int kb = 0; kb < C::KB1; ++kb) {
            mbar_wait(&full[r.stage], r.phase);
            tc_fence_after();
            const uint64_t ad = smem_desc_sw128(t_base + kb * BOX, 16, 1024);
            const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_H, (kb | k) ? 1u : 0u);
            umma_commit(&empty[r.stage]);
            r.advance(C::RING);
          }
          umma_commit(&h_full[buf]);
        };
        mbar_wait(t_full, (uint32_t)(it & 1));
        tc_fence_after();
        g1(0);
        g1(1);
        if (C::NCH == 2) umma_commit(t_empty);
        // the output accumulator of this tile must have been drained by the epilogue of tile it - NY
        mbar_wait(&y_empty[it % C::NY], (uint32_t)((it / C::NY) & 1) ^ 1);
        tc_fence_after();
        if (has_sc) {
          for (int kb = 0; kb < C::KB1; ++kb) {
            mbar_wait(&full[r.stage], r.phase);
            const int xs = r.stage;
            r.advance(C::RING);
            const uint64_t ad = smem_desc_sw128(ring_base + xs * BOX, 16, 1024);
            for (int h = 0; h < C::NH; ++h) {
              mbar_wait(&full[r.stage], r.phase);
              tc_fence_after();
              const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(y_tmem + h * 128, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_Y, (kb | k) ? 1u : 0u);
              umma_commit(&empty[r.stage]);
              r.advance(C::RING);
            }
            umma_commit(&empty[xs]);
          }
        }
        for (int j = 0; j < C::NCH; ++j) {
          if (j + 2 < C::NCH) {
            g1(j + 2);
            if (j + 3 == C::NCH) umma_commit(t_empty);   // last G1 of the tile issued: T may be overwritten once it completes
          }
          const int buf = j & 1;
          const uint32_t u = (uint32_t)(it * (C::NCH / 2) + (j >> 1));
          mbar_wait(&hs_full[buf], u & 1);
          tc_fence_after();
          for (int s = 0; s < 2; ++s) {
            const uint64_t ad = smem_desc_sw128(hs_base + (buf * 2 + s) * BOX, 16, 1024);
            for (int h = 0; h < C::NH; ++h) {
              mbar_wait(&full[r.stage], r.phase);
              tc_fence_after();
              const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(y_tmem + h * 128, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_Y,
                          (has_sc || j > 0 || s > 0 || k > 0) ? 1u : 0u);
              umma_commit(&empty[r.stage]);
              r.advance(C::RING);
            }
          }
          umma_commit(&hs_empty[buf]);
        }
        umma_commit(&y_full[it % C::NY]);
      }
    }
  } else {
    // ================= epilogue warps =================
    // Every warp owns a 32-row x 32-column piece of each hidden chunk and walks it in 16-column stages.  The tensor-memory
    // load of stage s+1 (also across chunk boundaries: the other accumulator buffer has long been complete) is issued BEFORE
    // the GELU arithmetic of stage s: TMEM drains at only ~64 B/clk/SM, and with load -> wait -> compute in sequence and all
    // 16 warps in lock-step that drain was a serial quarter of every chunk.
    const int widx = warp - 2;               // == physical warp id
    const int quarter = wr & 3;              // TMEM lane quarter this warp may access (physical warp id % 4)
    const int part = widx >> 2;              // 32-column group inside a 128-column chunk / tile-output column group
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint8_t* hs = smem + C::HS_OFF;
    const int sw = row_in_tile & 7;
    // one 16-column stage: + b1, GELU, bf16, two 16-byte chunks of the K-major / 128B-swizzled slab row
    auto consume = [&](int j, int sub, const uint32_t (&v)[16]) {
      float f[16];
      const float4* bp = reinterpret_cast<const float4*>(sb1 + j * 128 + part * 32 + sub * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b4 = bp[q];
        f[q * 4] = __uint_as_float(v[q * 4]) + b4.x; f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + b4.y;
        f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + b4.z; f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + b4.w;
      }
      act_fwd_fast_vec<16>(ACT_GELU, f);
      uint8_t* rowp = hs + ((j & 1) * 2 + (part >> 1)) * BOX + row_in_tile * 128;
      const int cbase = (part & 1) * 4 + sub * 2;
#pragma unroll
      for (int c = 0; c < 2; ++c)
        *reinterpret_cast<uint4*>(rowp + (((cbase + c) ^ sw) << 4)) =
            make_uint4(bf16x2_bits(f[c * 8], f[c * 8 + 1]), bf16x2_bits(f[c * 8 + 2], f[c * 8 + 3]),
                       bf16x2_bits(f[c * 8 + 4], f[c * 8 + 5]), bf16x2_bits(f[c * 8 + 6], f[c * 8 + 7]));
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      uint32_t va[16], vb[16];
      mbar_wait(&h_full[0], (uint32_t)(it * (C::NCH / 2)) & 1);
      tc_fence_after();
      tmem_ld_32x16(tmem_base + lane_addr + part * 32, va);
#pragma unroll 1
      for (int j = 0; j < C::NCH; ++j) {
        const int buf = j & 1;
        const uint32_t u = (uint32_t)(it * (C::NCH / 2) + (j >> 1));
        tmem_ld_wait();                                                           // va = stage (j, 0)
        tmem_ld_32x16(tmem_base + lane_addr + buf * 128 + part * 32 + 16, vb);    // prefetch stage (j, 1)
        mbar_wait(&hs_empty[buf], (u & 1) ^ 1);                                   // Hs[buf] is free once G2(j - 2) has read it
        consume(j, 0, va);
        tmem_ld_wait();                                                           // vb ready: this warp is done with H[buf]
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[buf]);                                // G1(j + 2) may overwrite the accumulator
        if (j + 1 < C::NCH) {                                                     // prefetch stage (j + 1, 0) from the other buffer
          mbar_wait(&h_full[buf ^ 1], (uint32_t)(it * (C::NCH / 2) + ((j + 1) >> 1)) & 1);
          tc_fence_after();
          tmem_ld_32x16(tmem_base + lane_addr + (buf ^ 1) * 128 + part * 32, va);
        }
        consume(j, 1, vb);
        fence_proxy_async();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&hs_full[buf]);
      }
      // ---- tile output: Y + b2 -> bf16 -> HBM ----
      const int yb = it % C::NY;
      mbar_wait(&y_full[yb], (uint32_t)((it / C::NY) & 1));
      tc_fence_after();
      const int row = tile * BM + row_in_tile;
#pragma unroll 1
      for (int g = part; g < NOUT / 32; g += 4) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_addr + 256 + yb * NOUT + g * 32, v);
        tmem_ld_wait();
        if (row < p.M) {
          float f[32];
          const float4* bp = reinterpret_cast<const float4*>(sb2 + g * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = bp[q];
            f[q * 4] = __uint_as_float(v[q * 4]) + b4.x; f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + b4.y;
            f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + b4.z; f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + b4.w;
          }
          store_row32(reinterpret_cast<bf16*>(p.Y) + (size_t)row * p.ld_y + g * 32, f);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_empty[yb]);
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int CIN, int NOUT>
int launch_fwd(const void* T, int ld_t, const void* X, int ld_x, long long M, const void* W1, const float* b1,
               const void* W2, const float* b2, const void* Ws, void* Y, int ld_y, cudaStream_t s) {
  using C = FwdCfg<CIN, NOUT>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd<CIN, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
    if (e != cudaSuccess) { set_error("fused_mlp_fwd smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr = true;
  }
  CUtensorMap tT, tX, tW1, tW2, tWs;
  if (get_map_2d(&tT, T, CIN, (uint64_t)M, (uint64_t)ld_t, 64, 128)) return 1;
  if (get_map_2d(&tX, X ? X : T, CIN, (uint64_t)M, (uint64_t)(X ? ld_x : ld_t), 64, 128)) return 1;
  if (get_map_2d(&tW1, W1, CIN, C::HID, CIN, 64, 128)) return 1;
  if (get_map_2d(&tW2, W2, C::HID, NOUT, C::HID, 64, C::NB)) return 1;
  if (get_map_2d(&tWs, Ws ? Ws : W1, CIN, Ws ? NOUT : C::HID, CIN, 64, C::NB)) return 1;
  FwdParams p;
  p.M = (int)M; p.m_tiles = (int)((M + BM - 1) / BM);
  p.b1 = b1; p.b2 = b2; p.Y = Y; p.ld_y = ld_y; p.has_sc = (X && Ws) ? 1 : 0;
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  k_mlp_fwd<CIN, NOUT><<<grid, NUM_THREADS, C::TOTAL, s>>>(tT, tX, tW1, tW2, tWs, p);
  return DS_LAUNCHED("fused_mlp_fwd");
}


// =====================================================================================================================
// Backward of the same MLP with the hidden RECOMPUTED (nothing of the 4C tensor was kept by the forward pass):
//   per 64-column hidden chunk j of a 128-pixel tile
//     R(j) : Hpre_j = T . W1_j^T              (K = C_in)    \  two accumulators, one TMEM buffer pair
//     D(j) : dH_j   = dY . W2[:, j]           (K = N_out)   /
//     epilogue: h = Hpre + b1;  A_j = GELU(h);  G_j = dH_j * GELU'(h)   -> bf16 slabs in shared memory
//     DT(j): dT    += G_j . W1_j               (K = 64, N = C_in)  <- slab as the A operand
//     A_j goes straight from the registers to A[M, 4C]; store warp: TMA-stores the G slab to G[M, 4C] (the operands of the
//     two weight-gradient GEMMs); the slab ring is four deep so that store / column sums never stall the epilogue
//     column-sum warp: adds the G slab's column sums to the CTA's partial pwconv1 bias gradient (flushed once per CTA)
//   tile end: dT -> bf16 -> HBM.
// The 4C tensor crosses HBM four times per Block and step (G and A written here, read once each by the weight gradients)
// instead of nine; its epilogue writes never leave the SM as register stores (slab -> TMA store).
constexpr int BOX8 = 8192;                 // ring box: 64 rows x 64 bf16
constexpr int BWD_THREADS = 128 + 32 * EPI_WARPS;   // producer, MMA issuer, store warp, column-sum warp + 16 epilogue warps

template <int CIN, int NOUT>
struct BwdCfg {
  static constexpr int KB1 = CIN / 64;
  static constexpr int KBO = NOUT / 64;
  static constexpr int HID = 4 * CIN;
  static constexpr int NCH = HID / 64;                  // hidden chunks of 64 columns
  static constexpr int T_OFF = 0;
  static constexpr int DY_OFF = T_OFF + KB1 * BOX;
  static constexpr int NGS = 4;                         // G slab buffers: the TMA store / column sums of chunk j may lag
  static constexpr int GS_OFF = DY_OFF + KBO * BOX;     //   behind the epilogue by up to three chunks
  static constexpr int RING_OFF = GS_OFF + NGS * BOX;
  static constexpr int FIXED = 1024 + 1024 + 1024 + 8 * HID;
  static constexpr int RING_MAX = (SMEM_LIMIT - RING_OFF - FIXED) / BOX8;
  static constexpr int RING = RING_MAX > 16 ? 16 : RING_MAX;
  static constexpr int BAR_OFF = RING_OFF + RING * BOX8;
  static constexpr int BIAS_OFF = BAR_OFF + 1024;
  static constexpr int TOTAL = BIAS_OFF + 8 * HID + 1024 + 1024;   // b1 + the CTA's partial bias gradients (pwconv1: HID, pwconv2: <= 256)
  static_assert(RING >= 4, "weight ring too small");
  static_assert(NCH % NGS == 0, "hidden must be a multiple of 256");
  static_assert(256 + CIN <= 512, "tensor memory budget");
  static_assert(TOTAL <= SMEM_LIMIT, "shared memory budget");
};

struct BwdParams {
  int M, m_tiles;
  const float* b1;
  void* dT;
  int ld_dt;
  float* db1;
  float* db2;   // pwconv2 bias gradient = column sums of dY (may be null)
  void* A;      // bf16 [M, HID]: written straight from the epilogue registers (one 32-byte store per lane and chunk)
};

template <int CIN, int NOUT>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_mlp_bwd(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmDY,
          const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
          const __grid_constant__ CUtensorMap tmG, const BwdParams p) {
  using C = BwdCfg<CIN, NOUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + C::RING;
  uint64_t* td_full = empty + C::RING;
  uint64_t* td_empty = td_full + 1;
  uint64_t* hd_full = td_empty + 1;    // [2] both accumulators of a chunk complete
  uint64_t* hd_empty = hd_full + 2;    // [2]
  uint64_t* gs_full = hd_empty + 2;    // [NGS] slab written
  uint64_t* gs_empty = gs_full + 4;    // [NGS] slab read by DT(j), by the TMA store and by the column-sum warp
  uint64_t* dt_full = gs_empty + 4;
  uint64_t* dt_empty = dt_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dt_empty + 1);
  float* sb1 = reinterpret_cast<float*>(smem + C::BIAS_OFF);
  float* sdb = sb1 + C::HID;
  float* sdb2 = sdb + C::HID;         // [NOUT]

  const int wr = threadIdx.x >> 5, lane = threadIdx.x & 31;      // physical warp; helper roles on the highest ids (see forward)
  const int warp = (wr + 4) % (BWD_THREADS / 32);                // logical: 0 producer, 1 MMA, 2 store, 3 column sums, 4.. epilogue

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmG);
    for (int i = 0; i < C::RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(td_full, 1); mbar_init(td_empty, 2);   // released by the MMA issuer's commit and by the column-sum warp
    for (int i = 0; i < 2; ++i) { mbar_init(&hd_full[i], 1); mbar_init(&hd_empty[i], EPI_WARPS); }
    for (int i = 0; i < C::NGS; ++i) {
      mbar_init(&gs_full[i], EPI_WARPS); mbar_init(&gs_empty[i], 3);   // DT(j) commit + store warp + column-sum warp
    }
    mbar_init(dt_full, 1); mbar_init(dt_empty, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < C::HID; i += BWD_THREADS) { sb1[i] = __ldg(p.b1 + i); sdb[i] = 0.f; }
  for (int i = threadIdx.x; i < NOUT; i += BWD_THREADS) sdb2[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      Ring r{0, 0};
      auto push = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(&empty[r.stage], r.phase ^ 1);
        mbar_expect_tx(&full[r.stage], BOX8);
        tma_load_2d(smem + C::RING_OFF + r.stage * BOX8, map, &full[r.stage], c0, c1);
        r.advance(C::RING);
      };
      auto push_rd = [&](int j) {
        for (int kb = 0; kb < C::KB1; ++kb) push(&tmW1, kb * 64, j * 64);   // W1_j as [n = hidden][k = c_in]  (K-major B)
        for (int kb = 0; kb < C::KBO; ++kb) push(&tmW2, j * 64, kb * 64);   // W2 as   [k = n_out][n = hidden] (MN-major B)
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int m0 = tile * BM;
        mbar_wait(td_empty, (uint32_t)(it & 1) ^ 1);
        mbar_expect_tx(td_full, (C::KB1 + C::KBO) * BOX);
        for (int kb = 0; kb < C::KB1; ++kb) tma_load_2d(smem + C::T_OFF + kb * BOX, &tmT, td_full, kb * 64, m0);
        for (int kb = 0; kb < C::KBO; ++kb) tma_load_2d(smem + C::DY_OFF + kb * BOX, &tmDY, td_full, kb * 64, m0);
        push_rd(0);
        push_rd(1);
        for (int j = 0; j < C::NCH; ++j) {
          if (j + 2 < C::NCH) push_rd(j + 2);
          for (int nb = 0; nb < C::KB1; ++nb) push(&tmW1, nb * 64, j * 64);  // W1_j as [k = hidden][n = c_in] (MN-major B)
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC_KK = idesc_bf16(BM, 64, false, false);   // B K-major
      constexpr uint32_t IDESC_KM = idesc_bf16(BM, 64, false, true);    // B MN-major
      Ring r{0, 0};
      const uint32_t ring_base = smem_u32(smem + C::RING_OFF);
      const uint32_t t_base = smem_u32(smem + C::T_OFF), dy_base = smem_u32(smem + C::DY_OFF);
      const uint32_t gs_base = smem_u32(smem + C::GS_OFF);
      const uint32_t dt_tmem = tmem_base + 256;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        auto rd = [&](int j) {
          const int buf = j & 1;
          const uint32_t u = (uint32_t)(it * (C::NCH / 2) + (j >> 1));
          mbar_wait(&hd_empty[buf], (u & 1) ^ 1);
          tc_fence_after();
          const uint32_t dh = tmem_base + buf * 128;
          for (int kb = 0; kb < C::KB1; ++kb) {       // Hpre_j = T . W1_j^T
            mbar_wait(&full[r.stage], r.phase);
            tc_fence_after();
            const uint64_t ad = smem_desc_sw128(t_base + kb * BOX, 16, 1024);
            const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX8, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(dh, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_KK, (kb | k) ? 1u : 0u);
            umma_commit(&empty[r.stage]);
            r.advance(C::RING);
          }
          for (int kb = 0; kb < C::KBO; ++kb) {       // dH_j = dY . W2[:, j]
            mbar_wait(&full[r.stage], r.phase);
            tc_fence_after();
            const uint64_t ad = smem_desc_sw128(dy_base + kb * BOX, 16, 1024);
            const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX8, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(dh + 64, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 128), IDESC_KM, (kb | k) ? 1u : 0u);
            umma_commit(&empty[r.stage]);
            r.advance(C::RING);
          }
          umma_commit(&hd_full[buf]);
        };
        mbar_wait(td_full, (uint32_t)(it & 1));
        tc_fence_after();
        rd(0);
        rd(1);
        if (C::NCH == 2) umma_commit(td_empty);
        mbar_wait(dt_empty, (uint32_t)(it & 1) ^ 1);   // the previous tile's dT has been drained
        tc_fence_after();
        for (int j = 0; j < C::NCH; ++j) {
          if (j + 2 < C::NCH) {
            rd(j + 2);
            if (j + 3 == C::NCH) umma_commit(td_empty);
          }
          const int gb = j & (C::NGS - 1);
          const uint32_t ug = (uint32_t)(it * (C::NCH / C::NGS) + j / C::NGS);
          mbar_wait(&gs_full[gb], ug & 1);
          tc_fence_after();
          const uint64_t ad = smem_desc_sw128(gs_base + gb * BOX, 16, 1024);
          for (int nb = 0; nb < C::KB1; ++nb) {       // dT[:, nb*64 ..] += G_j . W1_j[:, nb*64 ..]
            mbar_wait(&full[r.stage], r.phase);
            tc_fence_after();
            const uint64_t bd = smem_desc_sw128(ring_base + r.stage * BOX8, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(dt_tmem + nb * 64, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 128), IDESC_KM, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[r.stage]);
            r.advance(C::RING);
          }
          umma_commit(&gs_empty[gb]);
        }
        umma_commit(dt_full);
      }
    }
  } else if (warp == 2) {
    // ================= store warp: slabs -> HBM through TMA =================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int m0 = tile * BM;
        for (int j = 0; j < C::NCH; ++j) {
          const int gb = j & (C::NGS - 1);
          const uint32_t ug = (uint32_t)(it * (C::NCH / C::NGS) + j / C::NGS);
          mbar_wait(&gs_full[gb], ug & 1);
          tma_store_2d(&tmG, smem + C::GS_OFF + gb * BOX, j * 64, m0);
          tma_store_commit();
          tma_store_wait_read();               // the slab has been read: it may be overwritten
          mbar_arrive(&gs_empty[gb]);
        }
      }
      tma_store_wait_all();
    }
  } else if (warp == 3) {
    // ================= column-sum warp: pwconv1 bias gradient = column sums of G =================
    // lane = (16-byte chunk c = 8 columns, row offset ro): one LDS.128 per 4 rows, conflict-free under the 128-byte swizzle
    const int c = lane & 7, ro = lane >> 3;
    // column sums of one [128 x 64] bf16 slab (128-byte swizzled rows) added to dst[64]
    auto slab_sums = [&](const uint8_t* gsl, float* dst) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll 8
      for (int i = 0; i < BM / 4; ++i) {
        const int rr = i * 4 + ro;
        const uint4 w4 = *reinterpret_cast<const uint4*>(gsl + rr * 128 + ((c ^ (rr & 7)) << 4));
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[2 * e] += __uint_as_float(w[e] << 16);
          v[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
      }
      if (ro == 0) {
        float* d = dst + c * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += v[e];
      }
      __syncwarp();
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      // pwconv2 bias gradient: the dY tile is resident for the whole tile; rows past M arrive as zeros
      mbar_wait(td_full, (uint32_t)(it & 1));
      if (p.db2) {
#pragma unroll 1
        for (int kb = 0; kb < C::KBO; ++kb) slab_sums(smem + C::DY_OFF + kb * BOX, sdb2 + kb * 64);
      }
      if (lane == 0) mbar_arrive(td_empty);
#pragma unroll 1
      for (int j = 0; j < C::NCH; ++j) {
        const int gb = j & (C::NGS - 1);
        const uint32_t ug = (uint32_t)(it * (C::NCH / C::NGS) + j / C::NGS);
        mbar_wait(&gs_full[gb], ug & 1);
        slab_sums(smem + C::GS_OFF + gb * BOX, sdb + j * 64);
        if (lane == 0) mbar_arrive(&gs_empty[gb]);
      }
    }
    __syncwarp();
    if (p.db1) {
      for (int i = lane; i < C::HID; i += 32) atomicAdd(p.db1 + i, sdb[i]);
    }
    if (p.db2) {
      for (int i = lane; i < NOUT; i += 32) atomicAdd(p.db2 + i, sdb2[i]);
    }
  } else {
    // ================= epilogue warps: 32 rows x 16 columns of every chunk per warp, in two 8-column stages whose
    // tensor-memory loads run one stage ahead of the arithmetic (see the forward kernel) =================
    const int widx = warp - 4;               // == physical warp id
    const int quarter = wr & 3;
    const int part = widx >> 2;              // 16-column group inside a 64-column chunk / tile-output column group
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int sw = row_in_tile & 7;
    // one stage: 8 columns of Hpre (vh) and dH (vd) -> A = GELU(h) (4 packed words) and G = dH * GELU'(h) (one slab chunk)
    auto consume = [&](int j, int sub, const uint32_t (&vh)[8], const uint32_t (&vd)[8], uint32_t* wa, uint8_t* grow) {
      const float2* bp = reinterpret_cast<const float2*>(sb1 + j * 64 + part * 16 + sub * 8);
      uint32_t wg[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 b2v = bp[e];
        const float2 h = make_float2(__uint_as_float(vh[2 * e]) + b2v.x, __uint_as_float(vh[2 * e + 1]) + b2v.y);
        float2 cdf, pdf;
        gelu_parts_fast2(h, cdf, pdf);
        const float2 dg = make_float2(fmaf(h.x, pdf.x, cdf.x), fmaf(h.y, pdf.y, cdf.y));   // GELU'(h) = Phi(h) + h phi(h)
        wa[e] = bf16x2_bits(h.x * cdf.x, h.y * cdf.y);
        wg[e] = bf16x2_bits(__uint_as_float(vd[2 * e]) * dg.x, __uint_as_float(vd[2 * e + 1]) * dg.y);
      }
      *reinterpret_cast<uint4*>(grow + (((part * 2 + sub) ^ sw) << 4)) = make_uint4(wg[0], wg[1], wg[2], wg[3]);
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int row = tile * BM + row_in_tile;
      uint32_t ha[8], da[8], hb[8], db[8];
      mbar_wait(&hd_full[0], (uint32_t)(it * (C::NCH / 2)) & 1);
      tc_fence_after();
      tmem_ld_32x8(tmem_base + lane_addr + part * 16, ha);
      tmem_ld_32x8(tmem_base + lane_addr + 64 + part * 16, da);
#pragma unroll 1
      for (int j = 0; j < C::NCH; ++j) {
        const int buf = j & 1;
        const int gb = j & (C::NGS - 1);
        const uint32_t ug = (uint32_t)(it * (C::NCH / C::NGS) + j / C::NGS);
        uint8_t* grow = smem + C::GS_OFF + gb * BOX + row_in_tile * 128;
        uint32_t wa[8];
        tmem_ld_wait();                                                          // stage (j, 0)
        tmem_ld_32x8(tmem_base + lane_addr + buf * 128 + part * 16 + 8, hb);     // prefetch stage (j, 1)
        tmem_ld_32x8(tmem_base + lane_addr + buf * 128 + 64 + part * 16 + 8, db);
        mbar_wait(&gs_empty[gb], (ug & 1) ^ 1);
        consume(j, 0, ha, da, wa, grow);
        tmem_ld_wait();                                                          // this warp is done with the accumulators
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hd_empty[buf]);
        if (j + 1 < C::NCH) {                                                    // prefetch stage (j + 1, 0) from the other buffer
          mbar_wait(&hd_full[buf ^ 1], (uint32_t)(it * (C::NCH / 2) + ((j + 1) >> 1)) & 1);
          tc_fence_after();
          tmem_ld_32x8(tmem_base + lane_addr + (buf ^ 1) * 128 + part * 16, ha);
          tmem_ld_32x8(tmem_base + lane_addr + (buf ^ 1) * 128 + 64 + part * 16, da);
        }
        consume(j, 1, hb, db, wa + 4, grow);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gs_full[gb]);
        // A = GELU(Hpre): operand of the pwconv2 weight gradient, straight to HBM (32 bytes per lane).  AFTER the fence: the
        // proxy fence is a CTA-wide memory barrier and would otherwise wait for this global store to be acknowledged.
        if (row < p.M) {
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = wa[e];
          st_global_v8(reinterpret_cast<bf16*>(p.A) + (size_t)row * C::HID + j * 64 + part * 16, w8);
        }
      }
      // ---- tile output: dT -> bf16 -> HBM ----
      mbar_wait(dt_full, (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll 1
      for (int g = part; g < CIN / 32; g += 4) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_addr + 256 + g * 32, v);
        tmem_ld_wait();
        if (row < p.M) {
          float f[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
          store_row32(reinterpret_cast<bf16*>(p.dT) + (size_t)row * p.ld_dt + g * 32, f);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dt_empty);
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int CIN, int NOUT>
int launch_bwd(const void* T, int ld_t, const void* dY, int ld_dy, long long M, const void* W1, const float* b1,
               const void* W2, void* dT, int ld_dt, void* G, void* A, float* db1, float* db2, cudaStream_t s) {
  using C = BwdCfg<CIN, NOUT>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd<CIN, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
    if (e != cudaSuccess) { set_error("fused_mlp_bwd smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr = true;
  }
  CUtensorMap tT, tDY, tW1, tW2, tG;
  if (get_map_2d(&tT, T, CIN, (uint64_t)M, (uint64_t)ld_t, 64, 128)) return 1;
  if (get_map_2d(&tDY, dY, NOUT, (uint64_t)M, (uint64_t)ld_dy, 64, 128)) return 1;
  if (get_map_2d(&tW1, W1, CIN, C::HID, CIN, 64, 64)) return 1;
  if (get_map_2d(&tW2, W2, C::HID, NOUT, C::HID, 64, 64)) return 1;
  if (get_map_2d(&tG, G, C::HID, (uint64_t)M, C::HID, 64, 128)) return 1;
  BwdParams p;
  p.M = (int)M; p.m_tiles = (int)((M + BM - 1) / BM);
  p.b1 = b1; p.dT = dT; p.ld_dt = ld_dt; p.db1 = db1; p.db2 = db2; p.A = A;
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  k_mlp_bwd<CIN, NOUT><<<grid, BWD_THREADS, C::TOTAL, s>>>(tT, tDY, tW1, tW2, tG, p);
  return DS_LAUNCHED("fused_mlp_bwd");
}

}  // namespace

extern "C" {
int dsgan_fused_mlp_supported(int Cin, int Nout) {
  // (256, 256) is no generator Block and leaves the backward kernel a 3-box weight ring: not built
  return (Cin == 64 || Cin == 128 || Cin == 256) && (Nout == 64 || Nout == 128 || Nout == 256) && !(Cin == 256 && Nout == 256) ? 1 : 0;
}

int dsgan_fused_mlp_fwd(const void* T, int ld_t, const void* X, int ld_x, long long M, int Cin, int Nout, const void* W1,
                        const float* b1, const void* W2, const float* b2, const void* Ws, void* Y, int ld_y,
                        void* stream) {
  DS_REQUIRE(dsgan_fused_mlp_supported(Cin, Nout), "fused_mlp_fwd: unsupported channels Cin=%d Nout=%d", Cin, Nout);
  DS_REQUIRE(M >= 1 && ld_t % 8 == 0 && ld_y % 16 == 0 && (!X || ld_x % 8 == 0), "fused_mlp_fwd: bad pitches");
  DS_REQUIRE(((uintptr_t)T % 16 == 0) && ((uintptr_t)W1 % 16 == 0) && ((uintptr_t)W2 % 16 == 0) && ((uintptr_t)Y % 32 == 0) &&
                 (!X || (uintptr_t)X % 16 == 0) && (!Ws || (uintptr_t)Ws % 16 == 0),
             "fused_mlp_fwd: unaligned pointer");
  DS_REQUIRE((X == nullptr) == (Ws == nullptr), "fused_mlp_fwd: shortcut needs both X and Ws");
  DS_REQUIRE(b1 != nullptr, "fused_mlp_fwd: pwconv1 bias required");
  cudaStream_t s = (cudaStream_t)stream;
#define MLP_CASE(CI, NO) if (Cin == CI && Nout == NO) return launch_fwd<CI, NO>(T, ld_t, X, ld_x, M, W1, b1, W2, b2, Ws, Y, ld_y, s);
  MLP_CASE(64, 64) MLP_CASE(64, 128) MLP_CASE(64, 256)
  MLP_CASE(128, 64) MLP_CASE(128, 128) MLP_CASE(128, 256)
  MLP_CASE(256, 64) MLP_CASE(256, 128)
#undef MLP_CASE
  set_error("fused_mlp_fwd: no kernel");
  return 1;
}

int dsgan_fused_mlp_bwd(const void* T, int ld_t, const void* dY, int ld_dy, long long M, int Cin, int Nout, const void* W1,
                        const float* b1, const void* W2, void* dT, int ld_dt, void* G, void* A, float* db1, float* db2,
                        void* stream) {
  DS_REQUIRE(dsgan_fused_mlp_supported(Cin, Nout), "fused_mlp_bwd: unsupported channels Cin=%d Nout=%d", Cin, Nout);
  DS_REQUIRE(M >= 1 && ld_t % 8 == 0 && ld_dy % 8 == 0 && ld_dt % 16 == 0, "fused_mlp_bwd: bad pitches");
  DS_REQUIRE(((uintptr_t)T % 16 == 0) && ((uintptr_t)dY % 16 == 0) && ((uintptr_t)W1 % 16 == 0) && ((uintptr_t)W2 % 16 == 0) &&
                 ((uintptr_t)dT % 32 == 0) && ((uintptr_t)G % 16 == 0) && ((uintptr_t)A % 32 == 0),
             "fused_mlp_bwd: unaligned pointer");
  DS_REQUIRE(b1 && G && A && dT, "fused_mlp_bwd: null argument");
  cudaStream_t s = (cudaStream_t)stream;
#define MLP_CASE(CI, NO) if (Cin == CI && Nout == NO) return launch_bwd<CI, NO>(T, ld_t, dY, ld_dy, M, W1, b1, W2, dT, ld_dt, G, A, db1, db2, s);
  MLP_CASE(64, 64) MLP_CASE(64, 128) MLP_CASE(64, 256)
  MLP_CASE(128, 64) MLP_CASE(128, 128) MLP_CASE(128, 256)
  MLP_CASE(256, 64) MLP_CASE(256, 128)
#undef MLP_CASE
  set_error("fused_mlp_bwd: no kernel");
  return 1;
}
}
