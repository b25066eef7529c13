// tcgen05 / TMEM / TMA GEMM family for the pointwise layers (nn.Linear and 1x1 nn.Conv2d: MixConvNeXtML.py:218,222-224,
// 335-425 ...), bf16 operands, fp32 accumulation in tensor memory.
//
//   FWD   : C[M,N]  = epi( A[M,K] . B[N,K]^T )        A K-major (activations), B K-major (weight [out,in])
//   DGRAD : C[M,N]  = epi( A[M,K] . B[K,N]   )        A K-major (dY),          B MN-major (same weight, read as [k][n])
//   WGRAD : C[M,N] += A[K,M]^T . B[K,N]   (fp32)      A MN-major (dY as [k][m]), B MN-major (X as [k][n]); split over K
//
// One persistent CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer (single elected thread) + TMEM owner,
// warps 2..5 = epilogue (one TMEM lane quarter each).  4-stage smem ring (full/empty mbarriers) and a double-buffered
// TMEM accumulator (tmem_full/tmem_empty) so the epilogue of tile i overlaps the main loop of tile i+1.
#include "tc_common.cuh"
#include "../../include/dsgan_b200.h"
#include <map>
#include <mutex>
#include <tuple>
#include <string.h>

using namespace dsgan;
using namespace dsgan::tc;

namespace dsgan {
namespace tc {

// ---- host: driver entry point + tensor-map cache ------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return 1;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,..] stride0 %llu box [%u,%u,..]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return 1;
  }
  return 0;
}
// ---- host side --------------------------------------------------------------------------------------------------
struct MapKey {
  const void* base; uint64_t d0, d1, s0; uint32_t b0, b1;
  bool operator<(const MapKey& o) const {
    return std::tie(base, d0, d1, s0, b0, b1) < std::tie(o.base, o.d0, o.d1, o.s0, o.b0, o.b1);
  }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

// 2-D bf16 tensor map: dims {inner, outer}, row pitch in elements, box {b0, b1}
int get_map_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t b0,
               uint32_t b1) {
  MapKey k{base, inner, outer, pitch_elems * 2, b0, b1};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  uint64_t dims[2] = {inner, outer}, strides[1] = {pitch_elems * 2};
  uint32_t box[2] = {b0, b1};
  if (encode_tmap_bf16(out, base, 2, dims, strides, box, nullptr)) return 1;
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[k] = *out;
  return 0;
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

}  // namespace tc
}  // namespace dsgan

namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 16;              // four per TMEM lane quarter, each takes a share of the 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;

struct GemmParams {
  int M, N, K;            // problem (K = contraction length)
  int m_tiles, n_tiles, splits, kb_per_split;
  // epilogue
  void* C; int ldc;       // bf16 (FWD/DGRAD) or fp32 (WGRAD) output
  const float* bias;
  void* pre; int ld_pre;  // optional bf16 pre-activation copy
  const void* aux; int ld_aux;
  int act, dact, accumulate;
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // the ring is latency-bound (TMA round trip ~1-2 us): keep ~190 KB in flight -> 8 / 6 / 4 stages for BN = 64 / 128 / 256
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 512;       // fp32 bias vector (N <= MAX_SMEM_BIAS), read by every tile's epilogue
  static constexpr int TOTAL = BIAS_OFF + 4 * 4096 + 1024;  // barriers + bias + slack for 1024-B alignment
};
constexpr int MAX_SMEM_BIAS = 4096;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// MODE 0: A K-major, B K-major (FWD); 1: A K-major, B MN-major (DGRAD); 2: both MN-major, fp32 atomic output (WGRAD)
template <int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr bool A_MN = (MODE == 2), B_MN = (MODE >= 1);
  using SL = SmemLayout<BN>;
  // Epilogue teams (see tc_conv.cu): a 128 x BN accumulator is 4*BN/32 warp tasks; with BN = 64 the 16 epilogue warps split
  // into two teams that drain different tiles concurrently, the TMEM ring holds two accumulators per team.
  constexpr int CHUNKS = BN / 32;
  constexpr int TEAMS = CHUNKS >= 4 ? 1 : 4 / CHUNKS;
  constexpr int TEAM_WARPS = EPI_WARPS / TEAMS;
  constexpr int NACC = 2 * TEAMS;
  constexpr int TMEM_COLS = NACC * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty = full + SL::STAGES;
  uint64_t* tfull = empty + SL::STAGES;
  uint64_t* tempty = tfull + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SL::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TEAM_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  // the bias is the same for every tile with the same column block: keep it in shared memory (a global load per 32-column
  // chunk sat on the epilogue's critical path: 11 % of uc4.pwconv1)
  float* sbias = reinterpret_cast<float*>(smem + SL::BIAS_OFF);
  const bool bias_smem = MODE != 2 && p.bias != nullptr && p.N <= MAX_SMEM_BIAS;
  if (bias_smem)
    for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) sbias[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles, split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, (p.K + BK - 1) / BK);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          mbar_expect_tx(&full[stage], SL::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, &full[stage], m_blk * BM + j * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, &full[stage], n_blk * BN + j * 64, kb * BK);
          } else {
            tma_load_2d(sb, &tmB, &full[stage], kb * BK, n_blk * BN);
          }
          if (++stage == SL::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_bf16(BM, BN, A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_SBO = 1024, A_KSTEP = A_MN ? 16 * 128 : 32;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_SBO = 1024, B_KSTEP = B_MN ? 16 * 128 : 32;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, (p.K + BK - 1) / BK);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint32_t sb = sa + SL::A_BYTES;
          const uint64_t adesc = smem_desc_sw128(sa, A_LBO, A_SBO);
          const uint64_t bdesc = smem_desc_sw128(sb, B_LBO, B_SBO);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(d_tmem, adesc + (uint64_t)((k * A_KSTEP) >> 4), bdesc + (uint64_t)((k * B_KSTEP) >> 4), IDESC,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == SL::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);      // accumulator complete -> epilogue
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps (2..5): TMEM -> registers -> global =================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;
    const int team = CHUNKS >= 4 ? 0 : part / CHUNKS;
    const int c_first = CHUNKS >= 4 ? part : part % CHUNKS, c_step = CHUNKS >= 4 ? 4 : CHUNKS;
    int it = team;
    for (int tile = blockIdx.x + team * gridDim.x; tile < num_tiles; tile += TEAMS * gridDim.x, it += TEAMS) {
      const int acc = it % NACC;
      const uint32_t acc_phase = (uint32_t)(it / NACC) & 1u;
      const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      // 256-bit row accesses when every row segment is 32-byte aligned
      const bool wide = MODE != 2 && (p.ldc % 16 == 0) && ((uintptr_t)p.C % 32 == 0) &&
                        (!p.pre || (p.ld_pre % 16 == 0 && (uintptr_t)p.pre % 32 == 0)) &&
                        (!p.aux || (p.ld_aux % 16 == 0 && (uintptr_t)p.aux % 32 == 0));
#pragma unroll 1
      for (int c = c_first; c < CHUNKS; c += c_step) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < p.N) {
        if (MODE == 2) {
          float* o = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            // 8 vector reductions instead of 32 scalar atomics: the split-K epilogue is bound by L2 atomic transactions
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        } else {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>((bias_smem ? sbias : p.bias) + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = bp[q];
              f[q * 4] += b4.x; f[q * 4 + 1] += b4.y; f[q * 4 + 2] += b4.z; f[q * 4 + 3] += b4.w;
            }
          }
          bf16* o = reinterpret_cast<bf16*>(p.C) + (size_t)row * p.ldc + col0;
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
            const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + (size_t)row * p.ld_aux + col0);
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
            uint4* pp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.pre) + (size_t)row * p.ld_pre + col0);
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
        }
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge before the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int BN, int MODE>
int launch(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
  static bool attr = false;
  constexpr int smem = SmemLayout<BN>::TOTAL;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("tc_gemm smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  k_tc_gemm<BN, MODE><<<grid, NUM_THREADS, smem, s>>>(a, b, p);
  return DS_LAUNCHED("tc_gemm");
}

int pick_bn(int N) { return N >= 256 ? 256 : (N >= 128 ? 128 : 64); }
}  // namespace

extern "C" {
int dsgan_tc_gemm_supported(int mode, long long M, int N, int K, int lda, int ldb, int ldc) {
  if (M < 1 || N < 64 || K < 64) return 0;
  if (N % 8 || K % 8 || lda % 8 || ldb % 8 || ldc % 8) return 0;
  if (mode == 0 || mode == 1) return (K % 64 == 0) && (N % 32 == 0);
  if (mode == 2) return (M % 8 == 0);
  return 0;
}

// mode 0 (FWD):   C[M,N] = epi(A[M,K] . W[N,K]^T)   A: bf16 [M,lda], W: bf16 [N,ldb]
// mode 1 (DGRAD): C[M,N] = epi(A[M,K] . W[K,N])     A: bf16 [M,lda], W: bf16 [K,ldb]   (the forward weight [out=K, in=N])
int dsgan_tc_gemm(int mode, const void* A, int lda, const void* Wt, int ldb, long long M, int N, int K, void* C, int ldc,
                  const float* bias, void* pre, int ld_pre, const void* aux, int ld_aux, int act, int dact,
                  int accumulate, void* stream) {
  DS_REQUIRE(mode == 0 || mode == 1, "tc_gemm: mode must be 0 (fwd) or 1 (dgrad)");
  DS_REQUIRE(dsgan_tc_gemm_supported(mode, M, N, K, lda, ldb, ldc), "tc_gemm: unsupported shape M=%lld N=%d K=%d", M, N, K);
  DS_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Wt % 16 == 0) && ((uintptr_t)C % 16 == 0), "tc_gemm: unaligned pointer");
  DS_REQUIRE(!dact || aux, "tc_gemm: dact needs aux");
  DS_REQUIRE((!pre || (((uintptr_t)pre % 16 == 0) && ld_pre % 8 == 0)) && (!aux || (((uintptr_t)aux % 16 == 0) && ld_aux % 8 == 0)),
             "tc_gemm: pre/aux must be 16-byte aligned with pitches that are multiples of 8");
  const int BN = pick_bn(N);
  CUtensorMap ta, tb;
  if (get_map_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM)) return 1;
  if (mode == 0) {
    if (get_map_2d(&tb, Wt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)BN)) return 1;
  } else {
    if (get_map_2d(&tb, Wt, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64)) return 1;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N; p.K = K;
  p.m_tiles = (int)((M + BM - 1) / BM); p.n_tiles = (N + BN - 1) / BN; p.splits = 1; p.kb_per_split = (K + BK - 1) / BK;
  p.C = C; p.ldc = ldc; p.bias = bias; p.pre = pre; p.ld_pre = ld_pre; p.aux = aux; p.ld_aux = ld_aux;
  p.act = act; p.dact = dact; p.accumulate = accumulate;
  cudaStream_t s = (cudaStream_t)stream;
#define TC_CASE(BN_, MODE_) if (BN == BN_ && mode == MODE_) return launch<BN_, MODE_>(ta, tb, p, s);
  TC_CASE(256, 0) TC_CASE(128, 0) TC_CASE(64, 0) TC_CASE(256, 1) TC_CASE(128, 1) TC_CASE(64, 1)
#undef TC_CASE
  set_error("tc_gemm: no kernel for BN=%d mode=%d", BN, mode);
  return 1;
}

// WGRAD: dW[Co,Ci] (fp32, ld_dw) += dY[P,Co]^T . X[P,Ci]     dY: bf16 [P,ld_dy], X: bf16 [P,ld_x]
int dsgan_tc_wgrad(const void* dY, int ld_dy, const void* X, int ld_x, long long P, int Co, int Ci, float* dW, int ld_dw,
                   void* stream) {
  DS_REQUIRE(Co % 8 == 0 && Ci % 8 == 0 && ld_dy % 8 == 0 && ld_x % 8 == 0 && Ci >= 64 && Co >= 64,
             "tc_wgrad: unsupported shape Co=%d Ci=%d", Co, Ci);
  DS_REQUIRE(((uintptr_t)dY % 16 == 0) && ((uintptr_t)X % 16 == 0), "tc_wgrad: unaligned pointer");
  const int BN = pick_bn(Ci);
  CUtensorMap ta, tb;
  if (get_map_2d(&ta, dY, (uint64_t)Co, (uint64_t)P, (uint64_t)ld_dy, 64, 64)) return 1;
  if (get_map_2d(&tb, X, (uint64_t)Ci, (uint64_t)P, (uint64_t)ld_x, 64, 64)) return 1;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = Co; p.N = Ci; p.K = (int)P;
  p.m_tiles = (Co + BM - 1) / BM; p.n_tiles = (Ci + BN - 1) / BN;
  const int kb_total = (int)((P + BK - 1) / BK);
  // Split the pixel range so that ONE wave of CTAs covers the machine (every extra split costs a full BM x BN fp32 reduction
  // into L2) and no split is shorter than 8 k-blocks: with the former 2 x SMs / tiles rule the small 1x1 layers (P = 16384,
  // one or two output tiles) ran 256 one-k-block CTAs whose atomic epilogues took 60-70 us for 0.3 GFLOP.
  int splits = num_sms() / (p.m_tiles * p.n_tiles);
  if (splits > kb_total / 8) splits = kb_total / 8;
  if (splits < 1) splits = 1;
  p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.C = dW; p.ldc = ld_dw;
  cudaStream_t s = (cudaStream_t)stream;
  if (BN == 256) return launch<256, 2>(ta, tb, p, s);
  if (BN == 128) return launch<128, 2>(ta, tb, p, s);
  return launch<64, 2>(ta, tb, p, s);
}

// fp32 -> bf16 copy of a flat buffer (packed GEMM operands of parameters)
__global__ void k_cast_bf16(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}
int dsgan_pack_bf16(const float* src, void* dst, long long n, void* stream) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_cast_bf16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  return DS_LAUNCHED("pack_bf16");
}
}
