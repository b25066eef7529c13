// Narrow-input convolutions (Ci in {1,3,6,12} -> Co <= 64) as a warp-level tensor-core implicit GEMM.
//
// Reference op sites: VGG16 conv1_1 (models/vgg.py:16, 3->64 k3), the PatchGAN first layer (networks.py:544, 6->32 k4 s2), the
// generator's input 1x1 / 3x3 convolutions on 3-channel images (MixConvNeXtML.py:335-338) and block c1's 12->64 pwconv.
// On the tcgen05 path these layers pad Ci to 64 (5-20 TF/s of useful work); on the CUDA cores (sc_conv.cu) they are FMA-issue
// bound (3->64 k3 at 16x256x256: 0.18 ms for 151 MB of traffic).  Here the K dimension is (tap, 8 or 16 padded channels):
// one 16-byte row of an ldmatrix tile is one input pixel of one tap, so the im2col operand is never materialised -- every
// lane just points ldmatrix at its pixel of the staged NHWC tile.  K = 72 for a 3x3 on 3 channels instead of 576.
//
// CTA = 8 x 32 output positions x all Co; warp w owns tile row w (two m16 tiles), acc[2][Co/8][4] in registers.
// Weights: bf16 slabs [slab][co_pad][ci_pad] (the tcgen05 packing) -> shared [co][tap*CP + ci] once per CTA.
// Output: fragments -> per-warp shared staging -> 16-byte NHWC stores (whole 128-byte lines for Co = 64).
// Same contract as dsgan_tc_conv for the supported subset (one parity class, unit output stride, no accumulate / dact).
#include "common.cuh"
#include "nm_conv.cuh"
#include <stdlib.h>
#include <string.h>

namespace dsgan {
namespace nm {
namespace {

constexpr int TH = 8, TW = 32, THREADS = 256;

struct NmParams {
  int N, Hg, Wg, Hi, Wi, Ho, Wo, Ci, Co, co_pad, ci_pad, ld_in, ldc, ld_pre;
  int is_, oy0, ox0, ntaps;
  int dy[16], dx[16], slab[16];
  int dy_min, dx_min, rows, cols;   // staged input region of a tile
  int tiles_x, tiles_y, total_tiles;
  int act;
  const bf16* in; const bf16* w; bf16* out; bf16* pre; const float* bias;
};

__device__ __forceinline__ void ldsm_x4(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem, const void* gmem, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// CP: padded channels per pixel in the staged tile (8 or 16); NTN: n8 tiles (Co padded to 8*NTN)
template <int CP, int NTN>
__global__ void __launch_bounds__(THREADS, 2) k_nm_conv(const NmParams p) {
  constexpr int CPB = CP * 2, CON = NTN * 8, SPITCH = CON * 2 + 16;
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nchunks = CP == 8 ? ((p.ntaps + 1) & ~1) : 2 * p.ntaps;   // 16-byte k chunks per weight row (even)
  const int wpitch = ((nchunks & 1) ? nchunks : nchunks + 1) * 16;    // odd number of 16-byte units: conflict-free ldmatrix rows
  unsigned char* wsm = dsm;
  int* toff = reinterpret_cast<int*>(dsm + CON * wpitch);
  unsigned char* stage = reinterpret_cast<unsigned char*>(toff + 16);
  unsigned char* tile_s = stage + 8 * 32 * SPITCH;                    // two buffers of tile_bytes
  const int tile_bytes = (p.rows * p.cols * CPB + 127) & ~127;

  // weights -> shared [co][k chunk]; chunk = tap (CP 8) or (tap, channel half) (CP 16); zero beyond the taps
  for (int i = tid; i < CON * nchunks; i += THREADS) {
    const int n = i / nchunks, kc = i % nchunks;
    const int tap = CP == 8 ? kc : kc >> 1, half = CP == 8 ? 0 : kc & 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tap < p.ntaps && n < p.co_pad)
      v = __ldg(reinterpret_cast<const uint4*>(p.w + ((size_t)p.slab[tap] * p.co_pad + n) * p.ci_pad + half * 8));
    *reinterpret_cast<uint4*>(wsm + n * wpitch + kc * 16) = v;
  }
  if (tid < 16) {
    const int tp = tid < p.ntaps ? tid : p.ntaps - 1;   // a dummy tap (zero weights) must still point inside the tile
    toff[tid] = ((p.dy[tp] - p.dy_min) * p.cols + (p.dx[tp] - p.dx_min)) * CPB;
  }
  // Pad lanes of a narrow tensor (channels Ci..CP-1) are never trusted to be finite: they are cleared in the A fragments
  // (register e of an ldmatrix.x4 holds channels 2t, 2t+1 of its 8-channel chunk), so the tile itself can be staged by
  // cp.async straight into shared memory, one tile ahead of the MMAs.
  auto cmask = [&](int c) { return (c < p.Ci ? 0x0000ffffu : 0u) | (c + 1 < p.Ci ? 0xffff0000u : 0u); };
  const uint32_t amask_lo = cmask(2 * t), amask_hi = CP == 8 ? amask_lo : cmask(8 + 2 * t);
  const uint32_t wsm_u = (uint32_t)__cvta_generic_to_shared(wsm), tile_u = (uint32_t)__cvta_generic_to_shared(tile_s);
  const int px_l = (lane & 7) + ((lane >> 3) & 1) * 8, sel = lane >> 4;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_kc = (lane >> 3) & 1;
  const int ksteps = nchunks / 2;
  unsigned char* my_stage = stage + warp * 32 * SPITCH;
  const int co8 = (p.Co + 7) / 8;   // 16-byte channel groups actually stored

  auto prefetch = [&](int tile, int buf) {
    const int img = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int iy0 = ty * TH * p.is_ + p.dy_min, ix0 = tx * TW * p.is_ + p.dx_min;
    const uint32_t dst = tile_u + buf * tile_bytes;
    for (int i = tid; i < p.rows * p.cols * (CP / 8); i += THREADS) {
      const int h = i % (CP / 8), pxl = (i / (CP / 8)) % p.cols, pyl = i / ((CP / 8) * p.cols);
      const int iy = iy0 + pyl, ix = ix0 + pxl;
      const bool ok = iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi && h * 8 < p.ld_in;
      const bf16* src = ok ? p.in + (((size_t)img * p.Hi + iy) * p.Wi + ix) * p.ld_in + h * 8 : p.in;
      cp_async16_zfill(dst + (pyl * p.cols + pxl) * CPB + h * 16, src, ok);
    }
    cp_async_commit();
  };

  if ((int)blockIdx.x < p.total_tiles) prefetch(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int img = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int gy0 = ty * TH, gx0 = tx * TW;
    const int buf = it & 1;
    cp_async_wait_all();
    __syncthreads();   // this tile has landed for everyone; the other buffer (tile it-1) is consumed; weights visible
    if (tile + (int)gridDim.x < p.total_tiles) prefetch(tile + gridDim.x, buf ^ 1);

    float acc[2][NTN][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < NTN; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
    uint32_t abase[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) abase[mt] = tile_u + buf * tile_bytes + ((warp * p.is_) * p.cols + (mt * 16 + px_l) * p.is_) * CPB;
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      const int off = CP == 8 ? toff[2 * ks + sel] : toff[ks] + sel * 16;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ldsm_x4(abase[mt] + off, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
        a[mt][0] &= amask_lo; a[mt][1] &= amask_lo; a[mt][2] &= amask_hi; a[mt][3] &= amask_hi;
      }
#pragma unroll
      for (int j2 = 0; j2 < NTN / 2; ++j2) {
        uint32_t b[4];
        ldsm_x4(wsm_u + (j2 * 16 + b_row) * wpitch + (2 * ks + b_kc) * 16, b[0], b[1], b[2], b[3]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * j2], a[mt], b[0], b[1]);
          mma16816(acc[mt][2 * j2 + 1], a[mt], b[2], b[3]);
        }
      }
    }
    // epilogue: v = acc + bias ; pre = v ; out = act(v), through the warp's staging rows
    const int gy = gy0 + warp;
#pragma unroll 1
    for (int pass = p.pre ? 0 : 1; pass < 2; ++pass) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NTN; ++j) {
        const int ch = 8 * j + 2 * t;
        const float b0 = (p.bias && ch < p.Co) ? __ldg(p.bias + ch) : 0.f, b1 = (p.bias && ch + 1 < p.Co) ? __ldg(p.bias + ch + 1) : 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float v[4] = {acc[mt][j][0] + b0, acc[mt][j][1] + b1, acc[mt][j][2] + b0, acc[mt][j][3] + b1};
          if (pass == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = act_fwd_fast(p.act, v[e]);
          }
          *reinterpret_cast<uint32_t*>(my_stage + (mt * 16 + g) * SPITCH + ch * 2) = pack_bf2(v[0], v[1]);
          *reinterpret_cast<uint32_t*>(my_stage + (mt * 16 + g + 8) * SPITCH + ch * 2) = pack_bf2(v[2], v[3]);
        }
      }
      __syncwarp();
      bf16* dst = pass == 0 ? p.pre : p.out;
      const int ld = pass == 0 ? p.ld_pre : p.ldc;
      if (gy < p.Hg) {
        const size_t rowbase = ((size_t)img * p.Ho + gy + p.oy0) * p.Wo + p.ox0;
        for (int i = lane; i < 32 * NTN; i += 32) {
          const int px = i / NTN, c16 = i % NTN, gx = gx0 + px;
          if (gx < p.Wg && c16 < co8)
            *reinterpret_cast<uint4*>(dst + (rowbase + gx) * ld + c16 * 8) =
                *reinterpret_cast<const uint4*>(my_stage + px * SPITCH + c16 * 16);
        }
      }
    }
  }
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

template <int CP, int NTN>
int launch(const NmParams& p, size_t smem, cudaStream_t s) {
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(k_nm_conv<CP, NTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  int grid = 2 * sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  const int per = (p.total_tiles + grid - 1) / grid;
  grid = (p.total_tiles + per - 1) / per;
  k_nm_conv<CP, NTN><<<grid, THREADS, smem, s>>>(p);
  return DS_LAUNCHED("nm_conv");
}
}  // namespace

bool conv_try(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out, void* pre_out,
              const void* aux, void* stream, int* rc) {
  {
    const char* e = getenv("DSGAN_NM_CONV");
    if (e && e[0] == '0') return false;
  }
  if (!(d->Ci == 1 || d->Ci == 3 || d->Ci == 6 || d->Ci == 12) || d->Co > 64 || d->Co < 8) return false;
  if (d->nclass != 1 || d->out_stride != 1 || d->accumulate || d->dact || aux) return false;
  if (d->ntaps[0] < 1 || d->ntaps[0] > 16 || (d->in_stride != 1 && d->in_stride != 2)) return false;
  const int CP = d->Ci <= 8 ? 8 : 16;
  const int co8 = (d->Co + 7) / 8 * 8;
  if (d->ld_in % 8 || d->ld_in < (d->Ci + 7) / 8 * 8 || (uintptr_t)in % 16 || (uintptr_t)w_slabs % 16 || d->ci_pad % 8) return false;
  auto out_ok = [&](const void* ptr, int ld) {
    if (!ptr) return true;
    return (uintptr_t)ptr % 16 == 0 && ld % 8 == 0 && ld >= co8;
  };
  if (!out_ok(out, d->ld_out) || !out_ok(pre_out, d->ld_pre)) return false;
  if (d->Co % 8 && (d->ld_out != co8 || (pre_out && d->ld_pre != co8))) return false;   // ragged Co: whole-pitch tensors only
  NmParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Hi = d->Hi; p.Wi = d->Wi; p.Ho = d->Ho; p.Wo = d->Wo;
  p.Ci = d->Ci; p.Co = d->Co; p.co_pad = d->co_pad; p.ci_pad = d->ci_pad;
  p.ld_in = d->ld_in; p.ldc = d->ld_out; p.ld_pre = d->ld_pre;
  p.is_ = d->in_stride; p.oy0 = d->oy0[0]; p.ox0 = d->ox0[0]; p.ntaps = d->ntaps[0];
  int dy0 = 1 << 30, dy1 = -(1 << 30), dx0 = 1 << 30, dx1 = -(1 << 30);
  for (int t = 0; t < p.ntaps; ++t) {
    p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.slab[t] = d->slab[t];
    dy0 = d->dy[t] < dy0 ? d->dy[t] : dy0; dy1 = d->dy[t] > dy1 ? d->dy[t] : dy1;
    dx0 = d->dx[t] < dx0 ? d->dx[t] : dx0; dx1 = d->dx[t] > dx1 ? d->dx[t] : dx1;
  }
  p.dy_min = dy0; p.dx_min = dx0;
  p.rows = (TH - 1) * p.is_ + (dy1 - dy0) + 1;
  p.cols = (TW - 1) * p.is_ + (dx1 - dx0) + 1;
  p.tiles_x = (d->Wg + TW - 1) / TW; p.tiles_y = (d->Hg + TH - 1) / TH;
  const long long total = (long long)d->N * p.tiles_x * p.tiles_y;
  if (total >= (1LL << 31)) return false;
  p.total_tiles = (int)total;
  p.act = d->act;
  p.in = (const bf16*)in; p.w = (const bf16*)w_slabs; p.out = (bf16*)out; p.pre = (bf16*)pre_out; p.bias = bias;
  const int NTN = co8 <= 16 ? 2 : (co8 <= 32 ? 4 : 8);
  const int CON = NTN * 8;
  const int nchunks = CP == 8 ? ((p.ntaps + 1) & ~1) : 2 * p.ntaps;
  const int wpitch = ((nchunks & 1) ? nchunks : nchunks + 1) * 16;
  const size_t smem = (size_t)CON * wpitch + 64 + (size_t)8 * 32 * (CON * 2 + 16) + 2 * (size_t)((p.rows * p.cols * CP * 2 + 127) & ~127);
  if (smem > 100 * 1024) return false;
  cudaStream_t s = (cudaStream_t)stream;
#define NM_CASE(CPV, NT) if (CP == CPV && NTN == NT) { *rc = launch<CPV, NT>(p, smem, s); return true; }
  NM_CASE(8, 2) NM_CASE(8, 4) NM_CASE(8, 8) NM_CASE(16, 2) NM_CASE(16, 4) NM_CASE(16, 8)
#undef NM_CASE
  return false;
}

}  // namespace nm
}  // namespace dsgan
