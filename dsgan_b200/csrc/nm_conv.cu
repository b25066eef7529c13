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
  int tiles_x, tiles_y, total_tiles, nbuf;
  int act, dact, accumulate, ld_aux;
  const bf16* in; const bf16* w; bf16* out; bf16* pre; const float* bias; const bf16* aux;
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

// CP: channels per pixel in the staged tile (8 / 16: narrow inputs padded; 32 / 64: wide inputs of the narrow-OUTPUT layers);
// NTN: n8 tiles (Co padded to 8*NTN)
template <int CP, int NTN>
__global__ void __launch_bounds__(THREADS, 2) k_nm_conv(const NmParams p) {
  constexpr int CPB = CP >= 32 ? CP * 2 + 16 : CP * 2;   // pixel pitch: 8 consecutive pixels must hit 8 different 16-byte bank groups
  constexpr int CON = NTN * 8, SPITCH = CON * 2 + 16, C8 = CP / 8, SPT = CP >= 16 ? CP / 16 : 1;
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nchunks = CP == 8 ? ((p.ntaps + 1) & ~1) : C8 * p.ntaps;  // 16-byte k chunks per weight row (even)
  const int wpitch = ((nchunks & 1) ? nchunks : nchunks + 1) * 16;    // odd number of 16-byte units: conflict-free ldmatrix rows
  unsigned char* wsm = dsm;
  int* toff = reinterpret_cast<int*>(dsm + CON * wpitch);
  float* sbias = reinterpret_cast<float*>(toff + 16);                  // [CON] (0 beyond Co / without bias)
  unsigned char* stage = reinterpret_cast<unsigned char*>(sbias + CON);
  unsigned char* tile_s = stage + 8 * 32 * SPITCH;                    // two buffers of tile_bytes
  const int tile_bytes = (p.rows * p.cols * CPB + 127) & ~127;

  // weights -> shared [co][k chunk]; chunk = tap (CP 8) or (tap, channel half) (CP 16); zero beyond the taps
  for (int i = tid; i < CON * nchunks; i += THREADS) {
    const int n = i / nchunks, kc = i % nchunks;
    const int tap = kc / C8, half = kc % C8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tap < p.ntaps && n < p.co_pad)
      v = __ldg(reinterpret_cast<const uint4*>(p.w + ((size_t)p.slab[tap] * p.co_pad + n) * p.ci_pad + half * 8));
    *reinterpret_cast<uint4*>(wsm + n * wpitch + kc * 16) = v;
  }
  if (tid >= 32 && tid < 32 + CON) sbias[tid - 32] = (p.bias && tid - 32 < p.Co) ? __ldg(p.bias + tid - 32) : 0.f;
  if (tid < 16) {
    const int tp = tid < p.ntaps ? tid : p.ntaps - 1;   // a dummy tap (zero weights) must still point inside the tile
    toff[tid] = ((p.dy[tp] - p.dy_min) * p.cols + (p.dx[tp] - p.dx_min)) * CPB;
  }
  // Pad lanes of a narrow tensor (channels Ci..CP-1) are never trusted to be finite: they are cleared in the A fragments
  // (register e of an ldmatrix.x4 holds channels 2t, 2t+1 of its 8-channel chunk), so the tile itself can be staged by
  // cp.async straight into shared memory, one tile ahead of the MMAs.
  auto cmask = [&](int c) { return (c < p.Ci ? 0x0000ffffu : 0u) | (c + 1 < p.Ci ? 0xffff0000u : 0u); };
  const uint32_t amask_lo = CP >= 32 ? 0xffffffffu : cmask(2 * t);   // (wide inputs: Ci == CP, nothing to clear)
  const uint32_t amask_hi = CP >= 32 ? 0xffffffffu : (CP == 8 ? amask_lo : cmask(8 + 2 * t));
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
    const int rowchunks = p.cols * C8;   // 16-byte chunks per staged row; a warp walks one row (no runtime divisions)
    for (int pyl = warp; pyl < p.rows; pyl += THREADS / 32) {
      const int iy = iy0 + pyl;
      const bool rok = iy >= 0 && iy < p.Hi;
      const bf16* rsrc = p.in + (((size_t)img * p.Hi + (rok ? iy : 0)) * p.Wi) * p.ld_in;
      const uint32_t rdst = dst + pyl * p.cols * CPB;
      for (int c = lane; c < rowchunks; c += 32) {
        const int h = c % C8, pxl = c / C8, ix = ix0 + pxl;
        const bool ok = rok && ix >= 0 && ix < p.Wi && h * 8 < p.ld_in;
        cp_async16_zfill(rdst + pxl * CPB + h * 16, ok ? rsrc + (size_t)ix * p.ld_in + h * 8 : p.in, ok);
      }
    }
    cp_async_commit();
  };

  if ((int)blockIdx.x < p.total_tiles && p.nbuf == 2) prefetch(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int img = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int gy0 = ty * TH, gx0 = tx * TW;
    int buf = 0;
    if (p.nbuf == 2) {
      buf = it & 1;
      cp_async_wait_all();
      __syncthreads();   // this tile has landed for everyone; the other buffer (tile it-1) is consumed; weights visible
      if (tile + (int)gridDim.x < p.total_tiles) prefetch(tile + gridDim.x, buf ^ 1);
    } else {             // (64-channel tiles: one buffer, so that two CTAs still fit an SM)
      __syncthreads();
      prefetch(tile, 0);
      cp_async_wait_all();
      __syncthreads();
    }

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
      const int off = CP == 8 ? toff[2 * ks + sel] : toff[ks / SPT] + (ks % SPT) * 32 + sel * 16;
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
    // epilogue: v = acc + bias (+ out) ; v *= dact'(aux) ; pre = v ; out = act(v), through the warp's staging rows
    const int gy = gy0 + warp;
    // Input-gradient calls (accumulate / dact): the sums are staged un-activated and the rest of the contract is applied on the
    // 16-byte vectors of the store loop (old output and aux read coalesced).  pass 0 = pre-activation copy, 1 = activated output,
    // 2 = staged raw for the input-gradient form.
    const bool igrad = p.accumulate || p.dact;
#pragma unroll 1
    for (int pass = igrad ? 2 : (p.pre ? 0 : 1); pass < 3; ++pass) {
      if (pass == 2 && !igrad) break;
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float v[NTN * 4];
#pragma unroll
        for (int j = 0; j < NTN; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(sbias + 8 * j + 2 * t);
          v[j * 4 + 0] = acc[mt][j][0] + bb.x; v[j * 4 + 1] = acc[mt][j][1] + bb.y;
          v[j * 4 + 2] = acc[mt][j][2] + bb.x; v[j * 4 + 3] = acc[mt][j][3] + bb.y;
        }
        if (pass == 1) act_fwd_fast_vec<NTN * 4>(p.act, v);   // one uniform test of `act`, then a straight loop
#pragma unroll
        for (int j = 0; j < NTN; ++j) {
          const int ch = 8 * j + 2 * t;
          *reinterpret_cast<uint32_t*>(my_stage + (mt * 16 + g) * SPITCH + ch * 2) = pack_bf2(v[j * 4], v[j * 4 + 1]);
          *reinterpret_cast<uint32_t*>(my_stage + (mt * 16 + g + 8) * SPITCH + ch * 2) = pack_bf2(v[j * 4 + 2], v[j * 4 + 3]);
        }
      }
      __syncwarp();
      if (gy < p.Hg) {
        const size_t rowbase = ((size_t)img * p.Ho + gy + p.oy0) * p.Wo + p.ox0;
        if (pass < 2) {
          bf16* dst = pass == 0 ? p.pre : p.out;
          const int ld = pass == 0 ? p.ld_pre : p.ldc;
          for (int i = lane; i < 32 * NTN; i += 32) {
            const int px = i / NTN, c16 = i % NTN, gx = gx0 + px;
            if (gx < p.Wg && c16 < co8)
              *reinterpret_cast<uint4*>(dst + (rowbase + gx) * ld + c16 * 8) =
                  *reinterpret_cast<const uint4*>(my_stage + px * SPITCH + c16 * 16);
          }
        } else {
#pragma unroll 1
          for (int i = lane; i < 32 * NTN; i += 32) {
            const int px = i / NTN, c16 = i % NTN, gx = gx0 + px;
            if (gx >= p.Wg || c16 >= co8) continue;
            const uint4 sv = *reinterpret_cast<const uint4*>(my_stage + px * SPITCH + c16 * 16);
            const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
            float f[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(sw[e] << 16); f[2 * e + 1] = __uint_as_float(sw[e] & 0xffff0000u); }
            bf16* op = p.out + (rowbase + gx) * p.ldc + c16 * 8;
            if (p.accumulate) {
              const uint4 ov = *reinterpret_cast<const uint4*>(op);
              const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { f[2 * e] += __uint_as_float(ow[e] << 16); f[2 * e + 1] += __uint_as_float(ow[e] & 0xffff0000u); }
            }
            if (p.dact) {
              const uint4 av = *reinterpret_cast<const uint4*>(p.aux + (rowbase + gx) * p.ld_aux + c16 * 8);
              const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
              float a[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) { a[2 * e] = __uint_as_float(aw[e] << 16); a[2 * e + 1] = __uint_as_float(aw[e] & 0xffff0000u); }
              act_bwd_fast_mul<8>(p.dact, f, a);
            }
            if (p.pre)
              *reinterpret_cast<uint4*>(p.pre + (rowbase + gx) * p.ld_pre + c16 * 8) =
                  make_uint4(pack_bf2(f[0], f[1]), pack_bf2(f[2], f[3]), pack_bf2(f[4], f[5]), pack_bf2(f[6], f[7]));
            act_fwd_fast_vec<8>(p.act, f);
            *reinterpret_cast<uint4*>(op) = make_uint4(pack_bf2(f[0], f[1]), pack_bf2(f[2], f[3]), pack_bf2(f[4], f[5]), pack_bf2(f[6], f[7]));
          }
        }
      }
      if (pass == 2) break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Weight gradient of the narrow layers:  dW[tap][m][n] += sum_px G[px][m] * X[px*xs + tap][n]  with one side of 1/3/6/12
// channels.  M = the wide side's channels (m16 tiles), N = the narrow side padded to 8/16, K = grid positions: both operands
// are NHWC tiles whose rows are pixels, read through ldmatrix.trans; the tap shift is a per-lane row address on the X tile.
// CTA = 8 x 32 grid positions; warp = (m16 tile of the wide side, pixel group); accumulators persist over all tiles of the
// CTA and are reduced once (shared, then one atomicAdd per weight and CTA).
struct NmWgParams {
  int N, Hg, Wg, Cg, ld_g, Hx, Wx, Cx, ld_x, xs, ntaps;
  int dy[16], dx[16];
  long long tap_off[16];
  long long s_g, s_x;
  int wide_is_g, mt_total, cgp, cxp, gpitch, xpitch;
  int dy_min, dx_min, xrows, xcols;
  int tiles_x, tiles_y, total_tiles, nbuf, gbytes, xbytes;
  const bf16* G; const bf16* X; float* dW;
};

__device__ __forceinline__ void ldsm_x4_t(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t a, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}

template <int NTAPS, int NTn>
__global__ void __launch_bounds__(THREADS, 2) k_nm_wgrad(const NmWgParams p) {
  constexpr int CNP = 8 * NTn;
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int cwp = 16 * p.mt_total;
  float* sdw = reinterpret_cast<float*>(dsm);                        // [NTAPS][cwp][CNP]
  int* toff = reinterpret_cast<int*>(sdw + NTAPS * cwp * CNP);
  unsigned char* gbuf = reinterpret_cast<unsigned char*>(toff + 16);
  unsigned char* xbuf = gbuf + p.nbuf * p.gbytes;
  for (int i = tid; i < NTAPS * cwp * CNP; i += THREADS) sdw[i] = 0.f;
  if (tid < NTAPS) toff[tid] = ((p.dy[tid] - p.dy_min) * p.xcols + (p.dx[tid] - p.dx_min)) * p.xpitch;
  const uint32_t g_u = (uint32_t)__cvta_generic_to_shared(gbuf), x_u = (uint32_t)__cvta_generic_to_shared(xbuf);

  auto prefetch = [&](int tile, int buf) {
    const int img = tile / (p.tiles_x * p.tiles_y), ty = (tile / p.tiles_x) % p.tiles_y, tx = tile % p.tiles_x;
    const int gy0 = ty * TH, gx0 = tx * TW;
    const int cg8 = p.cgp / 8, cx8 = p.cxp / 8;   // powers of two (1..8)
    const int gsh = __ffs(cg8) - 1, xsh = __ffs(cx8) - 1;
    // a warp walks one staged row, a lane its 16-byte chunks: no divisions by run-time extents in the copy loops
    for (int py = warp; py < TH; py += THREADS / 32) {
      const int gy = gy0 + py;
      const bool rok = gy < p.Hg;
      const bf16* rsrc = p.G + (((size_t)img * p.Hg + (rok ? gy : 0)) * p.Wg) * p.ld_g;
      const uint32_t rdst = g_u + buf * p.gbytes + py * TW * p.gpitch;
      for (int c = lane; c < TW * cg8; c += 32) {
        const int px = c >> gsh, h = c & (cg8 - 1), gx = gx0 + px;
        const bool ok = rok && gx < p.Wg && h * 8 < p.ld_g;
        cp_async16_zfill(rdst + px * p.gpitch + h * 16, ok ? rsrc + (size_t)gx * p.ld_g + h * 8 : p.G, ok);
      }
    }
    const int iy0 = gy0 * p.xs + p.dy_min, ix0 = gx0 * p.xs + p.dx_min;
    for (int pyl = warp; pyl < p.xrows; pyl += THREADS / 32) {
      const int iy = iy0 + pyl;
      const bool rok = iy >= 0 && iy < p.Hx;
      const bf16* rsrc = p.X + (((size_t)img * p.Hx + (rok ? iy : 0)) * p.Wx) * p.ld_x;
      const uint32_t rdst = x_u + buf * p.xbytes + pyl * p.xcols * p.xpitch;
      for (int c = lane; c < p.xcols * cx8; c += 32) {
        const int pxl = c >> xsh, h = c & (cx8 - 1), ix = ix0 + pxl;
        const bool ok = rok && ix >= 0 && ix < p.Wx && h * 8 < p.ld_x;
        cp_async16_zfill(rdst + pxl * p.xpitch + h * 16, ok ? rsrc + (size_t)ix * p.ld_x + h * 8 : p.X, ok);
      }
    }
    cp_async_commit();
  };

  float acc[NTAPS][NTn][4];
#pragma unroll
  for (int tp = 0; tp < NTAPS; ++tp)
#pragma unroll
    for (int nt = 0; nt < NTn; ++nt) acc[tp][nt][0] = acc[tp][nt][1] = acc[tp][nt][2] = acc[tp][nt][3] = 0.f;

  const int mt = warp % p.mt_total, pg = warp / p.mt_total, npg = 8 / p.mt_total;
  // A (wide side): matrices (k 0-7 | k 8-15) x (m 0-7 | m 8-15);  B (narrow side): k 0-7, k 8-15 for lanes 0..15
  const int a_px = (lane & 7) + (lane >> 4) * 8, a_ch = ((lane >> 3) & 1) * 16 + mt * 32;
  const int b_px = lane & 15;

  if ((int)blockIdx.x < p.total_tiles && p.nbuf == 2) prefetch(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    int buf = 0;
    if (p.nbuf == 2) {
      buf = it & 1;
      cp_async_wait_all();
      __syncthreads();
      if (tile + (int)gridDim.x < p.total_tiles) prefetch(tile + gridDim.x, buf ^ 1);
    } else {
      __syncthreads();           // previous tile consumed
      prefetch(tile, 0);
      cp_async_wait_all();
      __syncthreads();
    }
    const uint32_t gt = g_u + buf * p.gbytes, xt = x_u + buf * p.xbytes;
#pragma unroll 1
    for (int ks = pg; ks < 16; ks += npg) {
      const int r = ks >> 1, c0 = (ks & 1) * 16;
      if (p.wide_is_g) {
        uint32_t a[4];
        ldsm_x4_t(gt + (r * TW + c0 + a_px) * p.gpitch + a_ch, a[0], a[1], a[2], a[3]);
        const uint32_t xb = xt + ((r * p.xs) * p.xcols + (c0 + b_px) * p.xs) * p.xpitch;
#pragma unroll
        for (int tp = 0; tp < NTAPS; ++tp) {
#pragma unroll
          for (int nt = 0; nt < NTn; ++nt) {
            uint32_t b0, b1;
            ldsm_x2_t(xb + toff[tp] + nt * 16, b0, b1);
            mma16816(acc[tp][nt], a, b0, b1);
          }
        }
      } else {
        uint32_t b[NTn][2];
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt) ldsm_x2_t(gt + (r * TW + c0 + b_px) * p.gpitch + nt * 16, b[nt][0], b[nt][1]);
        const uint32_t xa = xt + ((r * p.xs) * p.xcols + (c0 + a_px) * p.xs) * p.xpitch + a_ch;
#pragma unroll
        for (int tp = 0; tp < NTAPS; ++tp) {
          uint32_t a[4];
          ldsm_x4_t(xa + toff[tp], a[0], a[1], a[2], a[3]);
#pragma unroll
          for (int nt = 0; nt < NTn; ++nt) mma16816(acc[tp][nt], a, b[nt][0], b[nt][1]);
        }
      }
    }
  }
  // reduce: fragment element e of thread (g, t) is (m = mt*16 + g + 8*(e>>1), n = nt*8 + 2t + (e&1))
  const int cw = p.wide_is_g ? p.Cg : p.Cx, cn = p.wide_is_g ? p.Cx : p.Cg;
#pragma unroll
  for (int tp = 0; tp < NTAPS; ++tp)
#pragma unroll
    for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + g + 8 * (e >> 1), n = nt * 8 + 2 * t + (e & 1);
        if (m < cw && n < cn) atomicAdd(sdw + (tp * cwp + m) * CNP + n, acc[tp][nt][e]);
      }
  __syncthreads();
  for (int i = tid; i < NTAPS * cwp * CNP; i += THREADS) {
    const int n = i % CNP, m = (i / CNP) % cwp, tp = i / (CNP * cwp);
    if (m >= cw || n >= cn) continue;
    const long long off = p.tap_off[tp] + (p.wide_is_g ? m * p.s_g + n * p.s_x : n * p.s_g + m * p.s_x);
    atomicAdd(p.dW + off, sdw[i]);
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

template <int NTAPS, int NTn>
int launch_wg(const NmWgParams& p, size_t smem, cudaStream_t s) {
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(k_nm_wgrad<NTAPS, NTn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  int grid = (smem > 110 * 1024 ? 1 : 2) * sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  const int per = (p.total_tiles + grid - 1) / grid;
  grid = (p.total_tiles + per - 1) / per;
  k_nm_wgrad<NTAPS, NTn><<<grid, THREADS, smem, s>>>(p);
  return DS_LAUNCHED("nm_wgrad");
}
}  // namespace

bool conv_try(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out, void* pre_out,
              const void* aux, void* stream, int* rc) {
  {
    const char* e = getenv("DSGAN_NM_CONV");
    if (e && e[0] == '0') return false;
  }
  const bool narrow_in = (d->Ci == 1 || d->Ci == 3 || d->Ci == 6 || d->Ci == 12) && d->Co <= 64 && d->Co >= 8;
  const bool narrow_out = !narrow_in && d->Co <= 16 && (d->Ci == 16 || d->Ci == 32 || d->Ci == 64);
  if (!narrow_in && !narrow_out) return false;
  if (d->nclass != 1 || d->out_stride != 1) return false;
  if (d->ntaps[0] < 1 || d->ntaps[0] > 16 || (d->in_stride != 1 && d->in_stride != 2)) return false;
  const int CP = narrow_in ? (d->Ci <= 8 ? 8 : 16) : d->Ci;
  const int co8 = (d->Co + 7) / 8 * 8;
  if (d->ld_in % 8 || d->ld_in < (d->Ci + 7) / 8 * 8 || (uintptr_t)in % 16 || (uintptr_t)w_slabs % 16 || d->ci_pad % 8) return false;
  auto out_ok = [&](const void* ptr, int ld) {
    if (!ptr) return true;
    return (uintptr_t)ptr % 16 == 0 && ld % 8 == 0 && ld >= co8;
  };
  if (!out_ok(out, d->ld_out) || !out_ok(pre_out, d->ld_pre) || !out_ok(aux, d->ld_aux)) return false;
  if (d->dact && !aux) return false;
  if (d->Co % 8 && (d->ld_out != co8 || (pre_out && d->ld_pre != co8) || (aux && d->ld_aux != co8))) return false;   // ragged Co: whole-pitch tensors only
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
  p.act = d->act; p.dact = aux ? d->dact : 0; p.accumulate = d->accumulate; p.ld_aux = d->ld_aux;
  p.in = (const bf16*)in; p.w = (const bf16*)w_slabs; p.out = (bf16*)out; p.pre = (bf16*)pre_out; p.bias = bias;
  p.aux = (const bf16*)aux;
  const int NTN = co8 <= 16 ? 2 : (co8 <= 32 ? 4 : 8);
  const int CON = NTN * 8;
  const int nchunks = CP == 8 ? ((p.ntaps + 1) & ~1) : (CP / 8) * p.ntaps;
  const int wpitch = ((nchunks & 1) ? nchunks : nchunks + 1) * 16;
  const int ppb = CP >= 32 ? CP * 2 + 16 : CP * 2;
  const size_t tile_bytes = (size_t)((p.rows * p.cols * ppb + 127) & ~127);
  const size_t fixed = (size_t)CON * wpitch + 64 + (size_t)CON * 4 + (size_t)8 * 32 * (CON * 2 + 16);
  p.nbuf = fixed + 2 * tile_bytes <= 100 * 1024 ? 2 : 1;
  const size_t smem = fixed + p.nbuf * tile_bytes;
  if (smem > 100 * 1024) return false;
  cudaStream_t s = (cudaStream_t)stream;
#define NM_CASE(CPV, NT) if (CP == CPV && NTN == NT) { *rc = launch<CPV, NT>(p, smem, s); return true; }
  if (getenv("DSGAN_NM_TRACE")) fprintf(stderr, "nm_conv: Ci=%d Co=%d taps=%d CP=%d NTN=%d nbuf=%d smem=%zu\n", d->Ci, d->Co, p.ntaps, CP, NTN, p.nbuf, smem);
  NM_CASE(8, 2) NM_CASE(8, 4) NM_CASE(8, 8) NM_CASE(16, 2) NM_CASE(16, 4) NM_CASE(16, 8) NM_CASE(32, 2) NM_CASE(64, 2)
#undef NM_CASE
  return false;
}

bool wgrad_try(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream, int* rc) {
  {
    const char* e = getenv("DSGAN_NM_CONV");
    if (e && e[0] == '0') return false;
  }
  auto narrow = [](int c) { return c == 1 || c == 3 || c == 6 || c == 12; };
  const bool ng = narrow(d->Cg), nx = narrow(d->Cx);
  if (!ng && !nx) return false;
  const bool wide_is_g = nx && (!ng || d->Cg >= d->Cx);
  const int cw = wide_is_g ? d->Cg : d->Cx, cn = wide_is_g ? d->Cx : d->Cg;
  if (cw > 64 || (d->ntaps != 1 && d->ntaps != 9 && d->ntaps != 16) || (d->x_stride != 1 && d->x_stride != 2)) return false;
  if (d->ld_g % 8 || d->ld_x % 8 || (uintptr_t)G % 16 || (uintptr_t)X % 16) return false;
  NmWgParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hg = d->Hg; p.Wg = d->Wg; p.Cg = d->Cg; p.ld_g = d->ld_g; p.Hx = d->Hx; p.Wx = d->Wx; p.Cx = d->Cx;
  p.ld_x = d->ld_x; p.xs = d->x_stride; p.ntaps = d->ntaps;
  int dy0 = 1 << 30, dy1 = -(1 << 30), dx0 = 1 << 30, dx1 = -(1 << 30);
  for (int t = 0; t < d->ntaps; ++t) {
    p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.tap_off[t] = d->tap_off[t];
    dy0 = d->dy[t] < dy0 ? d->dy[t] : dy0; dy1 = d->dy[t] > dy1 ? d->dy[t] : dy1;
    dx0 = d->dx[t] < dx0 ? d->dx[t] : dx0; dx1 = d->dx[t] > dx1 ? d->dx[t] : dx1;
  }
  p.s_g = d->s_g; p.s_x = d->s_x; p.wide_is_g = wide_is_g ? 1 : 0;
  const int mt = (cw + 15) / 16;
  p.mt_total = mt == 3 ? 4 : mt;
  const int NTn = cn > 8 ? 2 : 1;
  const int cwp = 16 * p.mt_total, cnp = 8 * NTn;
  p.cgp = wide_is_g ? cwp : cnp; p.cxp = wide_is_g ? cnp : cwp;
  auto pitch = [](int cp) { const int b = cp * 2; return (b / 16) % 2 == 0 ? b + 16 : b; };   // odd number of 16-byte units
  p.gpitch = pitch(p.cgp); p.xpitch = pitch(p.cxp);
  p.dy_min = dy0; p.dx_min = dx0;
  p.xrows = (TH - 1) * p.xs + (dy1 - dy0) + 1;
  p.xcols = (TW - 1) * p.xs + (dx1 - dx0) + 1;
  p.tiles_x = (d->Wg + TW - 1) / TW; p.tiles_y = (d->Hg + TH - 1) / TH;
  const long long total = (long long)d->N * p.tiles_x * p.tiles_y;
  if (total >= (1LL << 31)) return false;
  p.total_tiles = (int)total;
  p.gbytes = (TH * TW * p.gpitch + 127) & ~127;
  p.xbytes = (p.xrows * p.xcols * p.xpitch + 127) & ~127;
  const size_t fixed = (size_t)d->ntaps * cwp * cnp * 4 + 64;
  p.nbuf = (fixed + 2 * (size_t)(p.gbytes + p.xbytes) <= 110 * 1024) ? 2 : 1;
  const size_t smem = fixed + (size_t)p.nbuf * (p.gbytes + p.xbytes);
  if (smem > 200 * 1024) return false;
  p.G = (const bf16*)G; p.X = (const bf16*)X; p.dW = dW;
  cudaStream_t s = (cudaStream_t)stream;
#define NMW_CASE(NT, NN) if (d->ntaps == NT && NTn == NN) { *rc = launch_wg<NT, NN>(p, smem, s); return true; }
  NMW_CASE(1, 1) NMW_CASE(1, 2) NMW_CASE(9, 1) NMW_CASE(9, 2) NMW_CASE(16, 1)   // (16 taps x 2 narrow tiles would need 128 accumulator registers: not built)
#undef NMW_CASE
  return false;
}

}  // namespace nm
}  // namespace dsgan
