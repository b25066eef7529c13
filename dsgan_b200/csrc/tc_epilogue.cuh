// Shared epilogue of the tcgen05 kernels: TMEM -> registers -> (bias, residual, act', act) -> bf16 -> global.
//
// tcgen05.ld hands every lane ONE accumulator row, so a naive store makes each warp instruction touch 32 different
// cache lines (32 LSU wavefronts per 512 bytes).  Here a warp stages its 32-row x 64-column slab in a private, XOR-swizzled
// 4 KB shared-memory tile and then writes it out with lanes running along the row: one instruction = 4 rows x 128 B =
// 4 wavefronts.  The same staging turns the row-per-lane reads of `aux` (activation-derivative operand) and of the old
// output (accumulate) into coalesced loads.
#pragma once
#include "tc_common.cuh"

namespace dsgan {
namespace tc {

struct EpiArgs {
  bf16* C; int ldc;
  const float* bias;
  bf16* pre; int ld_pre;
  const bf16* aux; int ld_aux;
  int act, dact, accumulate;
  int ncols;  // valid output columns (N / Co)
};

__device__ __forceinline__ uint32_t epi_pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void epi_unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}

// coalesced global -> stage -> this lane's row (8 x 16 B)
__device__ __forceinline__ void epi_load_rows(const bf16* src, int ld, int col0, int ncol, int ncols, int my_row, bool my_ok,
                                              uint4* stage, int lane, uint4 (&rowv)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), s = lane & 7;
    const int rr = __shfl_sync(0xffffffffu, my_row, r);
    const bool ok = __shfl_sync(0xffffffffu, (int)my_ok, r) != 0;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (ok && s * 8 < ncol && col0 + s * 8 < ncols) u = __ldg(reinterpret_cast<const uint4*>(src + (size_t)rr * ld + col0 + s * 8));
    stage[r * 8 + (s ^ (r & 7))] = u;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) rowv[j] = stage[lane * 8 + (j ^ (lane & 7))];
  __syncwarp();
}
// this lane's row (8 x 16 B) -> stage -> coalesced global
__device__ __forceinline__ void epi_store_rows(bf16* dst, int ld, int col0, int ncol, int ncols, int my_row, bool my_ok,
                                               uint4* stage, int lane, const uint4 (&rowv)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) stage[lane * 8 + (j ^ (lane & 7))] = rowv[j];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), s = lane & 7;
    const int rr = __shfl_sync(0xffffffffu, my_row, r);
    const bool ok = __shfl_sync(0xffffffffu, (int)my_ok, r) != 0;
    if (ok && s * 8 < ncol && col0 + s * 8 < ncols)
      *reinterpret_cast<uint4*>(dst + (size_t)rr * ld + col0 + s * 8) = stage[r * 8 + (s ^ (r & 7))];
  }
  __syncwarp();
}

// One 64-column group (ncol = 32 or 64) of this warp's 32 accumulator rows.  `taddr` = TMEM address of the group's first
// column for this warp's lane quarter.  Requires 16-byte aligned C/pre/aux with pitches % 8 == 0 and ncols % 32 == 0.
__device__ __forceinline__ void epi_group(const EpiArgs& p, uint32_t taddr, int ncol, int col0, int my_row, bool my_ok,
                                          uint4* stage, int lane) {
  uint4 auxv[8], oldv[8], outv[8], prev[8];
  if (p.dact) epi_load_rows(p.aux, p.ld_aux, col0, ncol, p.ncols, my_row, my_ok, stage, lane, auxv);
  if (p.accumulate) epi_load_rows(p.C, p.ldc, col0, ncol, p.ncols, my_row, my_ok, stage, lane, oldv);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h * 32 < ncol) {  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(taddr + h * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const int c = col0 + h * 32;
      if (p.bias && c < p.ncols) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + c + j);
      }
      if (p.accumulate) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[8];
          epi_unpack8(oldv[h * 4 + q], o);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[q * 8 + e] += o[e];
        }
      }
      if (p.dact) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a[8];
          epi_unpack8(auxv[h * 4 + q], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[q * 8 + e] *= act_bwd_fast(p.dact, a[e]);
        }
      }
      if (p.pre) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          prev[h * 4 + q] = make_uint4(epi_pack2(f[q * 8], f[q * 8 + 1]), epi_pack2(f[q * 8 + 2], f[q * 8 + 3]),
                                       epi_pack2(f[q * 8 + 4], f[q * 8 + 5]), epi_pack2(f[q * 8 + 6], f[q * 8 + 7]));
      }
      if (p.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = act_fwd_fast(p.act, f[j]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        outv[h * 4 + q] = make_uint4(epi_pack2(f[q * 8], f[q * 8 + 1]), epi_pack2(f[q * 8 + 2], f[q * 8 + 3]),
                                     epi_pack2(f[q * 8 + 4], f[q * 8 + 5]), epi_pack2(f[q * 8 + 6], f[q * 8 + 7]));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) outv[h * 4 + q] = prev[h * 4 + q] = make_uint4(0, 0, 0, 0);
    }
  }
  if (p.pre) epi_store_rows(p.pre, p.ld_pre, col0, ncol, p.ncols, my_row, my_ok, stage, lane, prev);
  epi_store_rows(p.C, p.ldc, col0, ncol, p.ncols, my_row, my_ok, stage, lane, outv);
}

}  // namespace tc
}  // namespace dsgan
