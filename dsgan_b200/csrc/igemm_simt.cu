// CUDA-core implicit-GEMM convolution family: nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward,
// input-gradient and weight-gradient for ANY shape, in fp32 or bf16 activations with fp32 weights and
// fp32 accumulation.  This is the fp32 validation mode (north_star: "1e-4 in an fp32 validation mode") and
// the path for shapes the tcgen05 kernels do not take (K=3/6/12, N=1/3 ...).
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T, int TRANSPOSED>
__global__ void __launch_bounds__(NT) k_igemm(dsgan_conv_desc d, const T* __restrict__ in,
                                               const float* __restrict__ w, const float* __restrict__ bias,
                                               T* __restrict__ out, T* __restrict__ pre_out,
                                               const T* __restrict__ aux) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const long long M = (long long)d.N * d.Ho * d.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load mapping: 4 elements per thread: pixel am[i] = t/16 + 16*i, channel lane ak = t%16
  const int ak = t & 15;
  int a_n[4], a_y[4], a_x[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + (t >> 4) + 16 * i;
    a_ok[i] = m < M;
    long long mm = a_ok[i] ? m : 0;
    a_x[i] = (int)(mm % d.Wo);
    long long r = mm / d.Wo;
    a_y[i] = (int)(r % d.Ho);
    a_n[i] = (int)(r / d.Ho);
  }
  // B-load mapping: n = t%64, k = t/64 + 4*i
  const int bn = t & 63, bk0 = t >> 6;

  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int ky = 0; ky < d.kh; ++ky) {
    for (int kx = 0; kx < d.kw; ++kx) {
      // per-tap source pixel of each of this thread's 4 A rows
      long long a_off[4];
      bool a_val[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int iy, ix;
        bool ok = a_ok[i];
        if (TRANSPOSED) {
          int ny = a_y[i] + d.pad - ky, nx = a_x[i] + d.pad - kx;
          ok = ok && ny >= 0 && nx >= 0 && (ny % d.stride) == 0 && (nx % d.stride) == 0;
          iy = ny / d.stride;
          ix = nx / d.stride;
        } else {
          iy = a_y[i] * d.stride - d.pad + ky;
          ix = a_x[i] * d.stride - d.pad + kx;
        }
        ok = ok && iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi;
        a_val[i] = ok;
        a_off[i] = ok ? (((long long)a_n[i] * d.Hi + iy) * d.Wi + ix) * d.ld_in : 0;
      }
      const float* wt = w + ky * d.w_sky + kx * d.w_skx;
      for (int c0 = 0; c0 < d.Ci; c0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + ak;
          As[ak][(t >> 4) + 16 * i] = (a_val[i] && c < d.Ci) ? ldf(in + a_off[i] + c) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = bk0 + 4 * i, c = c0 + kk, co = n0 + bn;
          Bs[kk][bn] = (c < d.Ci && co < d.Co) ? __ldg(wt + co * d.w_sco + c * d.w_sci) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= d.Co) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + co);
      T* o = out + m * d.ld_out + co;
      if (d.accumulate) v += ldf(o);
      if (d.dact) v *= act_bwd(d.dact, ldf(aux + m * d.ld_aux + co));
      if (pre_out) stf(pre_out + m * d.ld_pre + co, v);
      stf(o, act_fwd(d.act, v));
    }
  }
}

// weight gradient: for each tap, dW[co,ci] = sum over pixels dout[p,co] * in[p@tap,ci]
// grid: (co tiles * ci tiles, taps, pixel splits)
template <typename T>
__global__ void __launch_bounds__(NT) k_igemm_wgrad(dsgan_conv_desc d, const T* __restrict__ in,
                                                     const T* __restrict__ dout, float* __restrict__ dw,
                                                     int ci_tiles, long long chunk) {
  __shared__ float As[BK][BM + 4];  // dout: [pixel][co]
  __shared__ float Bs[BK][BN + 4];  // in:   [pixel][ci]
  const int t = threadIdx.x;
  const int co0 = (blockIdx.x / ci_tiles) * BM, ci0 = (blockIdx.x % ci_tiles) * BN;
  const int ky = blockIdx.y / d.kw, kx = blockIdx.y % d.kw;
  const long long M = (long long)d.N * d.Ho * d.Wo;
  const long long p_begin = (long long)blockIdx.z * chunk;
  const long long p_end = min(p_begin + chunk, M);
  const int ty = t >> 4, tx = t & 15;
  const int lc = t & 63, lp0 = t >> 6;  // load mapping: channel lane, pixel lane (4 pixels / thread)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = lp0 + 4 * i;
      const long long p = p0 + pp;
      float av = 0.f, bv = 0.f;
      if (p < p_end) {
        const int co = co0 + lc, ci = ci0 + lc;
        if (co < d.Co) av = ldf(dout + p * d.ld_out + co);
        if (ci < d.Ci) {
          int ox = (int)(p % d.Wo);
          long long r = p / d.Wo;
          int oy = (int)(r % d.Ho), n = (int)(r / d.Ho);
          int iy = oy * d.stride - d.pad + ky, ix = ox * d.stride - d.pad + kx;
          if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
            bv = ldf(in + (((long long)n * d.Hi + iy) * d.Wi + ix) * d.ld_in + ci);
        }
      }
      As[pp][lc] = av;
      Bs[pp][lc] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* wt = dw + ky * d.w_sky + kx * d.w_skx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= d.Co) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < d.Ci) atomicAdd(wt + co * d.w_sco + ci * d.w_sci, acc[i][j]);
    }
  }
}

template <typename T>
__global__ void k_colsum(const T* __restrict__ x, int ld, long long npix, int C, float* __restrict__ out,
                         long long chunk) {
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(p0 + chunk, npix);
  const int cl = min(C, (int)blockDim.x), pl = blockDim.x / cl;
  const int tc = threadIdx.x % cl, tp = threadIdx.x / cl;
  if (tp >= pl) return;
  for (int c = tc; c < C; c += cl) {
    float a = 0.f;
    for (long long p = p0 + tp; p < p1; p += pl) a += ldf(x + p * ld + c);
    atomicAdd(out + c, a);
  }
}

// bf16 fast path of colsum: 8 channels (16 B) per thread, 4 independent row loads in flight per thread, smem
// reduction over the block's pixel lanes, one global atomic per (block, channel).  grid = (pixel chunks, channel blocks).
__global__ void __launch_bounds__(256) k_colsum_v8(const bf16* __restrict__ x, int ld, long long npix, int C,
                                                    float* __restrict__ out, long long chunk) {
  __shared__ float sacc[256];
  const int groups = (C + 7) / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;  // ragged C: pitch == ceil8(C)
  const int tg = threadIdx.x % gl, tp = threadIdx.x / gl;
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(p0 + chunk, npix);
  const int cb = blockIdx.y * gl * 8, c0 = cb + tg * 8;
  sacc[threadIdx.x] = 0.f;
  __syncthreads();
  if (tp < pl && c0 < C) {
    float2 a[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    const bf16* xp = x + c0;
    long long p = p0 + tp;
    for (; p + 3 * pl < p1; p += 4 * pl) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(xp + (p + (long long)j * pl) * ld));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          a[e] = __fadd2_rn(a[e], __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e])));
      }
    }
    for (; p < p1; p += pl) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xp + p * ld));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        a[e] = __fadd2_rn(a[e], __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e])));
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&sacc[tg * 8 + 2 * e], a[e].x);
      atomicAdd(&sacc[tg * 8 + 2 * e + 1], a[e].y);
    }
  }
  __syncthreads();
  if (threadIdx.x < gl * 8 && cb + threadIdx.x < C) atomicAdd(out + cb + threadIdx.x, sacc[threadIdx.x]);
}
}  // namespace

extern "C" {
int dsgan_conv_fwd(const dsgan_conv_desc* d, const void* in, const float* w, const float* bias, void* out,
                   void* pre_out, const void* aux, void* stream) {
  DS_REQUIRE(d && in && w && out, "conv_fwd: null argument");
  DS_REQUIRE(!d->dact || aux, "conv_fwd: dact needs aux");
  DS_REQUIRE(d->stride >= 1 && d->kh >= 1 && d->kw >= 1, "conv_fwd: bad geometry");
  const long long M = (long long)d->N * d->Ho * d->Wo;
  dim3 grid(cdiv(M, BM), cdiv(d->Co, BN));
  cudaStream_t s = (cudaStream_t)stream;
  DS_DISPATCH_DT(d->dtype, {
    if (d->transposed)
      k_igemm<T, 1><<<grid, NT, 0, s>>>(*d, (const T*)in, w, bias, (T*)out, (T*)pre_out, (const T*)aux);
    else
      k_igemm<T, 0><<<grid, NT, 0, s>>>(*d, (const T*)in, w, bias, (T*)out, (T*)pre_out, (const T*)aux);
  });
  return DS_LAUNCHED("conv_fwd");
}

int dsgan_conv_wgrad(const dsgan_conv_desc* d, const void* in, const void* dout, float* dw, void* stream) {
  DS_REQUIRE(d && in && dout && dw, "conv_wgrad: null argument");
  const long long M = (long long)d->N * d->Ho * d->Wo;
  const int co_tiles = cdiv(d->Co, BM), ci_tiles = cdiv(d->Ci, BN), taps = d->kh * d->kw;
  long long base = (long long)co_tiles * ci_tiles * taps;
  long long splits = (148LL * 4 + base - 1) / base;
  long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long chunk = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + chunk - 1) / chunk;
  dim3 grid(co_tiles * ci_tiles, taps, (unsigned)splits);
  DS_DISPATCH_DT(d->dtype, (k_igemm_wgrad<T><<<grid, NT, 0, (cudaStream_t)stream>>>(*d, (const T*)in, (const T*)dout,
                                                                                   dw, ci_tiles, chunk)));
  return DS_LAUNCHED("conv_wgrad");
}

int dsgan_colsum(const void* x, int dtype, int ld, long long npix, int C, float* out, void* stream) {
  if (dtype == DT_BF16 && ld % 8 == 0 && (C % 8 == 0 || ld == (C + 7) / 8 * 8) && ((uintptr_t)x % 16 == 0)) {
    const int groups = (C + 7) / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;
    const int cblocks = (C + gl * 8 - 1) / (gl * 8);
    // ~8 CTAs per SM over the whole launch, at least 16 rows per pixel lane so the 4-deep unroll is used
    long long blocks = (148 * 8 + cblocks - 1) / cblocks;
    long long chunk = (npix + blocks - 1) / blocks;
    if (chunk < 16LL * pl) chunk = 16LL * pl;
    blocks = (npix + chunk - 1) / chunk;
    dim3 grid((unsigned)blocks, (unsigned)cblocks);
    k_colsum_v8<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld, npix, C, out, chunk);
    return DS_LAUNCHED("colsum_v8");
  }
  long long chunk = 1024;
  long long blocks = (npix + chunk - 1) / chunk;
  if (blocks > 148 * 8) { chunk = (npix + 148 * 8 - 1) / (148 * 8); blocks = (npix + chunk - 1) / chunk; }
  DS_DISPATCH_DT(dtype, (k_colsum<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, ld, npix, C,
                                                                                        out, chunk)));
  return DS_LAUNCHED("colsum");
}
}
