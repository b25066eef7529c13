// Loss kernels: GAN (BCE-with-logits / MSE), L1, TV, and the SSIM / MS-SSIM family (separable 11-tap Gaussian,
// valid padding) with their backward passes.  All are HBM-bound; SSIM stages image tiles + halos in shared
// memory and evaluates the five Gaussian moments in one pass.
// Reference: networks.py:143-163, pix2pix_model.py:177-197, MS_SSIM.py:9-225.
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
__constant__ float c_win[11];
bool g_win_ready = false;
int ensure_window() {
  if (g_win_ready) return 0;
  // _fspecial_gauss_1d(11, 1.5), MS_SSIM.py:9-23 (fp32 arithmetic like the reference)
  float g[11], s = 0.f;
  for (int i = 0; i < 11; ++i) { float c = (float)(i - 5); g[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f)); s += g[i]; }
  for (int i = 0; i < 11; ++i) g[i] /= s;
  cudaError_t e = cudaMemcpyToSymbol(c_win, g, sizeof(g));
  if (e != cudaSuccess) { set_error("gauss window upload: %s", cudaGetErrorString(e)); return 1; }
  g_win_ready = true;
  return 0;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

template <typename T>
__global__ void k_gan_loss(const T* __restrict__ pred, long long n, int ld, float target, int mode, float loss_scale,
                           float* __restrict__ loss, float grad_scale, T* __restrict__ dpred) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float invn = 1.0f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = ldf(pred + i * ld);
    float l, g;
    if (mode == 0) {  // BCEWithLogits: max(x,0) - x*t + log1p(exp(-|x|))   (Q12)
      l = fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x)));
      g = 1.0f / (1.0f + expf(-x)) - target;
    } else if (mode == 1) {  // MSELoss on the raw prediction
      const float dlt = x - target;
      l = dlt * dlt;
      g = 2.f * dlt;
    } else {  // MSELoss on sigmoid(x): the `--no_lsgan` pairing (pix2pix_model.py:98,112-114, Q6)
      const float sg = 1.0f / (1.0f + expf(-x)), dlt = sg - target;
      l = dlt * dlt;
      g = 2.f * dlt * sg * (1.f - sg);
    }
    acc += l;
    if (dpred) stf(dpred + i * ld, g * grad_scale * invn);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, acc * loss_scale * invn);
}

template <typename T>
__global__ void k_l1(const T* __restrict__ a, const T* __restrict__ b, long long n, float* __restrict__ loss,
                     float grad_scale, T* __restrict__ da, int accum, int relu_mask) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float invn = 1.0f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = ldf(a + i) - ldf(b + i);
    acc += fabsf(d);
    if (da) {
      float g = sgn(d) * grad_scale * invn;
      if (relu_mask && !(ldf(a + i) > 0.f)) g = 0.f;
      if (accum) g += ldf(da + i);
      stf(da + i, g);
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, acc * invn);
}

// bf16 fast path of k_l1: 8 elements (16 bytes) per thread and iteration
__global__ void __launch_bounds__(256) k_l1_v8(const bf16* __restrict__ a, const bf16* __restrict__ b, long long n8, float invn,
                                                float* __restrict__ loss, float grad_scale, bf16* __restrict__ da, int accum,
                                                int relu_mask) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float gs = grad_scale * invn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + i), ub = __ldg(reinterpret_cast<const uint4*>(b) + i);
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t wo[4] = {0, 0, 0, 0};
    if (da && accum) { const uint4 uo = reinterpret_cast<const uint4*>(da)[i]; wo[0] = uo.x; wo[1] = uo.y; wo[2] = uo.z; wo[3] = uo.w; }
    uint32_t wg[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wa[q]));
      const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb[q]));
      const float d0 = fa.x - fb.x, d1 = fa.y - fb.y;
      acc += fabsf(d0) + fabsf(d1);
      float g0 = sgn(d0) * gs, g1 = sgn(d1) * gs;
      if (relu_mask) { if (!(fa.x > 0.f)) g0 = 0.f; if (!(fa.y > 0.f)) g1 = 0.f; }
      if (accum) { const float2 fo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wo[q])); g0 += fo.x; g1 += fo.y; }
      __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
      wg[q] = *reinterpret_cast<uint32_t*>(&h);
    }
    if (da) reinterpret_cast<uint4*>(da)[i] = make_uint4(wg[0], wg[1], wg[2], wg[3]);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, acc * invn);
}

__global__ void k_tv(const float* __restrict__ x, int H, int W, long long total, float inv_denom,
                     float* __restrict__ loss, float grad_scale, float* __restrict__ dx) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W), yy = (int)((i / W) % H);
    const float v = x[i];
    float g = 0.f;
    if (xx + 1 < W) { const float d = x[i + 1] - v; acc += fabsf(d); g -= sgn(d); }
    if (xx > 0) g += sgn(v - x[i - 1]);
    if (yy + 1 < H) { const float d = x[i + W] - v; acc += fabsf(d); g -= sgn(d); }
    if (yy > 0) g += sgn(v - x[i - W]);
    if (dx) dx[i] += g * grad_scale * inv_denom;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_denom);
}

#include "ssim_v2.cuh"
#include "ssim_v3.cuh"

// ---------------------------------------------------------------------------------------------
// (first-generation kernels, kept as the readable reference of the tiling; the ABI dispatches to ssim_v2.cuh)
// SSIM forward: 32x32 output tile per block, 42x42 input tiles of X and Y staged in smem.
constexpr int TS = 32, HALO = 10, TI = TS + HALO;  // 42

__global__ void __launch_bounds__(256) k_ssim_fwd(const float* __restrict__ X, const float* __restrict__ Y, int H,
                                                   int W, float C1, float C2, float* __restrict__ sums) {
  __shared__ float sX[TI][TI + 1], sY[TI][TI + 1];
  __shared__ float sH[5][TI][TS + 1];
  __shared__ float red[32];
  const int nc = blockIdx.z, y0 = blockIdx.y * TS, x0 = blockIdx.x * TS;
  const float* xp = X + (long long)nc * H * W;
  const float* yp = Y + (long long)nc * H * W;
  for (int i = threadIdx.x; i < TI * TI; i += 256) {
    const int r = i / TI, c = i % TI, gy = y0 + r, gx = x0 + c;
    const bool ok = gy < H && gx < W;
    sX[r][c] = ok ? xp[(long long)gy * W + gx] : 0.f;
    sY[r][c] = ok ? yp[(long long)gy * W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TI * TS; i += 256) {
    const int r = i / TS, c = i % TS;
    float m1 = 0, m2 = 0, xx = 0, yy = 0, xy = 0;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float w = c_win[t], a = sX[r][c + t], b = sY[r][c + t];
      m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
      xx = fmaf(w, a * a, xx); yy = fmaf(w, b * b, yy); xy = fmaf(w, a * b, xy);
    }
    sH[0][r][c] = m1; sH[1][r][c] = m2; sH[2][r][c] = xx; sH[3][r][c] = yy; sH[4][r][c] = xy;
  }
  __syncthreads();
  float acc_s = 0.f, acc_c = 0.f;
  const int Hv = H - HALO, Wv = W - HALO;
  for (int i = threadIdx.x; i < TS * TS; i += 256) {
    const int r = i / TS, c = i % TS;
    if (y0 + r >= Hv || x0 + c >= Wv) continue;
    float m[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float w = c_win[t];
#pragma unroll
      for (int k = 0; k < 5; ++k) m[k] = fmaf(w, sH[k][r + t][c], m[k]);
    }
    const float mu1 = m[0], mu2 = m[1];
    const float s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
    const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
    const float ss = ((2.f * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs;
    acc_s += ss; acc_c += cs;
  }
  acc_s = block_sum(acc_s, red);
  acc_c = block_sum(acc_c, red);
  if (threadIdx.x == 0) { atomicAdd(sums + 2 * nc, acc_s); atomicAdd(sums + 2 * nc + 1, acc_c); }
}

// SSIM backward w.r.t. Y.  Output tile 32x32 of dY; a/b/c maps on 42x42; X,Y tile 52x52.
constexpr int TB = TS + 2 * HALO;  // 52
struct BwdSmem {
  float X[TB][TB + 1], Y[TB][TB + 1];
  float Hm[5][TB][TI + 1];  // horizontal moments; later reused for the horizontal pass of a,b,c
  float ABC[3][TI][TI + 1];
};

__global__ void __launch_bounds__(256) k_ssim_bwd(const float* __restrict__ X, const float* __restrict__ Y, int H,
                                                   int W, float C1, float C2, const float* __restrict__ coef,
                                                   float* __restrict__ dY, int accum) {
  extern __shared__ __align__(16) unsigned char smraw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(smraw);
  const int nc = blockIdx.z, qy0 = blockIdx.y * TS, qx0 = blockIdx.x * TS;
  const float gs = coef[2 * nc], gc = coef[2 * nc + 1];
  const float* xp = X + (long long)nc * H * W;
  const float* yp = Y + (long long)nc * H * W;
  for (int i = threadIdx.x; i < TB * TB; i += 256) {
    const int r = i / TB, c = i % TB, gy = qy0 - HALO + r, gx = qx0 - HALO + c;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    S.X[r][c] = ok ? xp[(long long)gy * W + gx] : 0.f;
    S.Y[r][c] = ok ? yp[(long long)gy * W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TB * TI; i += 256) {
    const int r = i / TI, c = i % TI;
    float m1 = 0, m2 = 0, xx = 0, yy = 0, xy = 0;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float w = c_win[t], a = S.X[r][c + t], b = S.Y[r][c + t];
      m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
      xx = fmaf(w, a * a, xx); yy = fmaf(w, b * b, yy); xy = fmaf(w, a * b, xy);
    }
    S.Hm[0][r][c] = m1; S.Hm[1][r][c] = m2; S.Hm[2][r][c] = xx; S.Hm[3][r][c] = yy; S.Hm[4][r][c] = xy;
  }
  __syncthreads();
  const int Hv = H - HALO, Wv = W - HALO;
  for (int i = threadIdx.x; i < TI * TI; i += 256) {
    const int r = i / TI, c = i % TI;
    const int py = qy0 - HALO + r, px = qx0 - HALO + c;  // window origin p
    float a = 0.f, b = 0.f, cc = 0.f;
    if (py >= 0 && py < Hv && px >= 0 && px < Wv) {
      float m[5] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int t = 0; t < 11; ++t) {
        const float w = c_win[t];
#pragma unroll
        for (int k = 0; k < 5; ++k) m[k] = fmaf(w, S.Hm[k][r + t][c], m[k]);
      }
      const float mu1 = m[0], mu2 = m[1];
      const float A1 = 2.f * mu1 * mu2 + C1, B1 = mu1 * mu1 + mu2 * mu2 + C1;
      const float A2 = 2.f * (m[4] - mu1 * mu2) + C2, B2 = (m[2] - mu1 * mu1) + (m[3] - mu2 * mu2) + C2;
      const float iB1 = 1.f / B1, iB2 = 1.f / B2;
      const float cs = A2 * iB2, lum = A1 * iB1, Sv = lum * cs;
      // d ssim / d{mu2, Eyy, Exy}
      const float as = 2.f * mu1 * (A2 - A1) * iB1 * iB2 + 2.f * mu2 * Sv * (iB2 - iB1);
      const float bs = -Sv * iB2;
      const float cS = 2.f * lum * iB2;
      // d cs / d{mu2, Eyy, Exy}
      const float ac = (-2.f * mu1 + 2.f * mu2 * cs) * iB2;
      const float bc = -cs * iB2;
      const float cC = 2.f * iB2;
      a = gs * as + gc * ac; b = gs * bs + gc * bc; cc = gs * cS + gc * cC;
    }
    S.ABC[0][r][c] = a; S.ABC[1][r][c] = b; S.ABC[2][r][c] = cc;
  }
  __syncthreads();
  // horizontal pass of the transposed filter: hA[k][i][v] = sum_t w[t] * abc[k][i][v + 10 - t]
  float (*hA)[TI][TS + 1] = reinterpret_cast<float (*)[TI][TS + 1]>(&S.Hm[0][0][0]);
  for (int i = threadIdx.x; i < TI * TS; i += 256) {
    const int r = i / TS, v = i % TS;
    float o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float w = c_win[t];
      o0 = fmaf(w, S.ABC[0][r][v + HALO - t], o0);
      o1 = fmaf(w, S.ABC[1][r][v + HALO - t], o1);
      o2 = fmaf(w, S.ABC[2][r][v + HALO - t], o2);
    }
    hA[0][r][v] = o0; hA[1][r][v] = o1; hA[2][r][v] = o2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TS * TS; i += 256) {
    const int u = i / TS, v = i % TS, gy = qy0 + u, gx = qx0 + v;
    if (gy >= H || gx >= W) continue;
    float o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float w = c_win[t];
      o0 = fmaf(w, hA[0][u + HALO - t][v], o0);
      o1 = fmaf(w, hA[1][u + HALO - t][v], o1);
      o2 = fmaf(w, hA[2][u + HALO - t][v], o2);
    }
    float g = o0 + 2.f * S.Y[u + HALO][v + HALO] * o1 + S.X[u + HALO][v + HALO] * o2;
    float* o = dY + (long long)nc * H * W + (long long)gy * W + gx;
    if (accum) g += *o;
    *o = g;
  }
}

// F.avg_pool2d(x, kernel_size=2, padding=(H % 2, W % 2)) (MS_SSIM.py:214-216; count_include_pad=True): an odd side is
// zero-padded by one row / column on BOTH sides, output floor(H/2) + 1, window i covers rows 2i - ph, 2i - ph + 1; the
// divisor stays 4.  Even sizes: ph = pw = 0, the plain 2x2 mean.
__global__ void k_avgpool2_fwd(const float* __restrict__ x, float* __restrict__ y, int H, int W, long long total) {
  const int ph = H & 1, pw = W & 1, Ho = H / 2 + ph, Wo = W / 2 + pw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const long long r = i / Wo;
    const int oy = (int)(r % Ho);
    const long long nc = r / Ho;
    const int y0 = 2 * oy - ph, x0 = 2 * ox - pw;
    const float* p = x + nc * H * W;
    float a = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) a += p[(long long)yy * W + xx];
      }
    y[i] = 0.25f * a;
  }
}
// the X and the Y plane sets of one pyramid level in one launch
__global__ void k_avgpool2_fwd_pair(const float* __restrict__ x1, float* __restrict__ y1, const float* __restrict__ x2,
                                    float* __restrict__ y2, int H, int W, long long total) {
  const int ph = H & 1, pw = W & 1, Ho = H / 2 + ph, Wo = W / 2 + pw;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < 2 * total; j += (long long)gridDim.x * blockDim.x) {
    const bool second = j >= total;
    const long long i = second ? j - total : j;
    const float* x = second ? x2 : x1;
    float* y = second ? y2 : y1;
    const int ox = (int)(i % Wo);
    const long long r = i / Wo;
    const int oy = (int)(r % Ho);
    const long long nc = r / Ho;
    const int y0 = 2 * oy - ph, x0 = 2 * ox - pw;
    const float* p = x + nc * H * W;
    float a = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) a += p[(long long)yy * W + xx];
      }
    y[i] = 0.25f * a;
  }
}
__global__ void k_avgpool2_bwd(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int accum,
                               long long total) {
  const int ph = H & 1, pw = W & 1, Ho = H / 2 + ph, Wo = W / 2 + pw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const long long r = i / W;
    const int yy = (int)(r % H);
    const long long nc = r / H;
    const int oy = (yy + ph) / 2, ox = (xx + pw) / 2;     // the one window that contains input pixel (yy, xx)
    float g = 0.f;
    if (oy < Ho && ox < Wo) g = 0.25f * dy[(nc * Ho + oy) * Wo + ox];
    dx[i] = accum ? dx[i] + g : g;
  }
}

__global__ void k_msssim_combine(const float* __restrict__ sums, const float* __restrict__ inv_sizes,
                                 const float* __restrict__ weights, int L, int NC, float out_scale,
                                 float* __restrict__ val, float grad_scale, float* __restrict__ coef) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int nc = blockIdx.x * blockDim.x + threadIdx.x; nc < NC; nc += gridDim.x * blockDim.x) {
    float v[8];
    float prod = 1.f;
    for (int l = 0; l < L; ++l) {
      const int slot = (l == L - 1) ? 0 : 1;
      float m = sums[((long long)l * NC + nc) * 2 + slot] * inv_sizes[l];
      m = m > 0.f ? m : 0.f;  // torch.relu, MS_SSIM.py:213,218
      v[l] = m;
      prod *= powf(m, weights[l]);
    }
    acc += prod;
    if (coef) {
      for (int l = 0; l < L; ++l) {
        const int slot = (l == L - 1) ? 0 : 1;
        // d prod / d mean_l = w_l * prod / v_l ; guarded to 0 where v_l <= 0 (the reference yields inf*0 = NaN, Q9)
        const float g = v[l] > 0.f ? weights[l] * prod / v[l] * inv_sizes[l] * grad_scale / (float)NC : 0.f;
        coef[((long long)l * NC + nc) * 2 + slot] = g;
        coef[((long long)l * NC + nc) * 2 + (1 - slot)] = 0.f;
      }
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(val, acc * out_scale / (float)NC);
}

__global__ void k_ssim_combine(const float* __restrict__ sums, float inv_size, int NC, float out_scale,
                               float* __restrict__ val, float grad_scale, float* __restrict__ coef) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int nc = blockIdx.x * blockDim.x + threadIdx.x; nc < NC; nc += gridDim.x * blockDim.x) {
    acc += sums[2 * nc] * inv_size;
    if (coef) { coef[2 * nc] = grad_scale * inv_size / (float)NC; coef[2 * nc + 1] = 0.f; }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(val, acc * out_scale / (float)NC);
}

inline int grid_for(long long n, int block, int cap = 148 * 8) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

// Row segments of the row-streaming kernels.  Every segment pays 10 warm-up rows, and the launch runs in waves of `slots`
// resident warps: pick the segment count that minimises waves x (rows per segment + 10).  (A fixed "two segments" choice put
// 3072 warps on 2960 slots -- two waves, the second 4 % full.)
template <typename K>
static int ssim_slots(K kern, int* cache) {
  if (*cache == 0) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * SS_WARPS, 0);
    *cache = sms * (occ > 0 ? occ : 1) * SS_WARPS;
  }
  return *cache;
}
static void ssim_tasks(int NC, int rows, int cols, int slots, int* nstrip, int* nseg, int* rs, long long* ntasks) {
  *nstrip = (cols + SS_W - 1) / SS_W;
  const long long per = (long long)NC * *nstrip;
  int maxseg = rows / 16;
  if (maxseg < 1) maxseg = 1;
  if (maxseg > 64) maxseg = 64;
  long long best = -1;
  int best_n = 1;
  for (int n = 1; n <= maxseg; ++n) {
    const long long waves = (per * n + slots - 1) / slots;
    const long long cost = waves * ((rows + n - 1) / n + 10);
    if (best < 0 || cost < best) { best = cost; best_n = n; }
  }
  *rs = (rows + best_n - 1) / best_n;
  *nseg = (rows + *rs - 1) / *rs;
  *ntasks = per * *nseg;
}

extern "C" {
int dsgan_gan_loss(const void* pred, int dtype, long long n, int ld, float target, int mode, float loss_scale,
                   float* loss, float grad_scale, void* dpred, void* stream) {
  DS_DISPATCH_DT(dtype, (k_gan_loss<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)pred, n, ld, target, mode, loss_scale, loss, grad_scale, (T*)dpred)));
  return DS_LAUNCHED("gan_loss");
}
int dsgan_l1_loss(const void* a, const void* b, int dtype, long long n, float* loss, float grad_scale, void* da,
                  int accumulate, int relu_mask, void* stream) {
  if (dtype == DT_BF16 && n % 8 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)da) % 16 == 0)) {
    k_l1_v8<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, n / 8, 1.0f / (float)n, loss,
                                                                   grad_scale, (bf16*)da, accumulate, relu_mask);
    return DS_LAUNCHED("l1_loss_v8");
  }
  DS_DISPATCH_DT(dtype, (k_l1<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, n, loss,
                                                                                    grad_scale, (T*)da, accumulate, relu_mask)));
  return DS_LAUNCHED("l1_loss");
}
int dsgan_tv_loss(const float* x, int NC, int H, int W, float denom, float* loss, float grad_scale, float* dx,
                  void* stream) {
  const long long total = (long long)NC * H * W;
  k_tv<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, H, W, total, 1.0f / denom, loss, grad_scale, dx);
  return DS_LAUNCHED("tv_loss");
}
int dsgan_ssim_fwd(const float* X, const float* Y, int NC, int H, int W, float C1, float C2, float* sums, float* moments,
                   void* stream) {
  DS_REQUIRE(H >= 11 && W >= 11, "ssim: image %dx%d smaller than the 11-tap window", H, W);
  DS_REQUIRE(NC <= 65535, "ssim: too many planes (%d)", NC);
  if (ensure_window()) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * NC, s);
  int nstrip, nseg, rs;
  long long ntasks;
  static int slots_s = 0, slots_n = 0;
  ssim_tasks(NC, H - 10, W - 10, moments ? ssim_slots(k_ssim_fwd3<true>, &slots_s) : ssim_slots(k_ssim_fwd3<false>, &slots_n),
             &nstrip, &nseg, &rs, &ntasks);
  const unsigned blocks = (unsigned)((ntasks + SS_WARPS - 1) / SS_WARPS);
  if (moments) k_ssim_fwd3<true><<<blocks, 32 * SS_WARPS, 0, s>>>(X, Y, H, W, C1, C2, sums, moments, NC, nstrip, nseg, rs, ntasks);
  else k_ssim_fwd3<false><<<blocks, 32 * SS_WARPS, 0, s>>>(X, Y, H, W, C1, C2, sums, nullptr, NC, nstrip, nseg, rs, ntasks);
  return DS_LAUNCHED("ssim_fwd");
}
int dsgan_ssim_bwd(const float* X, const float* Y, int NC, int H, int W, float C1, float C2, const float* coef,
                   const float* moments, float* dY, int accumulate, const float* gnext, void* stream) {
  DS_REQUIRE(H >= 11 && W >= 11, "ssim: image %dx%d smaller than the 11-tap window", H, W);
  DS_REQUIRE(NC <= 65535, "ssim: too many planes (%d)", NC);
  if (ensure_window()) return 1;
  if (moments) {   // from the moments the forward pass stored: row-streaming transposed filter (ssim_v3.cuh)
    int nstrip, nseg, rs;
    long long ntasks;
    static int slots_b = 0;
    ssim_tasks(NC, H, W, ssim_slots(k_ssim_bwd3, &slots_b), &nstrip, &nseg, &rs, &ntasks);
    const unsigned blocks = (unsigned)((ntasks + SS_WARPS - 1) / SS_WARPS);
    k_ssim_bwd3<<<blocks, 32 * SS_WARPS, 0, (cudaStream_t)stream>>>(X, Y, moments, H, W, C1, C2, coef, dY, accumulate, gnext, NC,
                                                                     nstrip, nseg, rs, ntasks);
    return DS_LAUNCHED("ssim_bwd");
  }
  DS_REQUIRE(gnext == nullptr, "ssim_bwd: the fused avg_pool adjoint needs the stored moments");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ssim_bwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Bwd2Smem));
    if (e != cudaSuccess) { set_error("ssim_bwd smem attr: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  dim3 grid(cdiv(W, V_TS), cdiv(H, V_TS), NC);
  k_ssim_bwd2<<<grid, 256, sizeof(Bwd2Smem), (cudaStream_t)stream>>>(X, Y, H, W, C1, C2, coef, dY, accumulate);
  return DS_LAUNCHED("ssim_bwd");
}
int dsgan_avgpool2_fwd2(const float* x1, float* y1, const float* x2, float* y2, int NC, int H, int W, void* stream) {
  const long long total = (long long)NC * (H / 2 + (H & 1)) * (W / 2 + (W & 1));
  k_avgpool2_fwd_pair<<<grid_for(2 * total, 256), 256, 0, (cudaStream_t)stream>>>(x1, y1, x2, y2, H, W, total);
  return DS_LAUNCHED("avgpool2_fwd");
}
int dsgan_avgpool2_fwd(const float* x, float* y, int NC, int H, int W, void* stream) {
  const long long total = (long long)NC * (H / 2 + (H & 1)) * (W / 2 + (W & 1));
  k_avgpool2_fwd<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, total);
  return DS_LAUNCHED("avgpool2_fwd");
}
int dsgan_avgpool2_bwd(const float* dy, float* dx, int NC, int H, int W, int accumulate, void* stream) {
  const long long total = (long long)NC * H * W;
  k_avgpool2_bwd<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, H, W, accumulate, total);
  return DS_LAUNCHED("avgpool2_bwd");
}
int dsgan_msssim_combine(const float* sums, const float* inv_sizes, const float* weights, int L, int NC,
                         float out_scale, float* val, float grad_scale, float* coef, void* stream) {
  DS_REQUIRE(L >= 1 && L <= 8, "msssim_combine: L=%d", L);
  k_msssim_combine<<<grid_for(NC, 128, 64), 128, 0, (cudaStream_t)stream>>>(sums, inv_sizes, weights, L, NC, out_scale,
                                                                            val, grad_scale, coef);
  return DS_LAUNCHED("msssim_combine");
}
int dsgan_ssim_combine(const float* sums, float inv_size, int NC, float out_scale, float* val, float grad_scale,
                       float* coef, void* stream) {
  k_ssim_combine<<<grid_for(NC, 128, 64), 128, 0, (cudaStream_t)stream>>>(sums, inv_size, NC, out_scale, val,
                                                                          grad_scale, coef);
  return DS_LAUNCHED("ssim_combine");
}
}
