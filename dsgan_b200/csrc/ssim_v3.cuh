// Row-streaming SSIM kernels (round 2; included by loss.cu inside its anonymous namespace, after c_win).
//
// One WARP owns a strip of 32 output columns of one plane and walks down a segment of rows; lane = column.  Per input row:
// the 42 columns the strip needs are staged in a per-warp shared-memory row, every lane applies the horizontal 11-tap filter
// to its column (five moments), and scatters the result into ELEVEN pending output rows held in registers (vertical filter
// in scatter form, slot = row mod 11 with the row loop unrolled by 11 so every slot index is static).  A row leaves the
// register window after exactly 11 updates.  Compared with the tiled kernels (ssim_v2.cuh: 32x32 tiles, both passes through
// shared memory, three block barriers, a 42/32 halo in BOTH directions): no block barrier, no shared-memory round trip between
// the passes, no vertical halo recompute except 10 warm-up rows per segment, exactly 110 filter FMAs per pixel.
//
// The forward pass can store the five filtered moments (mu1, mu2, E[xx], E[yy], E[xy]) of every window; the backward pass then
// starts from them instead of recomputing them on a 52x52 halo: it forms the three adjoint maps pointwise and applies the
// transposed (full-correlation) filter with the same streaming scheme.  20 bytes per pixel of extra traffic buy ~150 FMAs.
constexpr int SS_W = 32;            // output columns per warp
constexpr int SS_IN = SS_W + 10;    // staged input columns
constexpr int SS_WARPS = 4;

struct SsimTask {
  int nc, r0, r1, c0;
};
__device__ __forceinline__ bool ssim_task(long long task, long long ntasks, int nstrip, int nseg, int rs, int rows, SsimTask& t) {
  if (task >= ntasks) return false;
  t.c0 = (int)(task % nstrip) * SS_W;
  const int seg = (int)((task / nstrip) % nseg);
  t.nc = (int)(task / ((long long)nstrip * nseg));
  t.r0 = seg * rs;
  t.r1 = min(t.r0 + rs, rows);
  return t.r0 < t.r1;
}

template <bool STORE>
__global__ void __launch_bounds__(32 * SS_WARPS) k_ssim_fwd3(const float* __restrict__ X, const float* __restrict__ Y, int H,
                                                             int W, float C1, float C2, float* __restrict__ sums,
                                                             float* __restrict__ mom, int NC, int nstrip, int nseg, int rs,
                                                             long long ntasks) {
  __shared__ float srow[SS_WARPS][2][2][SS_IN + 2];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hv = H - 10, Wv = W - 10;
  SsimTask t;
  if (!ssim_task((long long)blockIdx.x * SS_WARPS + wid, ntasks, nstrip, nseg, rs, Hv, t)) return;
  const float* xp = X + (size_t)t.nc * H * W;
  const float* yp = Y + (size_t)t.nc * H * W;
  float w[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) w[k] = c_win[k];
  float acc[11][5];
#pragma unroll
  for (int s = 0; s < 11; ++s)
#pragma unroll
    for (int m = 0; m < 5; ++m) acc[s][m] = 0.f;
  float sum_s = 0.f, sum_c = 0.f;
  const int ox = t.c0 + lane;
  const int c_a = t.c0 + lane, c_b = t.c0 + 32 + lane;
  int buf = 0;
  const int r_end = t.r1 + 10;                    // input rows [r0, r1 + 10)
  for (int base = t.r0; base < r_end; base += 11) {
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int r = base + j;
      if (r < r_end) {                            // warp-uniform
        float* bx = srow[wid][buf][0];
        float* by = srow[wid][buf][1];
        buf ^= 1;
        bx[lane] = c_a < W ? __ldg(xp + (size_t)r * W + c_a) : 0.f;
        by[lane] = c_a < W ? __ldg(yp + (size_t)r * W + c_a) : 0.f;
        if (lane < 10) {
          bx[32 + lane] = c_b < W ? __ldg(xp + (size_t)r * W + c_b) : 0.f;
          by[32 + lane] = c_b < W ? __ldg(yp + (size_t)r * W + c_b) : 0.f;
        }
        __syncwarp();
        float h[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const float a = bx[lane + k], b = by[lane + k];
          h[0] = fmaf(w[k], a, h[0]);
          h[1] = fmaf(w[k], b, h[1]);
          h[2] = fmaf(w[k], a * a, h[2]);
          h[3] = fmaf(w[k], b * b, h[3]);
          h[4] = fmaf(w[k], a * b, h[4]);
        }
#pragma unroll
        for (int k = 0; k < 11; ++k) {            // input row r is tap k of output row r - k
          constexpr int dummy = 0; (void)dummy;
          const int s = (j - k + 11) % 11;
#pragma unroll
          for (int m = 0; m < 5; ++m) acc[s][m] = fmaf(w[k], h[m], acc[s][m]);
        }
        const int s = (j + 1) % 11;               // output row o = r - 10 has now seen all eleven taps
        const int o = r - 10;
        if (o >= t.r0) {                          // warp-uniform (rows before the segment were warm-up only)
          const float mu1 = acc[s][0], mu2 = acc[s][1];
          if (ox < Wv) {
            const float s1 = acc[s][2] - mu1 * mu1, s2 = acc[s][3] - mu2 * mu2, s12 = acc[s][4] - mu1 * mu2;
            const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
            sum_s += ((2.f * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs;
            sum_c += cs;
            if (STORE) {
              const size_t plane = (size_t)Hv * Wv, off = (size_t)t.nc * plane + (size_t)o * Wv + ox;
#pragma unroll
              for (int m = 0; m < 5; ++m) mom[(size_t)m * NC * plane + off] = acc[s][m];
            }
          }
        }
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[s][m] = 0.f;
      }
    }
  }
  sum_s = warp_sum(sum_s);
  sum_c = warp_sum(sum_c);
  if (lane == 0) { atomicAdd(sums + 2 * t.nc, sum_s); atomicAdd(sums + 2 * t.nc + 1, sum_c); }
}

// backward w.r.t. Y from the stored moments: dY(q) (=|+=) A(q) + 2 y(q) B(q) + x(q) C(q), (A,B,C) = w (*) (a,b,c) (full correlation)
__global__ void __launch_bounds__(32 * SS_WARPS) k_ssim_bwd3(const float* __restrict__ X, const float* __restrict__ Y,
                                                             const float* __restrict__ mom, int H, int W, float C1, float C2,
                                                             const float* __restrict__ coef, float* __restrict__ dY, int accum,
                                                             int NC, int nstrip, int nseg, int rs, long long ntasks) {
  __shared__ float srow[SS_WARPS][2][3][SS_IN + 2];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hv = H - 10, Wv = W - 10;
  SsimTask t;
  if (!ssim_task((long long)blockIdx.x * SS_WARPS + wid, ntasks, nstrip, nseg, rs, H, t)) return;
  const float gs = coef[2 * t.nc], gc = coef[2 * t.nc + 1];
  const size_t plane = (size_t)Hv * Wv;
  const float* mp = mom + (size_t)t.nc * plane;
  const size_t mstride = (size_t)NC * plane;
  float w[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) w[k] = c_win[k];
  float acc[11][3];
#pragma unroll
  for (int s = 0; s < 11; ++s) acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
  const int qx = t.c0 + lane;
  int buf = 0;
  // adjoint maps at window origin (py, px); zero outside the valid window range
  auto adjoint = [&](int py, int px, float& a, float& b, float& c) {
    a = b = c = 0.f;
    if (py < 0 || py >= Hv || px < 0 || px >= Wv) return;
    const size_t off = (size_t)py * Wv + px;
    const float mu1 = __ldg(mp + off), mu2 = __ldg(mp + mstride + off), exx = __ldg(mp + 2 * mstride + off),
                eyy = __ldg(mp + 3 * mstride + off), exy = __ldg(mp + 4 * mstride + off);
    const float A1 = 2.f * mu1 * mu2 + C1, B1 = mu1 * mu1 + mu2 * mu2 + C1;
    const float A2 = 2.f * (exy - mu1 * mu2) + C2, B2 = (exx - mu1 * mu1) + (eyy - mu2 * mu2) + C2;
    const float iB1 = 1.f / B1, iB2 = 1.f / B2;
    const float cs = A2 * iB2, lum = A1 * iB1, Sv = lum * cs;
    const float as = 2.f * mu1 * (A2 - A1) * iB1 * iB2 + 2.f * mu2 * Sv * (iB2 - iB1);
    const float bs = -Sv * iB2, cS = 2.f * lum * iB2;
    const float ac = (-2.f * mu1 + 2.f * mu2 * cs) * iB2, bc = -cs * iB2, cC = 2.f * iB2;
    a = gs * as + gc * ac; b = gs * bs + gc * bc; c = gs * cS + gc * cC;
  };
  const int p_begin = t.r0 - 10;                   // input (window-origin) rows [r0 - 10, r1): output row q completes at p = q
  for (int base = p_begin; base < t.r1; base += 11) {
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int py = base + j;
      if (py < t.r1) {                             // warp-uniform
        float* ba = srow[wid][buf][0];
        float* bb = srow[wid][buf][1];
        float* bc_ = srow[wid][buf][2];
        buf ^= 1;
        {
          float a, b, c;
          adjoint(py, t.c0 - 10 + lane, a, b, c);
          ba[lane] = a; bb[lane] = b; bc_[lane] = c;
          if (lane < 10) {
            adjoint(py, t.c0 + 22 + lane, a, b, c);
            ba[32 + lane] = a; bb[32 + lane] = b; bc_[32 + lane] = c;
          }
        }
        __syncwarp();
        float h[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {             // output column qx sees origin column qx - k = staged index lane + 10 - k
          h[0] = fmaf(w[k], ba[lane + 10 - k], h[0]);
          h[1] = fmaf(w[k], bb[lane + 10 - k], h[1]);
          h[2] = fmaf(w[k], bc_[lane + 10 - k], h[2]);
        }
#pragma unroll
        for (int k = 0; k < 11; ++k) {             // origin row py is tap k of output row py + k
          const int s = (j + k) % 11;
          acc[s][0] = fmaf(w[k], h[0], acc[s][0]);
          acc[s][1] = fmaf(w[k], h[1], acc[s][1]);
          acc[s][2] = fmaf(w[k], h[2], acc[s][2]);
        }
        const int s = j;                           // output row q = py has now seen origins py - 10 .. py
        if (py >= t.r0 && qx < W) {
          const size_t o = (size_t)t.nc * H * W + (size_t)py * W + qx;
          float g = acc[s][0] + 2.f * __ldg(Y + o) * acc[s][1] + __ldg(X + o) * acc[s][2];
          if (accum) g += dY[o];
          dY[o] = g;
        }
        acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
      }
    }
  }
}
