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
__device__ __forceinline__ float ss_rcp(float v) { return __fdividef(1.0f, v); }   // MUFU.RCP: ~1 ulp, far inside the 5e-5 / 5e-4 bars
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
  // acc[k] = partial sums of output row (r - k) while input row r is being processed (the row loop is NOT unrolled: the
  // window moves by writing every update one slot further, acc[k+1] = acc[k] + w[k] h -- no register copies, a small body)
  float acc[11][5];
#pragma unroll
  for (int k = 0; k < 11; ++k)
#pragma unroll
    for (int m = 0; m < 5; ++m) acc[k][m] = 0.f;
  float sum_s = 0.f, sum_c = 0.f;
  const int ox = t.c0 + lane;
  const int c_a = t.c0 + lane, c_b = t.c0 + 32 + lane;
  const bool ok_a = c_a < W, ok_b = lane < 10 && c_b < W;
  const int r_end = t.r1 + 10;                    // input rows [r0, r1 + 10)
  // the next row's global loads are in flight while the current row is filtered
  float xa = ok_a ? __ldg(xp + (size_t)t.r0 * W + c_a) : 0.f, ya = ok_a ? __ldg(yp + (size_t)t.r0 * W + c_a) : 0.f;
  float xb = ok_b ? __ldg(xp + (size_t)t.r0 * W + c_b) : 0.f, yb = ok_b ? __ldg(yp + (size_t)t.r0 * W + c_b) : 0.f;
  int buf = 0;
#pragma unroll 1
  for (int r = t.r0; r < r_end; ++r) {
    float* bx = srow[wid][buf][0];
    float* by = srow[wid][buf][1];
    buf ^= 1;
    bx[lane] = xa; by[lane] = ya;
    if (lane < 10) { bx[32 + lane] = xb; by[32 + lane] = yb; }
    __syncwarp();
    if (r + 1 < r_end) {
      const size_t ro = (size_t)(r + 1) * W;
      xa = ok_a ? __ldg(xp + ro + c_a) : 0.f; ya = ok_a ? __ldg(yp + ro + c_a) : 0.f;
      xb = ok_b ? __ldg(xp + ro + c_b) : 0.f; yb = ok_b ? __ldg(yp + ro + c_b) : 0.f;
    }
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
    // input row r is tap k of output row r - k; row r - 10 is complete after this update
    float done[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) done[m] = fmaf(w[10], h[m], acc[10][m]);
#pragma unroll
    for (int k = 9; k >= 0; --k)
#pragma unroll
      for (int m = 0; m < 5; ++m) acc[k + 1][m] = fmaf(w[k], h[m], acc[k][m]);   // acc[0] is always zero
    const int o = r - 10;
    if (o >= t.r0 && ox < Wv) {                    // rows before the segment were warm-up only
      const float mu1 = done[0], mu2 = done[1];
      const float s1 = done[2] - mu1 * mu1, s2 = done[3] - mu2 * mu2, s12 = done[4] - mu1 * mu2;
      const float cs = (2.f * s12 + C2) * ss_rcp(s1 + s2 + C2);
      sum_s += ((2.f * mu1 * mu2 + C1) * ss_rcp(mu1 * mu1 + mu2 * mu2 + C1)) * cs;
      sum_c += cs;
      if (STORE) {
        const size_t plane = (size_t)Hv * Wv, off = (size_t)t.nc * plane + (size_t)o * Wv + ox;
#pragma unroll
        for (int m = 0; m < 5; ++m) mom[(size_t)m * NC * plane + off] = done[m];
      }
    }
  }
  sum_s = warp_sum(sum_s);
  sum_c = warp_sum(sum_c);
  if (lane == 0) { atomicAdd(sums + 2 * t.nc, sum_s); atomicAdd(sums + 2 * t.nc + 1, sum_c); }
}

// backward w.r.t. Y from the stored moments: dY(q) (=|+=) A(q) + 2 y(q) B(q) + x(q) C(q), (A,B,C) = w (*) (a,b,c) (full correlation),
// plus (ms-ssim) the adjoint of the avg_pool that produced the next coarser level: + 0.25 gnext[(qy + ph) / 2][(qx + pw) / 2].
struct SsRaw { float v[5]; };
__global__ void __launch_bounds__(32 * SS_WARPS) k_ssim_bwd3(const float* __restrict__ X, const float* __restrict__ Y,
                                                             const float* __restrict__ mom, int H, int W, float C1, float C2,
                                                             const float* __restrict__ coef, float* __restrict__ dY, int accum,
                                                             const float* __restrict__ gnext, int NC, int nstrip, int nseg,
                                                             int rs, long long ntasks) {
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
  // acc[d] = partial sums of output row (py + d) while origin row py is being processed
  float acc[11][3];
#pragma unroll
  for (int d = 0; d < 11; ++d) acc[d][0] = acc[d][1] = acc[d][2] = 0.f;
  const int qx = t.c0 + lane;
  const int px0 = t.c0 - 10 + lane, px1 = t.c0 + 22 + lane;
  const bool col0 = px0 >= 0 && px0 < Wv, col1 = lane < 10 && px1 >= 0 && px1 < Wv;
  // the five moments of a window origin: loads only (issued a row ahead), the pointwise math runs after the filter passes
  auto fetch = [&](int py, int px, bool colok, SsRaw& r) {
    const bool ok = colok && py >= 0 && py < Hv;
    const size_t off = ok ? (size_t)py * Wv + px : 0;
#pragma unroll
    for (int m = 0; m < 5; ++m) r.v[m] = ok ? __ldg(mp + m * mstride + off) : 0.f;
    if (!ok) r.v[0] = __int_as_float(0x7fc00000);      // NaN marks "outside the valid window range"
  };
  auto adjoint = [&](const SsRaw& r, float& a, float& b, float& c) {
    const float mu1 = r.v[0], mu2 = r.v[1], exx = r.v[2], eyy = r.v[3], exy = r.v[4];
    if (mu1 != mu1) { a = b = c = 0.f; return; }
    const float A1 = 2.f * mu1 * mu2 + C1, B1 = mu1 * mu1 + mu2 * mu2 + C1;
    const float A2 = 2.f * (exy - mu1 * mu2) + C2, B2 = (exx - mu1 * mu1) + (eyy - mu2 * mu2) + C2;
    const float iB1 = ss_rcp(B1), iB2 = ss_rcp(B2);
    const float cs = A2 * iB2, lum = A1 * iB1, Sv = lum * cs;
    const float as = 2.f * mu1 * (A2 - A1) * iB1 * iB2 + 2.f * mu2 * Sv * (iB2 - iB1);
    const float bs = -Sv * iB2, cS = 2.f * lum * iB2;
    const float ac = (-2.f * mu1 + 2.f * mu2 * cs) * iB2, bc = -cs * iB2, cC = 2.f * iB2;
    a = gs * as + gc * ac; b = gs * bs + gc * bc; c = gs * cS + gc * cC;
  };
  const int ph = H & 1, pw = W & 1, Hn = H / 2 + ph, Wn = W / 2 + pw;   // geometry of the next coarser level (avg_pool, MS_SSIM.py:214)
  SsRaw r0, r1;
  fetch(t.r0 - 10, px0, col0, r0);
  fetch(t.r0 - 10, px1, col1, r1);
  int buf = 0;
  // origin rows [r0 - 10, r1): output row q completes at py = q
#pragma unroll 1
  for (int py = t.r0 - 10; py < t.r1; ++py) {
    float* ba = srow[wid][buf][0];
    float* bb = srow[wid][buf][1];
    float* bc_ = srow[wid][buf][2];
    buf ^= 1;
    {
      float a, b, c;
      adjoint(r0, a, b, c);
      ba[lane] = a; bb[lane] = b; bc_[lane] = c;
      if (lane < 10) {
        adjoint(r1, a, b, c);
        ba[32 + lane] = a; bb[32 + lane] = b; bc_[32 + lane] = c;
      }
    }
    __syncwarp();
    if (py + 1 < t.r1) {                       // next row's moments: in flight during the filter passes below
      fetch(py + 1, px0, col0, r0);
      fetch(py + 1, px1, col1, r1);
    }
    const bool emit = py >= t.r0 && qx < W;
    float xv = 0.f, yv = 0.f, old = 0.f, gn = 0.f;
    const size_t o = (size_t)t.nc * H * W + (size_t)(emit ? py : 0) * W + (emit ? qx : 0);
    if (emit) {
      xv = __ldg(X + o); yv = __ldg(Y + o);
      if (accum) old = dY[o];
      if (gnext) gn = __ldg(gnext + ((size_t)t.nc * Hn + (py + ph) / 2) * Wn + (qx + pw) / 2);
    }
    float h[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {             // output column qx sees origin column qx - k = staged index lane + 10 - k
      h[0] = fmaf(w[k], ba[lane + 10 - k], h[0]);
      h[1] = fmaf(w[k], bb[lane + 10 - k], h[1]);
      h[2] = fmaf(w[k], bc_[lane + 10 - k], h[2]);
    }
    // origin row py is tap d of output row py + d; row py is complete after this update, the window moves down one slot
    const float A = fmaf(w[0], h[0], acc[0][0]), B = fmaf(w[0], h[1], acc[0][1]), Cc = fmaf(w[0], h[2], acc[0][2]);
#pragma unroll
    for (int d = 1; d < 11; ++d) {
      acc[d - 1][0] = fmaf(w[d], h[0], acc[d][0]);
      acc[d - 1][1] = fmaf(w[d], h[1], acc[d][1]);
      acc[d - 1][2] = fmaf(w[d], h[2], acc[d][2]);
    }
    acc[10][0] = acc[10][1] = acc[10][2] = 0.f;
    if (emit) dY[o] = A + 2.f * yv * B + xv * Cc + old + 0.25f * gn;
  }
}
