// Register-tiled SSIM kernels (included by loss.cu inside its anonymous namespace, after c_win).
// Each work item produces FOUR adjacent outputs of an 11-tap pass from 14 inputs held in registers, and the five
// Gaussian moments travel as two float2 pairs + one scalar so the FMAs issue as packed FFMA2: ~130 instructions per
// pixel instead of ~410 for the one-output-per-thread version (the kernels are FMA-issue bound, not HBM bound:
// 110 FMAs per pixel are intrinsic to a separable 11-tap filter on five moments).
constexpr int V_TS = 32, V_HALO = 10;

__device__ __forceinline__ void ld16(const float* p, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void ld14x2(const float2* p, float2 (&v)[14]) {
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(p + 2 * q);
    v[2 * q] = make_float2(t.x, t.y); v[2 * q + 1] = make_float2(t.z, t.w);
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
constexpr int VF_IN = V_TS + V_HALO;  // 42 input rows / cols
constexpr int VF_P = 44;              // padded pitch (16-byte aligned rows, readable up to column 43)

__global__ void __launch_bounds__(256, 4) k_ssim_fwd2(const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                                                    float C1, float C2, float* __restrict__ sums) {
  __shared__ __align__(16) float sX[VF_IN][VF_P], sY[VF_IN][VF_P];
  __shared__ __align__(16) float2 sM[VF_IN][V_TS], sQ[VF_IN][V_TS];  // (mu1, mu2), (E xx, E yy) after the horizontal pass
  __shared__ __align__(16) float sXY[VF_IN][V_TS];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int nc = blockIdx.z, y0 = blockIdx.y * V_TS, x0 = blockIdx.x * V_TS;
  const float* xp = X + (size_t)nc * H * W;
  const float* yp = Y + (size_t)nc * H * W;
  for (int i = tid; i < VF_IN * VF_P; i += 256) {
    const int r = i / VF_P, c = i % VF_P, gy = y0 + r, gx = x0 + c;
    const bool ok = c < VF_IN && gy < H && gx < W;
    sX[r][c] = ok ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
    sY[r][c] = ok ? __ldg(yp + (size_t)gy * W + gx) : 0.f;
  }
  float w[11];
  float2 w2[11];
#pragma unroll
  for (int t = 0; t < 11; ++t) { w[t] = c_win[t]; w2[t] = make_float2(w[t], w[t]); }
  __syncthreads();
  for (int item = tid; item < VF_IN * (V_TS / 4); item += 256) {
    const int r = item / (V_TS / 4), c0 = (item % (V_TS / 4)) * 4;
    float a[16], b[16];
    ld16(&sX[r][c0], a);
    ld16(&sY[r][c0], b);
    float2 ab2[14], sq2[14];
    float xy[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) {
      ab2[j] = make_float2(a[j], b[j]);
      sq2[j] = __fmul2_rn(ab2[j], ab2[j]);
      xy[j] = a[j] * b[j];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 m = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) {
        m = __ffma2_rn(w2[t], ab2[o + t], m);
        q = __ffma2_rn(w2[t], sq2[o + t], q);
        s = fmaf(w[t], xy[o + t], s);
      }
      sM[r][c0 + o] = m; sQ[r][c0 + o] = q; sXY[r][c0 + o] = s;
    }
  }
  __syncthreads();
  float acc_s = 0.f, acc_c = 0.f;
  {
    const int c = tid & 31, r0 = (tid >> 5) * 4;
    float2 m[14], q[14];
    float xy[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) { m[j] = sM[r0 + j][c]; q[j] = sQ[r0 + j][c]; xy[j] = sXY[r0 + j][c]; }
    const int Hv = H - V_HALO, Wv = W - V_HALO;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 mm = make_float2(0.f, 0.f), qq = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) {
        mm = __ffma2_rn(w2[t], m[o + t], mm);
        qq = __ffma2_rn(w2[t], q[o + t], qq);
        s = fmaf(w[t], xy[o + t], s);
      }
      if (y0 + r0 + o < Hv && x0 + c < Wv) {
        const float mu1 = mm.x, mu2 = mm.y;
        const float s1 = qq.x - mu1 * mu1, s2 = qq.y - mu2 * mu2, s12 = s - mu1 * mu2;
        const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
        acc_s += ((2.f * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs;
        acc_c += cs;
      }
    }
  }
  acc_s = block_sum(acc_s, red);
  acc_c = block_sum(acc_c, red);
  if (tid == 0) { atomicAdd(sums + 2 * nc, acc_s); atomicAdd(sums + 2 * nc + 1, acc_c); }
}

// ---- backward w.r.t. Y ----------------------------------------------------------------------------------------------
constexpr int VB_IN = V_TS + 2 * V_HALO;  // 52 input rows / cols
constexpr int VB_P = 56;                  // padded input pitch
constexpr int VB_MC = 44, VB_MR = 56;     // moment maps: 44 columns (42 used), 56 rows (52 filled)
constexpr int VB_AC = 48, VB_AR = 44;     // adjoint maps a,b,c: 42x42 used

struct Bwd2Smem {
  float X[VB_IN][VB_P], Y[VB_IN][VB_P];
  float2 M[VB_MR][VB_MC], Q[VB_MR][VB_MC];   // later reused for the horizontally filtered adjoints
  float XY[VB_MR][VB_MC];
  float2 AB[VB_AR][VB_AC];
  float Cm[VB_AR][VB_AC];
};

__global__ void __launch_bounds__(256) k_ssim_bwd2(const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                                                    float C1, float C2, const float* __restrict__ coef,
                                                    float* __restrict__ dY, int accum) {
  extern __shared__ __align__(16) unsigned char smraw[];
  Bwd2Smem& S = *reinterpret_cast<Bwd2Smem*>(smraw);
  const int tid = threadIdx.x;
  const int nc = blockIdx.z, qy0 = blockIdx.y * V_TS, qx0 = blockIdx.x * V_TS;
  const float gs = coef[2 * nc], gc = coef[2 * nc + 1];
  const float* xp = X + (size_t)nc * H * W;
  const float* yp = Y + (size_t)nc * H * W;
  for (int i = tid; i < VB_IN * VB_P; i += 256) {
    const int r = i / VB_P, c = i % VB_P, gy = qy0 - V_HALO + r, gx = qx0 - V_HALO + c;
    const bool ok = c < VB_IN && gy >= 0 && gy < H && gx >= 0 && gx < W;
    S.X[r][c] = ok ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
    S.Y[r][c] = ok ? __ldg(yp + (size_t)gy * W + gx) : 0.f;
  }
  float w[11];
  float2 w2[11];
#pragma unroll
  for (int t = 0; t < 11; ++t) { w[t] = c_win[t]; w2[t] = make_float2(w[t], w[t]); }
  __syncthreads();
  // 1) horizontal moments on 52 rows x 44 columns
  for (int item = tid; item < VB_IN * (VB_MC / 4); item += 256) {
    const int r = item / (VB_MC / 4), c0 = (item % (VB_MC / 4)) * 4;
    float a[16], b[16];
    ld16(&S.X[r][c0], a);
    ld16(&S.Y[r][c0], b);
    float2 ab2[14], sq2[14];
    float xy[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) { ab2[j] = make_float2(a[j], b[j]); sq2[j] = __fmul2_rn(ab2[j], ab2[j]); xy[j] = a[j] * b[j]; }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 m = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) { m = __ffma2_rn(w2[t], ab2[o + t], m); q = __ffma2_rn(w2[t], sq2[o + t], q); s = fmaf(w[t], xy[o + t], s); }
      S.M[r][c0 + o] = m; S.Q[r][c0 + o] = q; S.XY[r][c0 + o] = s;
    }
  }
  __syncthreads();
  // 2) vertical moments at the 42x42 window origins p -> adjoint maps a (d/d mu2), b (d/d Eyy), c (d/d Exy)
  const int Hv = H - V_HALO, Wv = W - V_HALO;
  for (int item = tid; item < VB_MC * (VB_AR / 4); item += 256) {
    const int c = item % VB_MC, r0 = (item / VB_MC) * 4;
    float2 m[14], q[14];
    float xy[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) { m[j] = S.M[r0 + j][c]; q[j] = S.Q[r0 + j][c]; xy[j] = S.XY[r0 + j][c]; }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 mm = make_float2(0.f, 0.f), qq = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) { mm = __ffma2_rn(w2[t], m[o + t], mm); qq = __ffma2_rn(w2[t], q[o + t], qq); s = fmaf(w[t], xy[o + t], s); }
      const int r = r0 + o, py = qy0 - V_HALO + r, px = qx0 - V_HALO + c;
      float a = 0.f, b = 0.f, cc = 0.f;
      if (r < VF_IN && c < VF_IN && py >= 0 && py < Hv && px >= 0 && px < Wv) {
        const float mu1 = mm.x, mu2 = mm.y;
        const float A1 = 2.f * mu1 * mu2 + C1, B1 = mu1 * mu1 + mu2 * mu2 + C1;
        const float A2 = 2.f * (s - mu1 * mu2) + C2, B2 = (qq.x - mu1 * mu1) + (qq.y - mu2 * mu2) + C2;
        const float iB1 = 1.f / B1, iB2 = 1.f / B2;
        const float cs = A2 * iB2, lum = A1 * iB1, Sv = lum * cs;
        const float as = 2.f * mu1 * (A2 - A1) * iB1 * iB2 + 2.f * mu2 * Sv * (iB2 - iB1);
        const float bs = -Sv * iB2, cS = 2.f * lum * iB2;
        const float ac = (-2.f * mu1 + 2.f * mu2 * cs) * iB2, bc = -cs * iB2, cC = 2.f * iB2;
        a = gs * as + gc * ac; b = gs * bs + gc * bc; cc = gs * cS + gc * cC;
      }
      S.AB[r][c] = make_float2(a, b);
      S.Cm[r][c] = cc;
    }
  }
  __syncthreads();
  // 3) horizontal pass of the (symmetric) transposed filter over a,b,c: 42 rows x 32 columns, aliased onto M / XY
  float2 (*hAB)[V_TS] = reinterpret_cast<float2 (*)[V_TS]>(&S.M[0][0]);
  float (*hC)[V_TS] = reinterpret_cast<float (*)[V_TS]>(&S.XY[0][0]);
  for (int item = tid; item < VF_IN * (V_TS / 4); item += 256) {
    const int r = item / (V_TS / 4), v0 = (item % (V_TS / 4)) * 4;
    float2 ab[14];
    float cc[16];
    ld14x2(&S.AB[r][v0], ab);
    ld16(&S.Cm[r][v0], cc);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 s2 = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) { s2 = __ffma2_rn(w2[t], ab[o + t], s2); s = fmaf(w[t], cc[o + t], s); }
      hAB[r][v0 + o] = s2; hC[r][v0 + o] = s;
    }
  }
  __syncthreads();
  // 4) vertical pass + chain rule: dY(q) = A + 2 y(q) B + x(q) C
  {
    const int v = tid & 31, u0 = (tid >> 5) * 4;
    float2 ab[14];
    float cc[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) { ab[j] = hAB[u0 + j][v]; cc[j] = hC[u0 + j][v]; }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float2 s2 = make_float2(0.f, 0.f);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) { s2 = __ffma2_rn(w2[t], ab[o + t], s2); s = fmaf(w[t], cc[o + t], s); }
      const int u = u0 + o, gy = qy0 + u, gx = qx0 + v;
      if (gy < H && gx < W) {
        float g = s2.x + 2.f * S.Y[u + V_HALO][v + V_HALO] * s2.y + S.X[u + V_HALO][v + V_HALO] * s;
        float* op = dY + (size_t)nc * H * W + (size_t)gy * W + gx;
        if (accum) g += *op;
        *op = g;
      }
    }
  }
}
