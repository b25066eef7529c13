// Layout conversion and small elementwise kernels (HBM-bound; coalesced on the NHWC channel axis).
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

// ---- NCHW fp32 <-> NHWC T --------------------------------------------------------------------
template <typename T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, T* __restrict__ dst, int C, long long HW, int ld,
                               float scale, float shift, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over n*HW pixels
  if (i >= total) return;
  long long n = i / HW, p = i - n * HW;
  const float* s = src + n * C * HW + p;
  T* d = dst + i * ld;
  for (int c = 0; c < C; ++c) stf(d + c, s[(long long)c * HW] * scale + shift);
}
template <typename T>
__global__ void k_nhwc_to_nchw(const T* __restrict__ src, int ld, float* __restrict__ dst, int C, long long HW,
                               float alpha, int acc, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long n = i / HW, p = i - n * HW;
  const T* s = src + i * ld;
  float* d = dst + n * C * HW + p;
  for (int c = 0; c < C; ++c) {
    float v = alpha * ldf(s + c);
    if (acc) v += d[(long long)c * HW];
    d[(long long)c * HW] = v;
  }
}
template <typename T>
__global__ void k_copy_channels(const T* __restrict__ src, int lds, T* __restrict__ dst, int ldd, long long npix,
                                int C, int acc) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = npix * C;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long p = i / C;
    int c = (int)(i - p * C);
    float v = ldf(src + p * lds + c);
    if (acc) v += ldf(dst + p * ldd + c);
    stf(dst + p * ldd + c, v);
  }
}
template <typename T>
__global__ void k_add_n(T* __restrict__ out, long long n, const T* a, const T* b, const T* c, const T* d,
                        const T* e) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = ldf(a + i) + ldf(b + i);
    if (c) v += ldf(c + i);
    if (d) v += ldf(d + i);
    if (e) v += ldf(e + i);
    stf(out + i, v);
  }
}
template <typename T>
__global__ void k_scale_nc_fwd(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y,
                               long long HW, int C, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long n = i / ((long long)HW * C);
    stf(y + i, ldf(x + i) * s[n * C + c]);
  }
}
// grid: (pixel chunks, N); block: 256 threads = (C-lane, pixel-lane)
template <typename T>
__global__ void k_scale_nc_bwd_reduce(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ ds,
                                      long long HW, int C, int chunk) {
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = min(p0 + (long long)chunk, HW);
  const int cl = min(C, (int)blockDim.x);
  const int pl = blockDim.x / cl;
  const int tc = threadIdx.x % cl, tp = threadIdx.x / cl;
  if (tp >= pl) return;
  for (int c = tc; c < C; c += cl) {
    float acc = 0.f;
    for (long long p = p0 + tp; p < p1; p += pl) {
      long long o = ((long long)n * HW + p) * C + c;
      acc += ldf(x + o) * ldf(dy + o);
    }
    atomicAdd(ds + (long long)n * C + c, acc);
  }
}
template <typename T>
__global__ void k_scale_nc_bwd_apply(const T* __restrict__ dy, const float* __restrict__ s,
                                     const float* __restrict__ davg, const float* __restrict__ dmax,
                                     const int* __restrict__ amax, T* __restrict__ dx, long long HW, int C, int acc,
                                     long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float inv = 1.0f / (float)HW;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long np = i / C;
    long long n = np / HW;
    int p = (int)(np - n * HW);
    long long nc = n * C + c;
    float v = ldf(dy + i) * s[nc] + davg[nc] * inv + (amax[nc] == p ? dmax[nc] : 0.f);
    if (acc) v += ldf(dx + i);
    stf(dx + i, v);
  }
}

// bf16 fast paths: 8 channels (16 B) per thread
__device__ __forceinline__ void up8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
    f[2 * q] = t.x; f[2 * q + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pk8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    w[q] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// CA gate (y = x * s[n,c]) and its backward, 8 channels per thread
__global__ void k_scale_nc_fwd_v8(const bf16* __restrict__ x, const float* __restrict__ s, bf16* __restrict__ y, long long HW,
                                  int C, long long total8) {
  const int G = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / G;
    const int c0 = (int)(i - pix * G) * 8;
    const long long n = pix / HW;
    float v[8];
    up8(__ldg(reinterpret_cast<const uint4*>(x + i * 8)), v);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + n * C + c0)), s1 = __ldg(reinterpret_cast<const float4*>(s + n * C + c0 + 4));
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    *reinterpret_cast<uint4*>(y + i * 8) = pk8(v);
  }
}
// ds[n,c] += sum_p x*dy.  grid (pixel chunks, N, channel blocks of <= 256)
__global__ void __launch_bounds__(256) k_scale_nc_bwd_reduce_v8(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                                 float* __restrict__ ds, long long HW, int C, int chunk) {
  __shared__ float sacc[256];
  const int groups = C / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;
  const int tg = threadIdx.x % gl, tp = threadIdx.x / gl;
  const int n = blockIdx.y, cb = blockIdx.z * gl * 8, c0 = cb + tg * 8;
  sacc[threadIdx.x] = 0.f;
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * chunk, p1 = min(p0 + (long long)chunk, HW);
  if (tp < pl && c0 < C) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t base = (size_t)n * HW * C + c0;
#pragma unroll 4
    for (long long p = p0 + tp; p < p1; p += pl) {
      float xv[8], gv[8];
      up8(__ldg(reinterpret_cast<const uint4*>(x + base + p * C)), xv);
      up8(__ldg(reinterpret_cast<const uint4*>(dy + base + p * C)), gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = fmaf(xv[e], gv[e], a[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&sacc[tg * 8 + e], a[e]);
  }
  __syncthreads();
  if (threadIdx.x < gl * 8 && cb + threadIdx.x < C) atomicAdd(ds + (long long)n * C + cb + threadIdx.x, sacc[threadIdx.x]);
}
__global__ void k_scale_nc_bwd_apply_v8(const bf16* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ davg,
                                        const float* __restrict__ dmax, const int* __restrict__ amax, bf16* __restrict__ dx,
                                        long long HW, int C, int acc, long long total8) {
  const int G = C / 8;
  const float inv = 1.0f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / G;
    const int c0 = (int)(i - pix * G) * 8;
    const long long n = pix / HW;
    const int p = (int)(pix - n * HW);
    const long long nc = n * C + c0;
    float g[8], o[8];
    up8(__ldg(reinterpret_cast<const uint4*>(dy + i * 8)), g);
    if (acc) up8(*reinterpret_cast<const uint4*>(dx + i * 8), o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(g[e], __ldg(s + nc + e), __ldg(davg + nc + e) * inv) + (__ldg(amax + nc + e) == p ? __ldg(dmax + nc + e) : 0.f);
      if (acc) v += o[e];
      g[e] = v;
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pk8(g);
  }
}
__global__ void k_copy_channels_v8(const bf16* __restrict__ src, int lds, bf16* __restrict__ dst, int ldd, long long npix,
                                   int G, int acc) {
  const long long total = npix * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / G;
    const int c = (int)(i - p * G) * 8;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + p * lds + c));
    uint4* d = reinterpret_cast<uint4*>(dst + p * ldd + c);
    if (acc) {
      float a[8], b[8];
      up8(v, a); up8(*d, b);
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] += b[q];
      v = pk8(a);
    }
    *d = v;
  }
}
__global__ void k_add_n_v8(bf16* __restrict__ out, long long n8, const bf16* a, const bf16* b, const bf16* c, const bf16* d,
                           const bf16* e) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float s[8], t[8];
    up8(__ldg(reinterpret_cast<const uint4*>(a) + i), s);
    up8(__ldg(reinterpret_cast<const uint4*>(b) + i), t);
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] += t[q];
    if (c) { up8(__ldg(reinterpret_cast<const uint4*>(c) + i), t);
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] += t[q]; }
    if (d) { up8(__ldg(reinterpret_cast<const uint4*>(d) + i), t);
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] += t[q]; }
    if (e) { up8(__ldg(reinterpret_cast<const uint4*>(e) + i), t);
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] += t[q]; }
    reinterpret_cast<uint4*>(out)[i] = pk8(s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Input pipeline on the device (data/aligned_dataset.py:53-90): decoded uint8 RGB images (H x W x 3, what PIL hands over)
// -> transforms.ToTensor (x / 255) -> crop (h_off, w_off) -> Normalize(0.5, 0.5) -> horizontal flip -> optional RGB->gray,
// written as the fp32 NCHW batch the model's set_input expects.  Same op order as the reference, so the result is bit-exact
// with its CPU transforms; the host ships 1 byte per value instead of 4.
__global__ void k_preprocess_u8(const unsigned char* __restrict__ src, int Hs, int Ws, const int* __restrict__ h_off,
                                const int* __restrict__ w_off, const int* __restrict__ flip, float* __restrict__ dst,
                                int C_out, int H, int W, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    const int sx = (flip[n] ? (W - 1 - x) : x) + w_off[n], sy = y + h_off[n];
    const unsigned char* p = src + (((long long)n * Hs + sy) * Ws + sx) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p[c], 255.0f), 0.5f), 0.5f);
    float* o = dst + ((long long)n * C_out * H + y) * W + x;
    if (C_out == 3) {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[(long long)c * H * W] = v[c];
    } else {   // tmp = A[0]*0.299 + A[1]*0.587 + A[2]*0.114 (left to right, unfused like the reference's tensor ops)
      o[0] = __fadd_rn(__fadd_rn(__fmul_rn(v[0], 0.299f), __fmul_rn(v[1], 0.587f)), __fmul_rn(v[2], 0.114f));
    }
  }
}

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" {
int dsgan_nchw_to_nhwc(const float* src, void* dst, int dtype, int N, int C, int H, int W, int ld_dst, float scale,
                       float shift, void* stream) {
  long long HW = (long long)H * W, total = HW * N;
  DS_DISPATCH_DT(dtype, (k_nchw_to_nhwc<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            src, (T*)dst, C, HW, ld_dst, scale, shift, total)));
  return DS_LAUNCHED("nchw_to_nhwc");
}
int dsgan_nhwc_to_nchw(const void* src, int dtype, int ld_src, float* dst, int N, int C, int H, int W, float alpha,
                       int accumulate, void* stream) {
  long long HW = (long long)H * W, total = HW * N;
  DS_DISPATCH_DT(dtype, (k_nhwc_to_nchw<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)src, ld_src, dst, C, HW, alpha, accumulate, total)));
  return DS_LAUNCHED("nhwc_to_nchw");
}
int dsgan_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int dtype, long long npix, int C,
                        int accumulate, void* stream) {
  if (dtype == DT_BF16 && C % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    k_copy_channels_v8<<<grid_for(npix * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, (bf16*)dst,
                                                                                        ld_dst, npix, C / 8, accumulate);
    return DS_LAUNCHED("copy_channels_v8");
  }
  DS_DISPATCH_DT(dtype, (k_copy_channels<T><<<grid_for(npix * C, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)src, ld_src, (T*)dst, ld_dst, npix, C, accumulate)));
  return DS_LAUNCHED("copy_channels");
}
int dsgan_add_n(void* out, int dtype, long long n, const void* a, const void* b, const void* c, const void* d,
                const void* e, void* stream) {
  DS_REQUIRE(a && b, "add_n needs at least two inputs");
  if (dtype == DT_BF16 && n % 8 == 0 &&
      (((uintptr_t)out | (uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e) % 16 == 0)) {
    k_add_n_v8<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((bf16*)out, n / 8, (const bf16*)a, (const bf16*)b,
                                                                      (const bf16*)c, (const bf16*)d, (const bf16*)e);
    return DS_LAUNCHED("add_n_v8");
  }
  DS_DISPATCH_DT(dtype, (k_add_n<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
                            (T*)out, n, (const T*)a, (const T*)b, (const T*)c, (const T*)d, (const T*)e)));
  return DS_LAUNCHED("add_n");
}
int dsgan_scale_nc_fwd(const void* x, const float* s, void* y, int dtype, int N, long long HW, int C, void* stream) {
  long long total = (long long)N * HW * C;
  if (dtype == DT_BF16 && C % 8 == 0 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)s) % 16 == 0)) {
    k_scale_nc_fwd_v8<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, s, (bf16*)y, HW, C, total / 8);
    return DS_LAUNCHED("scale_nc_fwd_v8");
  }
  DS_DISPATCH_DT(dtype, (k_scale_nc_fwd<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, s, (T*)y, HW, C, total)));
  return DS_LAUNCHED("scale_nc_fwd");
}
int dsgan_scale_nc_bwd_reduce(const void* x, const void* dy, float* ds, int dtype, int N, long long HW, int C,
                              void* stream) {
  cudaMemsetAsync(ds, 0, sizeof(float) * N * C, (cudaStream_t)stream);
  if (dtype == DT_BF16 && C % 8 == 0 && (((uintptr_t)x | (uintptr_t)dy) % 16 == 0)) {
    const int groups = C / 8, gl = groups < 32 ? groups : 32, pl = 256 / gl;
    const int cblocks = (C + gl * 8 - 1) / (gl * 8);
    int ch = 2048;
    while (ch > 4 * pl && ((HW + ch - 1) / ch) * N * cblocks < 296) ch >>= 1;
    dim3 g8(cdiv(HW, ch), N, cblocks);
    k_scale_nc_bwd_reduce_v8<<<g8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, ds, HW, C, ch);
    return DS_LAUNCHED("scale_nc_bwd_reduce_v8");
  }
  int chunk = 256;
  dim3 grid(cdiv(HW, chunk), N);
  DS_DISPATCH_DT(dtype, (k_scale_nc_bwd_reduce<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)dy,
                                                                                         ds, HW, C, chunk)));
  return DS_LAUNCHED("scale_nc_bwd_reduce");
}
int dsgan_scale_nc_bwd_apply(const void* dy, const float* s, const float* davg, const float* dmax, const int* argmax,
                             void* dx, int dtype, int N, long long HW, int C, int accumulate, void* stream) {
  long long total = (long long)N * HW * C;
  if (dtype == DT_BF16 && C % 8 == 0 && (((uintptr_t)dy | (uintptr_t)dx) % 16 == 0)) {
    k_scale_nc_bwd_apply_v8<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)dy, s, davg, dmax, argmax, (bf16*)dx, HW, C, accumulate, total / 8);
    return DS_LAUNCHED("scale_nc_bwd_apply_v8");
  }
  DS_DISPATCH_DT(dtype, (k_scale_nc_bwd_apply<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)dy, s, davg, dmax, argmax, (T*)dx, HW, C, accumulate, total)));
  return DS_LAUNCHED("scale_nc_bwd_apply");
}
int dsgan_preprocess_u8(const unsigned char* src, int N, int Hs, int Ws, const int* h_off, const int* w_off, const int* flip,
                        float* dst, int C_out, int H, int W, void* stream) {
  DS_REQUIRE(C_out == 3 || C_out == 1, "preprocess_u8: C_out must be 3 or 1, got %d", C_out);
  DS_REQUIRE(H >= 1 && W >= 1 && H <= Hs && W <= Ws, "preprocess_u8: crop %dx%d does not fit the %dx%d source", H, W, Hs, Ws);
  const long long total = (long long)N * H * W;
  k_preprocess_u8<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, h_off, w_off, flip, dst, C_out, H, W,
                                                                         total);
  return DS_LAUNCHED("preprocess_u8");
}
}
