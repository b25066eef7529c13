// Fused Adam over flat fp32 parameter/gradient/moment buffers (torch.optim.Adam semantics,
// pix2pix_model.py:122-125) with optional bf16 shadow copy, plus the transposed bf16 packer used for
// input-gradient GEMM operands.
#include "common.cuh"
#include "../../include/dsgan_b200.h"
using namespace dsgan;

namespace {
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
                       float bc2_sqrt, float gscale, bf16* __restrict__ pb) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}
// same update with the step-dependent scalars read from device memory: hyper = {lr / (1 - beta1^t), sqrt(1 - beta2^t)}.
// Lets a captured CUDA graph of the training step be replayed while t and the learning-rate schedule advance.
__global__ void k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, long long n, const float* __restrict__ hyper, float b1, float b2, float eps,
                           float gscale, bf16* __restrict__ pb) {
  const float lr_bc1 = hyper[0], bc2_sqrt = hyper[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - lr_bc1 * (mi / denom);
    p[i] = pi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}
__global__ void k_set2(float* dst, float a, float b) { dst[0] = a; dst[1] = b; }
__global__ void k_pack_t(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(long long)c * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}
}  // namespace

extern "C" {
int dsgan_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, int step_t, float grad_scale, void* p_bf16, void* stream) {
  DS_REQUIRE(step_t >= 1, "adam: step_t must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step_t);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step_t));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, grad_scale,
                                                           (bf16*)p_bf16);
  return DS_LAUNCHED("adam_step");
}
int dsgan_adam_hyper(float* hyper, float lr, float beta1, float beta2, int step_t, void* stream) {
  DS_REQUIRE(hyper && step_t >= 1, "adam_hyper: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step_t);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step_t));
  k_set2<<<1, 1, 0, (cudaStream_t)stream>>>(hyper, lr / bc1, bc2);
  return DS_LAUNCHED("adam_hyper");
}
int dsgan_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper, float beta1,
                        float beta2, float eps, float grad_scale, void* p_bf16, void* stream) {
  DS_REQUIRE(hyper, "adam_step_dev: null hyper");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_adam_dev<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, beta1, beta2, eps, grad_scale,
                                                               (bf16*)p_bf16);
  return DS_LAUNCHED("adam_step_dev");
}
int dsgan_pack_transpose_bf16(const float* src, void* dst, int rows, int cols, void* stream) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
  k_pack_t<<<grid, block, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, rows, cols);
  return DS_LAUNCHED("pack_transpose_bf16");
}
}
