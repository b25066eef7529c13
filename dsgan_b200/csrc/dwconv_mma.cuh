// Tensor-core backend for the depthwise convolutions (dwconv_mma.cu); tried first by dsgan_dwconv_fwd / dsgan_dwconv_wgrad.
#pragma once
#include "common.cuh"
#include "../../include/dsgan_b200.h"
namespace dsgan {
namespace dwm {
// -> true if the shape was taken (then *rc holds the launch status); false: the caller runs the CUDA-core kernels
bool fwd_try(const bf16* x, int ldx, const float* w, const float* bias, bf16* y, int ldy, int N, int H, int W, int C, int k,
             int flip, int accumulate, cudaStream_t s, int* rc);
bool wgrad_try(const bf16* x, int ldx, const bf16* dy, int lddy, float* dw, float* db, int N, int H, int W, int C, int k,
               cudaStream_t s, int* rc);
bool multi_fwd_try(const bf16* x, int ldx, bf16* y, int ldy, int N, int H, int W, const dsgan_dw_branch* br, int nbr, int flip,
                   int accumulate, cudaStream_t s, int* rc);
bool multi_wgrad_try(const bf16* x, int ldx, const bf16* dy, int lddy, int N, int H, int W, const dsgan_dw_branch* br, int nbr,
                     cudaStream_t s, int* rc);
}  // namespace dwm
}  // namespace dsgan
