// Warp-level tensor-core backend for narrow-input convolutions (nm_conv.cu); tried before the CUDA-core sc_conv backend.
#pragma once
#include "../../include/dsgan_b200.h"
namespace dsgan {
namespace nm {
// -> true if the shape was taken (then *rc holds the launch status); false: the caller tries the next backend
bool conv_try(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out, void* pre_out,
              const void* aux, void* stream, int* rc);
bool wgrad_try(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream, int* rc);
}  // namespace nm
}  // namespace dsgan
