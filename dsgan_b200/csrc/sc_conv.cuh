// CUDA-core backend for small-channel convolutions (sc_conv.cu); tried first by dsgan_tc_conv / dsgan_tc_conv_wgrad.
#pragma once
#include "../../include/dsgan_b200.h"
namespace dsgan {
namespace sc {
// -> true if the shape was taken (then *rc holds the launch status); false: the caller runs the tcgen05 path
bool conv_try(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out, void* pre_out,
              const void* aux, void* stream, int* rc);
bool wgrad_try(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream, int* rc);
}  // namespace sc
}  // namespace dsgan
