/* dsgan_b200 — C ABI of the B200-native (sm_100a) kernels behind DS-GAN's adversarial training step.
 *
 * The reference (yglbgyx/DS-GAN) has no FFI: its extension point is the Python model/network API
 * (DSGAN/models/__init__.py:4-37, networks.py:81-163, MS_SSIM.py:95-225) and every op below that API is
 * a torch library call.  This header is the boundary a maintainer binds instead of those calls; each
 * entry cites the reference op site it replaces (paths under DSGAN/).
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated; `stream` is a
 *    cudaStream_t passed as void*.  The library never allocates or frees device memory and never
 *    synchronises.  Return 0 on success, non-zero on failure (see dsgan_last_error()).
 *  - activations are NHWC ("channels last"): element (n,y,x,c) of a tensor with pixel pitch `ld`
 *    lives at ((n*H+y)*W+x)*ld + c.  `dtype` selects the activation element type: 0 = fp32
 *    (validation mode), 1 = bf16 (production).  Parameters and parameter gradients are always fp32
 *    in the reference's own layouts (OIHW conv, IOHW transposed conv, (out,in) linear).
 *  - image-level tensors exchanged with the host API (real_A, real_B, fake_B, d fake_B) are NCHW fp32,
 *    exactly what the reference's set_input / forward hold (pix2pix_model.py:129-139).
 *  - there is NO CPU fallback: every entry point fails on a non-sm_100 device.
 */
#ifndef DSGAN_B200_H
#define DSGAN_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DSGAN_ABI_VERSION 1

enum { DSGAN_F32 = 0, DSGAN_BF16 = 1 };
enum { DSGAN_ACT_NONE = 0, DSGAN_ACT_RELU = 1, DSGAN_ACT_LEAKY = 2, DSGAN_ACT_GELU = 3, DSGAN_ACT_SIGMOID = 4 };

/* ---- runtime ------------------------------------------------------------------------------ */
int dsgan_abi_version(void);
const char* dsgan_last_error(void);
/* 0 iff the current device is sm_100 (B200); otherwise sets the error string. */
int dsgan_device_check(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long dsgan_launch_count(void);
int dsgan_memset(void* p, int byte_value, size_t bytes, void* stream);

/* ---- layout / elementwise ------------------------------------------------------------------ */
/* dst[n,y,x,c] = src[n,c,y,x]*scale + shift.  Replaces the implicit NCHW->kernel layout of
 * `input['A'].to(device)` (pix2pix_model.py:131-132) and torch.cat((real_A, fake_B),1) (:145,:153,:168)
 * when called twice into channel slices of one NHWC buffer. */
int dsgan_nchw_to_nhwc(const float* src, void* dst, int dtype, int N, int C, int H, int W, int ld_dst,
                       float scale, float shift, void* stream);
/* dst[n,c,y,x] (fp32) = alpha*src[n,y,x,c] (+ dst if accumulate). */
int dsgan_nhwc_to_nchw(const void* src, int dtype, int ld_src, float* dst, int N, int C, int H, int W,
                       float alpha, int accumulate, void* stream);
/* Device-side input pipeline (data/aligned_dataset.py:53-90): uint8 HWC RGB [N,Hs,Ws,3] -> ToTensor -> crop at
 * (h_off[n], w_off[n]) -> Normalize(0.5,0.5) -> flip[n] (horizontal) -> fp32 NCHW [N,C_out,H,W]; C_out = 1 applies the
 * reference's RGB->gray weights.  h_off / w_off / flip: int32 device arrays [N].  Bit-exact with the reference's CPU ops. */
int dsgan_preprocess_u8(const unsigned char* src, int N, int Hs, int Ws, const int* h_off, const int* w_off, const int* flip,
                        float* dst, int C_out, int H, int W, void* stream);
/* dst[p, 0:C] (=|+=) src[p, 0:C] for npix pixels with independent pitches: torch.cat / slicing and their
 * backward (MixConvNeXtML.py:66,110). */
int dsgan_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int dtype, long long npix, int C,
                        int accumulate, void* stream);
/* out = a + b (+ c + d + e): the multi-scale skip sums (MixConvNeXtML.py:482-491). Unused inputs NULL. */
int dsgan_add_n(void* out, int dtype, long long n, const void* a, const void* b, const void* c, const void* d,
                const void* e, void* stream);
/* y[n,p,c] = x[n,p,c] * s[n,c]  (out * CA(out), MixConvNeXtML.py:112);  s is fp32 [N,C]. */
int dsgan_scale_nc_fwd(const void* x, const float* s, void* y, int dtype, int N, long long HW, int C, void* stream);
/* ds[n,c] = sum_p dy*x  (fp32, overwritten). */
int dsgan_scale_nc_bwd_reduce(const void* x, const void* dy, float* ds, int dtype, int N, long long HW, int C,
                              void* stream);
/* dx = dy*s + davg[n,c]/HW + (p == argmax[n,c]) * dmax[n,c]   (=|+=) : backward of out*CA(out) including the
 * adaptive avg/max pooling inside CA (MixConvNeXtML.py:7-8,18-19). */
int dsgan_scale_nc_bwd_apply(const void* dy, const float* s, const float* davg, const float* dmax,
                             const int* argmax, void* dx, int dtype, int N, long long HW, int C, int accumulate,
                             void* stream);

/* ---- dense convolution family (CUDA-core implicit GEMM; fp32 validation + odd shapes) ------ */
typedef struct {
  int dtype;
  int N, Hi, Wi, Ci; /* input  tensor, pixel pitch ld_in  */
  int Ho, Wo, Co;    /* output tensor, pixel pitch ld_out */
  int kh, kw, stride, pad;
  /* 0: out[o]  = sum_k in[o*stride - pad + k] * w   (nn.Conv2d forward; ConvTranspose2d input-gradient)
   * 1: out[o]  = sum_k in[(o + pad - k)/stride] * w (nn.ConvTranspose2d forward; Conv2d input-gradient) */
  int transposed;
  int ld_in, ld_out, ld_aux, ld_pre;
  long long w_sco, w_sci, w_sky, w_skx; /* element strides of the fp32 weight for (out ch, in ch, ky, kx) */
  int act;        /* DSGAN_ACT_* applied last                                   */
  int dact;       /* multiply by act'(aux) before `act` (backward through an activation) */
  int accumulate; /* add the previous contents of `out` before dact/act           */
} dsgan_conv_desc;
/* v = sum + bias (+ out) ; v *= dact'(aux) ; pre_out = v ; out = act(v).
 * Replaces nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward and input-gradient:
 * MixConvNeXtML.py:53,122-159,218,222-224,335-425,459; networks.py:544-569; models/vgg.py:16-25. */
int dsgan_conv_fwd(const dsgan_conv_desc* d, const void* in, const float* w, const float* bias, void* out,
                   void* pre_out, const void* aux, void* stream);
/* dw[co,ci,ky,kx] += sum_{n,oy,ox} dout[n,oy,ox,co] * in[n, oy*stride-pad+ky, ox*stride-pad+kx, ci]
 * (d->N,Hi,Wi,Ci describe `in`; Ho,Wo,Co describe `dout`; `transposed` ignored). Weight gradient of the
 * same op sites. */
int dsgan_conv_wgrad(const dsgan_conv_desc* d, const void* in, const void* dout, float* dw, void* stream);
/* out[c] += sum_p x[p,c]: bias gradients. */
int dsgan_colsum(const void* x, int dtype, int ld, long long npix, int C, float* out, void* stream);

/* ---- tensor-core (tcgen05 + TMEM + TMA) GEMM family: nn.Linear / 1x1 nn.Conv2d in bf16, fp32 accumulate ------
 * Operands are bf16; weights are the bf16 shadow copy of the fp32 masters in the reference's own [out,in] layout. */
/* 1 if the shape/pitches are eligible for dsgan_tc_gemm (mode 0/1) or dsgan_tc_wgrad (mode 2). */
int dsgan_tc_gemm_supported(int mode, long long M, int N, int K, int lda, int ldb, int ldc);
/* mode 0 (forward):        C[M,N] = epi(A[M,K] . W[N,K]^T)   W = weight [out=N, in=K]   (MixConvNeXtML.py:218,222-224)
 * mode 1 (input-gradient): C[M,N] = epi(A[M,K] . W[K,N])     W = the same weight [out=K, in=N], read MN-major
 * epi: v = acc + bias (+ C) ; v *= dact'(aux) ; pre = v ; C = act(v)   — as dsgan_conv_fwd. */
int dsgan_tc_gemm(int mode, const void* A, int lda, const void* W, int ldb, long long M, int N, int K, void* C, int ldc,
                  const float* bias, void* pre, int ld_pre, const void* aux, int ld_aux, int act, int dact,
                  int accumulate, void* stream);
/* weight-gradient: dW[Co,Ci] (fp32, pitch ld_dw) += dY[P,Co]^T . X[P,Ci]; split over P, fp32 atomics. */
int dsgan_tc_wgrad(const void* dY, int ld_dy, const void* X, int ld_x, long long P, int Co, int Ci, float* dW, int ld_dw,
                   void* stream);

/* ---- fused ConvNeXt Block MLP (tcgen05, hidden kept on chip) ---------------------------------------------------
 * Y[M,Nout] = X[M,Cin] . Ws[Nout,Cin]^T  +  GELU( T[M,Cin] . W1[4Cin,Cin]^T + b1 ) . W2[Nout,4Cin]^T + b2
 * = Block.forward after the depthwise conv + norm (MixConvNeXtML.py:236-243: pwconv1 -> GELU -> pwconv2, + shortcut).
 * T, X, Y: bf16 NHWC pixels with pitches ld_*; W1, W2, Ws: bf16 row-major (the networks' bf16 shadow); b1, b2: fp32.
 * X / Ws may both be NULL (no shortcut), b2 may be NULL.  The 4Cin-wide hidden tensor never reaches HBM. */
int dsgan_fused_mlp_supported(int Cin, int Nout);
int dsgan_fused_mlp_fwd(const void* T, int ld_t, const void* X, int ld_x, long long M, int Cin, int Nout, const void* W1,
                        const float* b1, const void* W2, const float* b2, const void* Ws, void* Y, int ld_y,
                        void* stream);
/* Backward of the MLP part (shortcut excluded) with the hidden recomputed:
 *   Hpre = T.W1^T + b1;  A = GELU(Hpre);  G = (dY.W2) * GELU'(Hpre);  dT = G.W1      (MixConvNeXtML.py:236-240 backward)
 * dT: bf16 [M,Cin] (overwritten).  G, A: bf16 [M,4Cin] workspaces, fully written - the operands of the two weight-gradient
 * GEMMs (dW1 = G^T.T, dW2 = dY^T.A via dsgan_tc_wgrad).  db1 (fp32 [4Cin], +=, optional) = column sums of G;
 * db2 (fp32 [Nout], +=, optional) = column sums of dY, taken from the resident dY tile (no separate pass over dY). */
int dsgan_fused_mlp_bwd(const void* T, int ld_t, const void* dY, int ld_dy, long long M, int Cin, int Nout, const void* W1,
                        const float* b1, const void* W2, void* dT, int ld_dt, void* G, void* A, float* db1, float* db2,
                        void* stream);
/* dst (bf16) = src (fp32), n elements: refresh of the packed GEMM operands after an optimizer step. */
int dsgan_pack_bf16(const float* src, void* dst, long long n, void* stream);

/* tensor-core implicit-GEMM convolution (csrc/tc_conv.cu).  Grid position (y,x) of image n reads input pixel
 * (y*in_stride + dy[t], x*in_stride + dx[t]) for tap t (zero outside the image) against weight slab slab[t], and writes
 * output pixel (y*out_stride + oy0[c], x*out_stride + ox0[c]) of parity class c.  w_slabs: bf16 [nslabs][co_pad][ci_pad], zero padded
 * (dsgan_pack_conv_weight).  Any Ci/Co: the input pixel pitch must be a multiple of 8 elements (TMA zero-fills the
 * channels >= Ci), narrow outputs (Co = 1, 3, 6 ...) take a scalar epilogue.
 * Covers nn.Conv2d s1/s2 and nn.ConvTranspose2d(s2) forward and input-gradients: models/vgg.py:16-25,
 * networks.py:544-569, MixConvNeXtML.py:53,150.  Epilogue as dsgan_conv_fwd.
 * Back ends behind this one entry point (chosen by shape, same results): the generic tcgen05 implicit GEMM; a halo-staged
 * tcgen05 variant for 3x3 / stride-1 layers with 64 input and <= 64 output channels (vgg.py:17 conv1_2, MixConvNeXtML.py:459
 * `res`); and a CUDA-core FFMA2 direct convolution (csrc/sc_conv.cu) for layers with 1/3/6/12 input channels or <= 16 output
 * channels (vgg.py:16 conv1_1, networks.py:544 PatchGAN conv1, MixConvNeXtML.py:222-226 block c1).  With a ragged Co (not a
 * multiple of 8) the CUDA-core back end writes whole 16-byte vectors: out / pre_out must then be whole-pitch tensors
 * (ld == ceil8(Co)) whose extra lanes are padding. */
typedef struct {
  int N, Hi, Wi, Ci, ld_in;
  int Ho, Wo, Co, ld_out;
  int Hg, Wg;
  int in_stride, out_stride;
  int nclass;         /* 1, or up to 4 output-parity classes handled by one launch (same grid extent for all) */
  int oy0[4], ox0[4], ntaps[4]; /* per class: output offset and number of taps; taps of class c start at index 16*c */
  int nslabs;
  int co_pad, ci_pad; /* row / column padding of the packed slabs (>= Co, >= Ci and a multiple of 64) */
  int dy[64], dx[64], slab[64];
  int ld_aux, ld_pre, act, dact, accumulate;
} dsgan_tc_conv_desc;
int dsgan_tc_conv_supported(int Ci, int Co, int ld_in, int ld_out);
int dsgan_tc_conv(const dsgan_tc_conv_desc* d, const void* in, const void* w_slabs, const float* bias, void* out,
                  void* pre_out, const void* aux, void* stream);
/* tensor-core weight gradient of the same convolutions:
 *   dW[tap_off[t] + m*s_g + n*s_x] += sum_{img,y,x} G[img,y,x,m] * X[img, y*x_stride + dy[t], x*x_stride + dx[t], n]
 * G: bf16 NHWC on the (Hg,Wg) grid with Cg channels; X: bf16 NHWC (Hx,Wx) with Cx channels; dW: fp32 master-layout
 * gradient addressed with its own strides (OIHW for nn.Conv2d, IOHW for nn.ConvTranspose2d). */
typedef struct {
  int N, Hg, Wg, Cg, ld_g;
  int Hx, Wx, Cx, ld_x;
  int x_stride, ntaps;
  int dy[16], dx[16];
  long long tap_off[16];
  long long s_g, s_x;
} dsgan_tc_wgrad_desc;
int dsgan_tc_conv_wgrad_supported(int Cg, int Cx, int ld_g, int ld_x);
int dsgan_tc_conv_wgrad(const dsgan_tc_wgrad_desc* d, const void* G, const void* X, float* dW, void* stream);
/* dst[slab=ky*kw+kx][o < O_pad][i < I_pad] (bf16) = src[o*s_o + i*s_i + ky'*s_ky + kx'*s_kx] (0 in the padding),
 * (ky',kx') = flipped tap if flip. */
int dsgan_pack_conv_weight(const float* src, void* dst, int O, int I, int O_pad, int I_pad, int kh, int kw,
                           long long s_o, long long s_i, long long s_ky, long long s_kx, int flip, void* stream);

/* The same packing for a whole network in one launch: jobs_dev is a DEVICE array of njobs descriptors (no flip).  Every block
 * packs 1024 consecutive output elements of one job: block0 = first block of the job (ascending), total_blocks = their sum. */
typedef struct {
  unsigned long long src;  /* const float*  : master weight */
  unsigned long long dst;  /* bf16*         : slab buffer [kh*kw][O_pad][I_pad] */
  int O, I, O_pad, I_pad, kh, kw;
  long long s_o, s_i, s_ky, s_kx;
  int block0, pad_;
} dsgan_pack_job;
int dsgan_pack_conv_weights(const dsgan_pack_job* jobs_dev, int njobs, int total_blocks, void* stream);

/* ---- depthwise convolution (MixConvNeXtML.py:94-97,220: k = 3,5,7,9, stride 1, pad k/2) ---- */
/* flip=0: forward (w is [C,1,k,k] fp32, bias may be NULL); flip=1: input-gradient (correlate with the
 * flipped kernel, no bias). */
int dsgan_dwconv_fwd(const void* x, int ld_x, const float* w, const float* bias, void* y, int ld_y, int dtype,
                     int N, int H, int W, int C, int k, int flip, int accumulate, void* stream);
/* dw[c,ky,kx] += sum dy*x_shifted ; db[c] += sum dy  (db may be NULL). */
int dsgan_dwconv_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, float* dw, float* db, int dtype,
                       int N, int H, int W, int C, int k, void* stream);

/* Several depthwise convolutions with different kernel sizes over disjoint channel slices of ONE NHWC tensor in one launch
 * (MidMLKA.X3/X5/X7/X9 on the four quarters of its input, MixConvNeXtML.py:94-97,110).  Branch i maps channels
 * [c0, c0+c) of x to the same channels of y with its own fp32 weight [c,1,k,k] / bias [c] (bias may be NULL); semantics per
 * branch as dsgan_dwconv_fwd / dsgan_dwconv_wgrad (dw +=, db += if not NULL).  nbr <= 4. */
typedef struct {
  const float* w;
  const float* bias;
  float* dw;
  float* db;
  int k, c0, c, pad_;
} dsgan_dw_branch;
int dsgan_dwconv_multi_fwd(const void* x, int ld_x, void* y, int ld_y, int dtype, int N, int H, int W,
                           const dsgan_dw_branch* br, int nbr, int flip, int accumulate, void* stream);
int dsgan_dwconv_multi_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, int dtype, int N, int H, int W,
                             const dsgan_dw_branch* br, int nbr, void* stream);

/* ---- InstanceNorm2d(affine=False, eps=1e-5) fused with activation / residual ---------------- */
/* stats[n,c] = {shift, sum(x-shift), sum((x-shift)^2)} (fp32 [N,C,3], overwritten).  networks.py:25. */
int dsgan_inorm_stats(const void* x, int ld_x, int dtype, int N, long long HW, int C, float* stats, void* stream);
/* y = act( (x-mean)*rstd + res ).  Covers IN, IN+GELU (+concat slice via ld_y), IN+LeakyReLU, and
 * `IN(out) += x; GELU` (MixConvNeXtML.py:54,113-115,186-188,221,335-338; networks.py:556-566). res may be NULL. */
int dsgan_inorm_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, void* y, int ld_y,
                      int dtype, int N, long long HW, int C, int act, void* stream);
/* bstats[n,c] = {sum g, sum g*xhat}, g = dy*act'(xhat+res)  (fp32 [N,C,2], overwritten). */
int dsgan_inorm_bwd_stats(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, int dtype, int N, long long HW, int C, int act, float* bstats, void* stream);
/* dx (=|+=) rstd*(g - mean(g) - xhat*mean(g*xhat));  dres (=|+=) g  (dres may be NULL).
 * dbias (fp32 [C], +=) / dsum_nc (fp32 [N,C], +=), both optional: the sum over pixels of THIS call's dx values in fp32,
 * taken before the bf16 rounding and before any fan-in add.  It is the gradient of a bias feeding the norm (Conv2d /
 * ConvTranspose2d / depthwise conv + InstanceNorm2d: MixConvNeXtML.py:53-54,220-221; networks.py:556-566): zero in exact
 * arithmetic, fp32 round-off in the reference. */
int dsgan_inorm_bwd_apply(const void* x, int ld_x, const float* stats, const void* res, int ld_res, const void* dy,
                          int ld_dy, const float* bstats, void* dx, int ld_dx, int acc_dx, void* dres, int ld_dres,
                          int acc_dres, int dtype, int N, long long HW, int C, int act, float* dbias, float* dsum_nc,
                          void* stream);

/* ---- pooling ------------------------------------------------------------------------------- */
/* nn.MaxPool2d(k) (stride k), MixConvNeXtML.py:68-74,333-354; models/vgg.py pools. */
int dsgan_maxpool_fwd(const void* x, int ld_x, void* y, int ld_y, int dtype, int N, int H, int W, int C, int k,
                      void* stream);
/* dx (=|+=) dy routed to the first maximum in scan order; if relu_mask, the result is then multiplied by (x>0). */
int dsgan_maxpool_bwd(const void* x, int ld_x, const void* dy, int ld_dy, void* dx, int ld_dx, int dtype, int N,
                      int H, int W, int C, int k, int accumulate, int relu_mask, void* stream);
/* MaxPool2d(2), (4), ... (2^nlev) of ONE bf16 NHWC tensor in a single pass (nlev = 2..4), and the combined backward:
 * the multi-scale down-skips of an encoder stage (MixConvNeXtML.py:328-426) plus its own downSample (:68-74).  Outputs /
 * output-gradients are dense (pitch C); a NULL dy means no gradient flowed into that scale.  Same first-maximum routing as
 * nn.MaxPool2d.  dx (=|+=). */
int dsgan_multipool_supported(int dtype, int H, int W, int C, int nlev, int ld_x);
int dsgan_multipool_fwd(const void* x, int ld_x, void* y2, void* y4, void* y8, void* y16, int dtype, int N, int H, int W, int C,
                        int nlev, void* stream);
int dsgan_multipool_bwd(const void* x, int ld_x, const void* dy2, const void* dy4, const void* dy8, const void* dy16, void* dx,
                        int ld_dx, int accumulate, int dtype, int N, int H, int W, int C, int nlev, void* stream);
/* AdaptiveAvgPool2d(1) + AdaptiveMaxPool2d(1) + fc1 -> PReLU -> fc2 (shared) -> add -> sigmoid
 * (CA.forward, MixConvNeXtML.py:17-22).  avg,mx,s: fp32 [N,C]; argmax int32 [N,C] (pixel index);
 * fc1 [C/8,C], fc2 [C,C/8], slope [1]. */
int dsgan_ca_fwd(const void* x, int dtype, int N, long long HW, int C, const float* fc1, const float* slope,
                 const float* fc2, float* avg, float* mx, int* argmax, float* s, void* workspace /* >= 8*N*C bytes */,
                 void* stream);
/* given ds [N,C]: dfc1,dfc2,dslope += ; davg,dmax [N,C] overwritten. */
int dsgan_ca_bwd(const float* ds, const float* s, const float* avg, const float* mx, int N, int C, const float* fc1,
                 const float* slope, const float* fc2, float* dfc1, float* dslope, float* dfc2, float* davg,
                 float* dmax, void* stream);
/* Bias gradients of one MidMLKA from fp32 plane statistics (MixConvNeXtML.py:94-113): conv.bias[c] += sum_n(s*S + davg +
 * dmax), X{3,5,7,9}.bias += W_conv^T . d conv.bias, with S = dsum_nc of dsgan_inorm_bwd_apply.  w_conv: fp32 [C,C]. */
int dsgan_mid_bias_grads(const float* dsum_nc, const float* s, const float* davg, const float* dmax, int N, int C,
                         const float* w_conv, float* d_conv_bias, float* d_b3, float* d_b5, float* d_b7, float* d_b9,
                         void* stream);

/* ---- losses --------------------------------------------------------------------------------- */
/* GANLoss (networks.py:143-163): mode 0 = BCEWithLogits, 1 = MSE, 2 = MSE on sigmoid(pred) (the `--no_lsgan`
 * pairing of Sigmoid-D + MSELoss, pix2pix_model.py:98,112-114); target is the constant 1.0/0.0.
 * loss[0] += loss_scale*mean(...);  dpred (=) grad_scale * d mean/d pred  (dpred may be NULL). */
int dsgan_gan_loss(const void* pred, int dtype, long long n, int ld, float target, int mode, float loss_scale,
                   float* loss, float grad_scale, void* dpred, void* stream); /* element i lives at pred[i*ld] */
/* nn.L1Loss (pix2pix_model.py:115,177,182-186): loss[0] += mean|a-b|; da (=|+=) grad_scale*sign(a-b)/n,
 * multiplied by (a>0) when relu_mask (a is a ReLU output whose gradient is kept w.r.t. its pre-activation). */
int dsgan_l1_loss(const void* a, const void* b, int dtype, long long n, float* loss, float grad_scale, void* da,
                  int accumulate, int relu_mask, void* stream);
/* TV loss (pix2pix_model.py:189-191) on NCHW fp32: loss[0] += (sum|dx|+sum|dy|)/denom; dx (+=) grad_scale*d/dx. */
int dsgan_tv_loss(const float* x, int NC, int H, int W, float denom, float* loss, float grad_scale, float* dx,
                  void* stream);
/* sums[nc] = {sum over the valid windows of ssim_map, of cs_map} (fp32 [NC,2], overwritten) for X, Y: [NC,H,W] fp32 planes;
 * 11-tap sigma 1.5 Gaussian, valid padding (MS_SSIM.py:26-92).  moments (optional, fp32 [5][NC][H-10][W-10]) receives the
 * five filtered moments mu1, mu2, E[xx], E[yy], E[xy] of every window for the backward pass. */
int dsgan_ssim_fwd(const float* X, const float* Y, int NC, int H, int W, float C1, float C2, float* sums, float* moments,
                   void* stream);
/* dY (=|+=) coef[nc,0] * d sum(ssim_map)/dY + coef[nc,1] * d sum(cs_map)/dY  (gradient w.r.t. the second argument only, as
 * the training call needs it, pix2pix_model.py:193-195).  moments = the buffer dsgan_ssim_fwd filled for the same X, Y, or
 * NULL (the moments are then recomputed on a halo: slower).  gnext (optional, needs moments): gradient of the next coarser
 * ms-ssim level [NC, H/2 + H%2, W/2 + W%2]; its avg_pool adjoint (+ gnext/4) is added on the fly (MS_SSIM.py:214-216). */
int dsgan_ssim_bwd(const float* X, const float* Y, int NC, int H, int W, float C1, float C2, const float* coef,
                   const float* moments, float* dY, int accumulate, const float* gnext, void* stream);
/* F.avg_pool2d(k=2) on NCHW fp32 planes (MS_SSIM.py:214-216) and its backward (dx += dy/4). */
int dsgan_avgpool2_fwd(const float* x, float* y, int NC, int H, int W, void* stream);
int dsgan_avgpool2_fwd2(const float* x1, float* y1, const float* x2, float* y2, int NC, int H, int W, void* stream);
int dsgan_avgpool2_bwd(const float* dy, float* dx, int NC, int H, int W, int accumulate, void* stream);
/* ms_ssim combine (MS_SSIM.py:218-225): sums [L,NC,2], sizes[L] = map pixel count, weights[L].
 * val[0] += out_scale * mean_nc prod_l relu(v_l)^w_l ; coef [L,NC,2] = grad_scale * d val / d sums. */
int dsgan_msssim_combine(const float* sums, const float* inv_sizes, const float* weights, int L, int NC,
                         float out_scale, float* val, float grad_scale, float* coef, void* stream);
/* single-scale: val[0] += out_scale*mean_nc(ssim) ; coef[nc] = {grad_scale/(NC*size), 0}. */
int dsgan_ssim_combine(const float* sums, float inv_size, int NC, float out_scale, float* val, float grad_scale,
                       float* coef, void* stream);

/* ---- optimiser (torch.optim.Adam, pix2pix_model.py:122-125) over flat fp32 buffers ---------- */
/* p,g,m,v: flat fp32 [n]; step_t >= 1.  grad_scale multiplies g first (1/world_size after a sum all-reduce).
 * p_bf16 (may be NULL) receives the bf16 copy of the updated parameters. */
int dsgan_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, int step_t, float grad_scale, void* p_bf16, void* stream);
/* Graph-replayable form: the step-dependent scalars live in device memory.  dsgan_adam_hyper writes
 * hyper[0] = lr / (1 - beta1^t), hyper[1] = sqrt(1 - beta2^t) (a by-value kernel argument, so the host may run ahead);
 * dsgan_adam_step_dev is dsgan_adam_step reading them (lr*m_hat/(sqrt(v_hat)+eps) in the same operation order). */
int dsgan_adam_hyper(float* hyper, float lr, float beta1, float beta2, int step_t, void* stream);
int dsgan_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper, float beta1,
                        float beta2, float eps, float grad_scale, void* p_bf16, void* stream);
/* dst bf16 [cols,rows] = transpose(src fp32 [rows,cols]) — packed operand for input-gradient GEMMs. */
int dsgan_pack_transpose_bf16(const float* src, void* dst, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSGAN_B200_H */
