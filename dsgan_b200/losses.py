"""Loss graphs: GAN (BCE-with-logits / MSE), L1, VGG-feature L1, TV, SSIM and MS-SSIM — forward value and the
gradient w.r.t. the generated image, written by the bandwidth kernels of csrc/loss.cu.

Image-level tensors are NCHW fp32 (what the reference's API holds, pix2pix_model.py:129-139); loss values are
accumulated into fp32 device scalars (`slot` pointers) with no host synchronisation.
"""
from __future__ import annotations

import torch

from .engine import ACT_RELU, ACT_SIGMOID, F32, Ctx, Var

MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # MS_SSIM.py:200
_CONST = {}   # (device, sizes) -> (1/size per level, weights): built once, so the call never touches the host again
_WORK = {}    # (device, shape, levels, want_grad) -> reusable pyramid / statistics / gradient workspaces


def _consts(device, sizes):
    key = (str(device), tuple(sizes))
    if key not in _CONST:
        _CONST[key] = (torch.tensor([1.0 / s for s in sizes], dtype=torch.float32, device=device),
                       torch.tensor(MS_WEIGHTS, dtype=torch.float32, device=device))
    return _CONST[key]


def _workspace(X, levels, want):
    """Pyramid levels, per-plane sums, coefficients and per-level gradient buffers, allocated once per shape (they are
    fully overwritten by every call)."""
    N, C, H, W = X.shape
    key = (str(X.device), N, C, H, W, levels, want)
    ws = _WORK.get(key)
    if ws is None:
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=X.device)
        ws = {"sums": f(levels, N * C, 2), "coef": f(levels, N * C, 2) if want else None, "px": [], "py": [], "g": [],
              "mom": [f(5, N * C, H - 10, W - 10)] if want else [None]}   # filtered moments per level (forward -> backward)
        h, w = H, W
        for lv in range(1, levels):
            h, w = h // 2 + h % 2, w // 2 + w % 2   # avg_pool2d(kernel 2, padding = side % 2), MS_SSIM.py:214-216
            ws["px"].append(f(N, C, h, w))
            ws["py"].append(f(N, C, h, w))
            ws["g"].append(f(N, C, h, w) if want else None)
            ws["mom"].append(f(5, N * C, h - 10, w - 10) if want else None)
        _WORK[key] = ws
    return ws


def gan_loss(ctx: Ctx, pred: Var, target_is_real: bool, slot: int, loss_scale=1.0, grad_scale=1.0, use_lsgan=False,
             sigmoid_d=False, want_grad=True):
    """GANLoss.__call__ (networks.py:154-163).  Writes d(loss*grad_scale)/d pred into pred's gradient."""
    assert pred.C == 1, "PatchGAN logits have one channel"
    n = pred.npix
    mode = 0 if not use_lsgan else (2 if sigmoid_d else 1)
    gp = None
    if want_grad and not ctx.no_grad:
        gp, gld, gacc = pred.grad_out()
        assert gacc == 0 and gld == pred.ld and pred.fused_act is None
    ctx.L.gan_loss(pred.ptr, ctx.dt, n, pred.ld, 1.0 if target_is_real else 0.0, mode, loss_scale, slot, grad_scale,
                   gp, ctx.stream)


def l1_images(ctx: Ctx, fake: torch.Tensor, real: torch.Tensor, slot: int, grad_scale, dfake: torch.Tensor):
    """criterionL1(fake_B, real_B) (pix2pix_model.py:177); dfake += grad."""
    ctx.L.l1_loss(fake.data_ptr(), real.data_ptr(), F32, fake.numel(), slot, grad_scale,
                  dfake.data_ptr() if dfake is not None else None, 1, 0, ctx.stream)


def l1_features(ctx: Ctx, f: Var, r: Var, slot: int, grad_scale):
    """criterionL1 on one VGG tap (pix2pix_model.py:182-186); the tap is a fused-ReLU output, so its gradient is
    kept w.r.t. the pre-activation (masked by f > 0)."""
    assert f.ld == f.C and r.ld == r.C
    gp, acc, relu = None, 0, 0
    if not ctx.no_grad:
        gp, _gld, acc = f.grad_out()
        if f.fused_act is not None:
            assert f.fused_act[0] == ACT_RELU
            relu = 1
    ctx.L.l1_loss(f.ptr, r.ptr, ctx.dt, f.npix * f.C, slot, grad_scale, gp, acc, relu, ctx.stream)


def tv_loss(ctx: Ctx, fake: torch.Tensor, slot: int, grad_scale, dfake: torch.Tensor):
    """TV loss with the reference's literal 320*256 divisor and batch SUM (pix2pix_model.py:189-191, Q10)."""
    N, C, H, W = fake.shape
    ctx.L.tv_loss(fake.data_ptr(), N * C, H, W, float(320 * 256), slot, grad_scale,
                  dfake.data_ptr() if dfake is not None else None, ctx.stream)


def _affine(ctx: Ctx, img: torch.Tensor, scale, shift):
    out = torch.empty_like(img)
    N, C, H, W = img.shape
    ctx.L.nchw_to_nhwc(img.data_ptr(), out.data_ptr(), F32, N * C, 1, H, W, 1, scale, shift, ctx.stream)
    return out


def ssim_value_and_grad(ctx: Ctx, X: torch.Tensor, Y: torch.Tensor, slot: int, out_scale, dY: torch.Tensor = None,
                        grad_scale=0.0, data_range=1.0, K=(0.01, 0.03), multiscale=False, per_plane=None):
    """slot += out_scale * {ssim | ms_ssim}(X, Y)  (MS_SSIM.py:95-225, size_average=True) and, if dY is given,
    dY += grad_scale * d value / dY.  X, Y: NCHW fp32 in [0, data_range].  `per_plane` (fp32 [L,NC,2]) receives
    the raw per-plane map sums when given."""
    N, C, H, W = X.shape
    NC = N * C
    L_ = ctx.L
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    levels = 5 if multiscale else 1
    if multiscale and min(H, W) <= 160:
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    if min(H, W) < 11:
        raise ValueError("ssim needs H, W >= 11 (win_size), got %dx%d" % (H, W))
    want = dY is not None
    ws = _workspace(X, levels, want)
    sums = per_plane if per_plane is not None else ws["sums"]
    xs, ys, sizes = [X], [Y], []
    for lv in range(levels):
        h, w = xs[lv].shape[-2:]
        mom = ws["mom"][lv]
        L_.ssim_fwd(xs[lv].data_ptr(), ys[lv].data_ptr(), NC, h, w, C1, C2, sums[lv].data_ptr(),
                    mom.data_ptr() if mom is not None else None, ctx.stream)
        sizes.append((h - 10) * (w - 10))
        if lv < levels - 1:
            nx, ny = ws["px"][lv], ws["py"][lv]
            L_.avgpool2_fwd2(xs[lv].data_ptr(), nx.data_ptr(), ys[lv].data_ptr(), ny.data_ptr(), NC, h, w, ctx.stream)
            xs.append(nx)
            ys.append(ny)
    coef = ws["coef"]
    if multiscale:
        inv, wts = _consts(X.device, sizes)
        L_.msssim_combine(sums.data_ptr(), inv.data_ptr(), wts.data_ptr(), levels, NC, out_scale, slot, grad_scale,
                          coef.data_ptr() if want else None, ctx.stream)
    else:
        L_.ssim_combine(sums.data_ptr(), 1.0 / sizes[0], NC, out_scale, slot, grad_scale,
                        coef.data_ptr() if want else None, ctx.stream)
    if not want:
        return
    # backward: coarsest level first; each level adds the avg_pool adjoint of the next coarser level's dY on the fly
    g_next = None
    for lv in reversed(range(levels)):
        h, w = xs[lv].shape[-2:]
        g = dY if lv == 0 else ws["g"][lv - 1]
        L_.ssim_bwd(xs[lv].data_ptr(), ys[lv].data_ptr(), NC, h, w, C1, C2, coef[lv].data_ptr(), ws["mom"][lv].data_ptr(),
                    g.data_ptr(), 1 if lv == 0 else 0, g_next.data_ptr() if g_next is not None else None, ctx.stream)
        g_next = g


def ssim_training_loss(ctx: Ctx, real_B, fake_B, slot, w_ss, dfake):
    """loss_ssim = 1 - ssim((real_B+1)/2, (fake_B+1)/2, data_range=1) (pix2pix_model.py:193-195); the slot must
    start at 1.0.  dfake += w_ss * d loss / d fake_B."""
    X, Y = _affine(ctx, real_B, 0.5, 0.5), _affine(ctx, fake_B, 0.5, 0.5)
    dY = None
    if dfake is not None:
        dY = torch.empty_like(Y)
        ctx.zero_(dY)
    ssim_value_and_grad(ctx, X, Y, slot, -1.0, dY, grad_scale=-w_ss)
    if dfake is not None:  # dfake += 0.5 * dY
        N, C, H, W = Y.shape
        ctx.L.nhwc_to_nchw(dY.data_ptr(), F32, 1, dfake.data_ptr(), N * C, 1, H, W, 0.5, 1, ctx.stream)
