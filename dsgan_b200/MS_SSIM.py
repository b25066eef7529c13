"""Drop-in for the reference's DSGAN/MS_SSIM.py public functions (ssim :95-150, ms_ssim :153-225, the SSIM /
MS_SSIM modules :228-311) on CUDA tensors, backed by the sm_100a SSIM kernels.  Same argument names, same
ValueError / AssertionError behaviour; gradients flow to Y (and X by symmetry is not provided: the reference's
training call differentiates w.r.t. the second argument only, pix2pix_model.py:193-195)."""
import torch

from .engine import Ctx
from .losses import MS_WEIGHTS, ssim_value_and_grad

_ctx_cache = {}


def _ctx(device):
    key = str(device)
    if key not in _ctx_cache:
        _ctx_cache[key] = Ctx(device, "fp32")
    return _ctx_cache[key]


def _check(X, Y, win_size, win_sigma, win, weights=None):
    if not X.shape == Y.shape:
        raise ValueError(f"Input images should have the same dimensions, but got {X.shape} and {Y.shape}.")
    for d in range(len(X.shape) - 1, 1, -1):
        X, Y = X.squeeze(dim=d), Y.squeeze(dim=d)
    if len(X.shape) != 4:
        raise ValueError(f"Input images should be 4-d tensors, but got {X.shape}")
    if not X.type() == Y.type():
        raise ValueError(f"Input images should have the same dtype, but got {X.type()} and {Y.type()}.")
    if win is not None:
        win_size = win.shape[-1]
    if not (win_size % 2 == 1):
        raise ValueError("Window size should be odd.")
    if win is not None or win_size != 11 or win_sigma != 1.5:
        raise NotImplementedError("the sm_100a SSIM kernels implement the reference's default 11-tap, sigma=1.5 window")
    if weights is not None and tuple(weights) != MS_WEIGHTS:
        raise NotImplementedError("custom ms_ssim weights are not supported")
    if not X.is_cuda:
        raise RuntimeError("dsgan_b200.MS_SSIM runs on CUDA tensors only (no CPU fallback)")
    return X.contiguous().float(), Y.contiguous().float()


class _SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, X, Y, data_range, K, multiscale, size_average):
        ctx = _ctx(X.device)
        N, C = X.shape[:2]
        levels = 5 if multiscale else 1
        val = torch.zeros(1, dtype=torch.float32, device=X.device)
        sums = torch.empty((levels, N * C, 2), dtype=torch.float32, device=X.device)
        need = Y.requires_grad
        if X.requires_grad:
            raise NotImplementedError("gradient w.r.t. X is not provided (the training call differentiates Y only, "
                                      "pix2pix_model.py:193-195); pass X.detach()")
        if need and not size_average:
            raise NotImplementedError("per-image (size_average=False) gradients are not provided")
        dY = torch.zeros_like(Y) if need else None
        ssim_value_and_grad(ctx, X, Y, val.data_ptr(), 1.0, dY, 1.0, float(data_range), K, multiscale, per_plane=sums)
        fctx.dY, fctx.size_average = dY, size_average
        if size_average:
            return val[0]
        # per-image means (size_average=False): tiny host-side epilogue over the [L, N*C] plane statistics
        H, W = X.shape[-2:]
        if not multiscale:
            per = sums[0, :, 0] / float((H - 10) * (W - 10))
        else:
            per = torch.ones(N * C, device=X.device)
            h, w = H, W
            for lv in range(5):
                m = sums[lv, :, 0 if lv == 4 else 1] / float((h - 10) * (w - 10))
                per = per * torch.relu(m) ** MS_WEIGHTS[lv]
                h, w = h // 2 + h % 2, w // 2 + w % 2
        fctx.dY = None  # per-image gradients are not provided
        return per.view(N, C).mean(1)

    @staticmethod
    def backward(fctx, g):
        if fctx.dY is None:
            return None, None, None, None, None, None
        return None, fctx.dY * g, None, None, None, None


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03),
         nonnegative_ssim=False):
    X, Y = _check(X, Y, win_size, win_sigma, win)
    if nonnegative_ssim:
        raise NotImplementedError("nonnegative_ssim=True is off the reference's training path")
    return _SSIMFn.apply(X, Y, data_range, tuple(K), False, size_average)


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, weights=None,
            K=(0.01, 0.03)):
    X, Y = _check(X, Y, win_size, win_sigma, win, weights)
    smaller_side = min(X.shape[-2:])
    assert smaller_side > (win_size - 1) * (2 ** 4), \
        "Image size should be larger than %d due to the 4 downsamplings in ms-ssim" % ((win_size - 1) * (2 ** 4))
    return _SSIMFn.apply(X, Y, data_range, tuple(K), True, size_average)


class SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 K=(0.01, 0.03), nonnegative_ssim=False):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma, K=K,
                       nonnegative_ssim=nonnegative_ssim)

    def forward(self, X, Y):
        return ssim(X, Y, **self.kw)


class MS_SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 weights=None, K=(0.01, 0.03)):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma,
                       weights=weights, K=K)

    def forward(self, X, Y):
        return ms_ssim(X, Y, **self.kw)
