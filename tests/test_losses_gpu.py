"""GPU parity of the loss kernels (value and gradient) against the oracle's restatement of the reference losses."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import dsgan_oracle as O  # noqa: E402
from gpu_util import ctx_for, rel, to_var, var_grad  # noqa: E402
from dsgan_b200 import MS_SSIM as K  # noqa: E402
from dsgan_b200 import losses  # noqa: E402


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _slot(ctx, init=0.0):
    t = torch.full((1,), init, dtype=torch.float32, device="cuda:0")
    return t, t.data_ptr()


@pytest.mark.parametrize("mode", ["bce", "mse"])
@pytest.mark.parametrize("real", [True, False])
def test_gan_loss(mode, real):
    ctx = ctx_for("fp32")
    x = torch.randn(3, 1, 30, 30, generator=_g(1)) * 3
    xr = x.clone().requires_grad_(True)
    lr = O.gan_loss(xr, real, use_lsgan=(mode == "mse"))
    (0.5 * lr).backward()
    v = to_var(ctx, x)
    t, p = _slot(ctx)
    losses.gan_loss(ctx, v, real, p, 1.0, 0.5, use_lsgan=(mode == "mse"))
    assert abs(float(t) - float(lr)) < 1e-5 * max(1, abs(float(lr)))
    assert rel(var_grad(v), xr.grad) < 1e-5


def test_l1_and_tv():
    ctx = ctx_for("fp32")
    a, b = torch.randn(2, 3, 32, 48, generator=_g(1)), torch.randn(2, 3, 32, 48, generator=_g(2))
    ar = a.clone().requires_grad_(True)
    l1, tv = F.l1_loss(ar, b), O.tv_loss(ar)
    (l1 + 2.0 * tv).backward()
    ac, bc = a.cuda(), b.cuda()
    d = torch.zeros_like(ac)
    t1, p1 = _slot(ctx)
    t2, p2 = _slot(ctx)
    losses.l1_images(ctx, ac, bc, p1, 1.0, d)
    losses.tv_loss(ctx, ac, p2, 2.0, d)
    assert abs(float(t1) - float(l1)) < 1e-6
    assert abs(float(t2) - float(tv)) < 1e-5 * float(tv)
    assert rel(d.cpu(), ar.grad) < 1e-5


@pytest.mark.parametrize("shape,noise", [((2, 3, 256, 256), 0.1), ((2, 3, 256, 256), None), ((3, 3, 48, 40), 0.2),
                                         ((1, 2, 11, 11), 0.1), ((1, 1, 37, 75), 0.3)])
def test_ssim_value_and_grad(shape, noise):
    X = torch.rand(shape, generator=_g(3))
    Y = torch.rand(shape, generator=_g(4)) if noise is None else \
        (X + noise * torch.randn(shape, generator=_g(4))).clamp(0, 1)
    Yr = Y.clone().requires_grad_(True)
    want = O.ssim(X, Yr, 1.0)
    want.backward()
    Yc = Y.cuda().requires_grad_(True)
    got = K.ssim(X.cuda(), Yc, data_range=1, size_average=True)
    got.backward()
    assert abs(float(got) - float(want)) < 2e-5
    assert rel(Yc.grad.cpu(), Yr.grad) < 2e-4
    per = K.ssim(X.cuda(), Y.cuda(), data_range=1, size_average=False).cpu()
    assert torch.allclose(per, O.ssim(X, Y, 1.0, size_average=False), atol=2e-5)


@pytest.mark.parametrize("shape,noise", [((2, 3, 256, 256), 0.1), ((2, 3, 256, 256), None), ((1, 3, 176, 192), 0.05),
                                         ((1, 3, 200, 168), 0.05), ((2, 2, 161, 250), 0.1)])   # odd pyramid levels
def test_ms_ssim_value_and_grad(shape, noise):
    X = torch.rand(shape, generator=_g(3))
    Y = torch.rand(shape, generator=_g(4)) if noise is None else \
        (X + noise * torch.randn(shape, generator=_g(4))).clamp(0, 1)
    Yr = Y.clone().requires_grad_(True)
    want = O.ms_ssim(X, Yr, 1.0)
    want.backward()
    Yc = Y.cuda().requires_grad_(True)
    got = K.ms_ssim(X.cuda(), Yc, data_range=1, size_average=True)
    got.backward()
    assert abs(float(got) - float(want)) < 5e-5
    assert rel(Yc.grad.cpu(), Yr.grad) < 5e-4
    per = K.ms_ssim(X.cuda(), Y.cuda(), data_range=1, size_average=False).cpu()
    assert torch.allclose(per, O.ms_ssim(X, Y, 1.0, size_average=False), atol=5e-5)


def test_ms_ssim_against_reference_golden(golden_dir):
    """Fixtures written by the reference's own MS_SSIM.py (oracle/make_golden.py)."""
    recs = json.load(open(os.path.join(golden_dir, "ms_ssim.json")))
    for rec in recs[:2] + recs[4:5]:       # [4]: 200x200, odd pyramid levels
        g = _g(rec["seed"])
        n, hw = rec["n"], rec["hw"]
        X = torch.rand(n, 3, hw, hw, generator=g)
        Y = torch.rand(n, 3, hw, hw, generator=g) if rec["noise"] is None else \
            (X + rec["noise"] * torch.randn(n, 3, hw, hw, generator=g)).clamp(0, 1)
        for fn in ("ssim", "ms_ssim"):
            Yc = Y.cuda().requires_grad_(True)
            v = getattr(K, fn)(X.cuda(), Yc, data_range=1)
            v.backward()
            assert abs(float(v) - rec[fn]) < 5e-5, (fn, float(v), rec[fn])
            fp = O.fingerprint(Yc.grad.cpu())
            assert abs(fp[0] - rec[fn + "_grad"][0]) < 1e-3 * rec[fn + "_grad"][0]


def test_ssim_identity_and_errors():
    X = torch.rand(1, 3, 176, 176, generator=_g(1)).cuda()
    assert abs(float(K.ssim(X, X, data_range=1)) - 1.0) < 1e-5
    assert abs(float(K.ms_ssim(X, X, data_range=1)) - 1.0) < 1e-4
    with pytest.raises(ValueError):
        K.ssim(X, X[:, :, :100], data_range=1)
    with pytest.raises(AssertionError):
        K.ms_ssim(X[:, :, :160, :160].contiguous(), X[:, :, :160, :160].contiguous(), data_range=1)
    with pytest.raises(ValueError):
        K.ssim(X, X, data_range=1, win_size=10)


def test_ms_ssim_full_size_properties():
    """BASELINE configs[3] size (64 x 3 x 256 x 256): size-independent properties, plus the CPU oracle on a 2-image slice."""
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.rand(64, 3, 256, 256, device="cuda", generator=g)
    Y = (X + 0.1 * torch.randn(X.shape, device="cuda", generator=g)).clamp(0, 1)
    for fn, tol in ((K.ms_ssim, 1e-4), (K.ssim, 1e-5)):
        per = fn(X, Y, data_range=1, size_average=False)
        assert per.shape == (64,)
        sub = fn(X[8:12].contiguous(), Y[8:12].contiguous(), data_range=1, size_average=False)
        assert torch.allclose(per[8:12], sub, atol=1e-6)                       # images are independent
        assert abs(float(fn(X, Y, data_range=1)) - float(per.mean())) < 1e-5   # the mean is the mean of the per-image values
        assert float((fn(X, X, data_range=1, size_average=False) - 1).abs().max()) < tol   # identity
        Yc = Y.clone().requires_grad_(True)
        fn(X, Yc, data_range=1).backward()
        Ys = Y[8:12].clone().requires_grad_(True)
        fn(X[8:12].contiguous(), Ys, data_range=1).backward()
        assert rel(Yc.grad[8:12].cpu() * 16, Ys.grad.cpu()) < 1e-4            # d mean / dY of a slice, rescaled 64/4
        ref = getattr(O, fn.__name__)(X[:2].cpu(), Y[:2].cpu(), 1.0, size_average=False)
        assert torch.allclose(per[:2].cpu(), ref, atol=5e-5)
