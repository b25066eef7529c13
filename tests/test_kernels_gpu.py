"""GPU parity of every kernel family against plain torch CPU fp32 (the arithmetic the reference dispatches to),
called through the C ABI.  Tolerances: fp32 validation mode 1e-4, bf16 mode 2e-2 (north_star)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import TOL, ctx_for, from_nhwc, make_params, q, rel, set_grad, to_var, var_data, var_grad  # noqa: E402
from dsgan_b200 import engine as E  # noqa: E402

PREC = ["fp32", "bf16"]


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("cfg", [
    # (N, Ci, H, W, Co, k, stride, pad, act)
    (2, 3, 16, 16, 12, 1, 1, 0, "gelu"), (2, 64, 8, 8, 256, 1, 1, 0, "gelu"), (1, 70, 9, 7, 33, 1, 1, 0, None),
    (2, 6, 32, 32, 32, 4, 2, 1, "leaky"), (2, 32, 16, 16, 64, 4, 2, 1, None), (1, 16, 9, 9, 8, 4, 1, 1, None),
    (2, 64, 8, 8, 3, 3, 1, 1, None), (1, 3, 16, 16, 64, 3, 1, 1, "relu"), (1, 256, 5, 5, 1, 4, 1, 1, None),
    # shapes eligible for the tcgen05 implicit GEMM in bf16 mode (Ci % 64 == 0, Co % 32 == 0)
    (2, 64, 16, 16, 64, 3, 1, 1, "relu"), (1, 128, 32, 32, 256, 3, 1, 1, None), (1, 64, 9, 13, 96, 3, 1, 1, "relu"),
    (2, 64, 32, 32, 128, 4, 2, 1, None), (1, 128, 32, 32, 256, 4, 1, 1, None), (1, 64, 18, 10, 32, 4, 2, 1, "leaky"),
    (1, 512, 8, 8, 512, 3, 1, 1, "relu"),
    # small-channel layers (CUDA-core sc_conv backend in bf16 mode), incl. widths that are not a multiple of 4
    (2, 12, 10, 14, 64, 1, 1, 0, None), (1, 64, 12, 10, 12, 1, 1, 0, "gelu"), (2, 12, 8, 8, 3, 1, 1, 0, None),
    (1, 3, 10, 6, 32, 1, 1, 0, None), (1, 1, 9, 9, 16, 3, 1, 1, None), (2, 6, 30, 26, 32, 4, 2, 1, "leaky"),
    (1, 64, 7, 9, 3, 3, 1, 1, None), (2, 3, 17, 19, 64, 3, 1, 1, "relu"), (1, 32, 12, 12, 6, 4, 2, 1, None),
])
def test_conv2d(prec, cfg):
    N, Ci, H, W, Co, k, s, p, act = cfg
    ctx = ctx_for(prec)
    x = q(torch.randn(N, Ci, H, W, generator=_g(1)), prec)
    w = torch.randn(Co, Ci, k, k, generator=_g(2)) * (1.0 / (Ci * k * k) ** 0.5)
    b = torch.randn(Co, generator=_g(3)) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = F.conv2d(xr, wr, br, stride=s, padding=p)
    actf = {None: lambda t: t, "gelu": F.gelu, "relu": F.relu, "leaky": lambda t: F.leaky_relu(t, 0.2)}[act]
    yr = actf(pre)
    dy = torch.randn(yr.shape, generator=_g(4))
    dyq = q(dy, prec)
    # the engine keeps the gradient of a fused-activation output w.r.t. its pre-activation
    pre.backward(dyq)
    P = make_params({"w": w if k > 1 else w.reshape(Co, Ci), "b": b})
    xv = to_var(ctx, x)
    code = {None: E.ACT_NONE, "gelu": E.ACT_GELU, "relu": E.ACT_RELU, "leaky": E.ACT_LEAKY}[act]
    yv = E.conv2d(ctx, xv, P["w"], P["b"], k, s, p, act=code)
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec]
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    assert rel(P["w"].grad.cpu().reshape(w.shape), wr.grad) < tol
    assert rel(P["b"].grad.cpu(), br.grad) < tol


@pytest.mark.parametrize("prec", PREC)
def test_conv_dgrad_fused_act_chain(prec):
    """conv(act) -> conv: the second conv's input-gradient applies act' of the first (GELU uses the pre-activation)."""
    ctx = ctx_for(prec)
    for act, code in (("gelu", E.ACT_GELU), ("leaky", E.ACT_LEAKY), ("relu", E.ACT_RELU)):
        x = q(torch.randn(2, 8, 6, 6, generator=_g(1)), prec)
        w1, b1 = torch.randn(16, 8, generator=_g(2)) * 0.3, torch.randn(16, generator=_g(3)) * 0.1
        w2 = torch.randn(5, 16, generator=_g(4)) * 0.3
        xr, w1r, w2r = (t.clone().requires_grad_(True) for t in (x, w1, w2))
        actf = {"gelu": F.gelu, "relu": F.relu, "leaky": lambda t: F.leaky_relu(t, 0.2)}[act]
        yr = F.conv2d(actf(F.conv2d(xr, w1r[:, :, None, None], b1)), w2r[:, :, None, None])
        dy = q(torch.randn(yr.shape, generator=_g(5)), prec)
        yr.backward(dy)
        P = make_params({"w1": w1, "b1": b1, "w2": w2})
        xv = to_var(ctx, x)
        h = E.conv2d(ctx, xv, P["w1"], P["b1"], 1, act=code)
        yv = E.conv2d(ctx, h, P["w2"], None, 1)
        set_grad(ctx, yv, dy)
        ctx.backward()
        tol = TOL[prec] * 1.5
        assert rel(var_data(yv), yr.detach()) < tol, act
        assert rel(var_grad(xv), xr.grad) < tol, act
        assert rel(P["w1"].grad.cpu(), w1r.grad) < tol, act
        assert rel(P["w2"].grad.cpu(), w2r.grad) < tol, act


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("cfg", [(2, 16, 4, 4, 8), (1, 128, 8, 8, 64), (2, 5, 3, 6, 7), (2, 256, 16, 16, 128),
                                 (1, 128, 5, 9, 64), (1, 1024, 4, 4, 512)])
def test_conv_transpose2d(prec, cfg):
    N, Ci, H, W, Co = cfg
    ctx = ctx_for(prec)
    x = q(torch.randn(N, Ci, H, W, generator=_g(1)), prec)
    w = torch.randn(Ci, Co, 3, 3, generator=_g(2)) * (1.0 / (Ci * 2.25) ** 0.5)
    b = torch.randn(Co, generator=_g(3)) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, br, stride=2, padding=1, output_padding=1)
    dy = q(torch.randn(yr.shape, generator=_g(4)), prec)
    yr.backward(dy)
    P = make_params({"w": w, "b": b})
    xv = to_var(ctx, x)
    yv = E.conv_transpose2d(ctx, xv, P["w"], P["b"])
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec]
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    assert rel(P["w"].grad.cpu(), wr.grad) < tol
    assert rel(P["b"].grad.cpu(), br.grad) < tol


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("k", [3, 5, 7, 9])
@pytest.mark.parametrize("C", [3, 8, 64])
def test_dwconv(prec, k, C):
    ctx = ctx_for(prec)
    x = q(torch.randn(2, C, 12, 10, generator=_g(1)), prec)
    w = torch.randn(C, 1, k, k, generator=_g(2)) / k
    b = torch.randn(C, generator=_g(3)) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=k // 2, groups=C)
    dy = q(torch.randn(yr.shape, generator=_g(4)), prec)
    yr.backward(dy)
    P = make_params({"w": w, "b": b})
    xv = to_var(ctx, x)
    yv = E.dwconv(ctx, xv, P["w"], P["b"], k)
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec]
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    assert rel(P["w"].grad.cpu(), wr.grad) < tol
    assert rel(P["b"].grad.cpu(), br.grad) < tol


def _dwconv_run(x, w, b, dy, k, twice):
    """forward + backward of the bf16 depthwise conv through the engine; twice: x feeds two convs (gradient fan-in)"""
    ctx = ctx_for("bf16")
    P = make_params({"w": w, "b": b})
    xv = to_var(ctx, x)
    yv = E.dwconv(ctx, xv, P["w"], P["b"], k)
    if twice:
        y2 = E.dwconv(ctx, xv, P["w"], P["b"], k)
        set_grad(ctx, y2, dy)
    set_grad(ctx, yv, dy)
    ctx.backward()
    return var_data(yv), var_grad(xv), P["w"].grad.cpu().clone(), P["b"].grad.cpu().clone()


@pytest.mark.parametrize("cfg", [
    # (N, C, H, W, k): shapes taken by the tensor-core path (C >= 16, C % 8 == 0, H >= 16, W >= 32), ragged tiles included
    (2, 16, 32, 64, 7), (1, 24, 40, 72, 7), (2, 64, 64, 32, 3), (1, 32, 16, 32, 9), (1, 128, 48, 96, 5), (2, 32, 33, 47, 7),
    (1, 40, 70, 130, 9),
])
@pytest.mark.parametrize("twice", [False, True])
def test_dwconv_mma(cfg, twice, monkeypatch):
    """mma.sync depthwise path (dwconv_mma.cu) against torch fp32 AND against the CUDA-core kernels it replaces"""
    N, C, H, W, k = cfg
    x = q(torch.randn(N, C, H, W, generator=_g(1)), "bf16")
    w = torch.randn(C, 1, k, k, generator=_g(2)) / k
    b = torch.randn(C, generator=_g(3)) * 0.1
    dy = q(torch.randn(N, C, H, W, generator=_g(4)), "bf16")
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=k // 2, groups=C)
    yr.backward(dy * (2 if twice else 1))
    monkeypatch.setenv("DSGAN_DW_MMA", "1")
    y1, dx1, dw1, db1 = _dwconv_run(x, w, b, dy, k, twice)
    monkeypatch.setenv("DSGAN_DW_MMA", "0")
    y0, dx0, dw0, db0 = _dwconv_run(x, w, b, dy, k, twice)
    tol = TOL["bf16"]
    assert rel(y1, yr.detach()) < tol and rel(dx1, xr.grad) < tol
    assert rel(dw1, wr.grad) < 1e-3 and rel(db1, br.grad) < 1e-3     # exact bf16 products, fp32 sums
    # against the fp32-weight CUDA-core kernels: the tensor-core path rounds the taps to bf16 (2^-9 relative per tap)
    assert rel(y1, y0) < 5e-3 and rel(dx1, dx0) < 5e-3
    assert rel(dw1, dw0) < 1e-4 and rel(db1, db0) < 1e-4


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("shape", [(2, 128, 32, 64), (1, 64, 48, 40), (2, 32, 12, 10), (1, 256, 16, 16)])
def test_dwconv_multi(prec, shape):
    """MidMLKA's four depthwise branches (k = 3, 5, 7, 9 on the channel quarters) in one launch per pass"""
    N, C, H, W = shape
    q4 = C // 4
    ctx = ctx_for(prec)
    x = q(torch.randn(N, C, H, W, generator=_g(1)), prec)
    dy = q(torch.randn(N, C, H, W, generator=_g(4)), prec)
    ws = [torch.randn(q4, 1, k, k, generator=_g(10 + k)) / k for k in (3, 5, 7, 9)]
    bs = [torch.randn(q4, generator=_g(20 + k)) * 0.1 for k in (3, 5, 7, 9)]
    xr = x.clone().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    yr = torch.cat([F.conv2d(xr[:, i * q4:(i + 1) * q4], wr[i], br[i], padding=k // 2, groups=q4)
                    for i, k in enumerate((3, 5, 7, 9))], 1)
    yr.backward(dy)
    P = make_params(dict([("w%d" % i, w) for i, w in enumerate(ws)] + [("b%d" % i, b) for i, b in enumerate(bs)]))
    xv = to_var(ctx, x)
    yv = E.dwconv_multi(ctx, xv, [(P["w%d" % i], P["b%d" % i], k) for i, k in enumerate((3, 5, 7, 9))])
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec]
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    for i in range(4):
        assert rel(P["w%d" % i].grad.cpu(), wr[i].grad) < tol
        assert rel(P["b%d" % i].grad.cpu(), br[i].grad) < tol


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("act", [None, "gelu", "leaky"])
@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (2, 64, 31, 31), (1, 130, 4, 4), (2, 128, 32, 32), (1, 16, 64, 64),
                                   (3, 256, 16, 16), (1, 64, 72, 72)])
def test_instance_norm(prec, act, with_res, shape):
    ctx = ctx_for(prec)
    x = q(torch.randn(shape, generator=_g(1)) * 2 + 3, prec)   # non-zero mean: exercises the shifted sums
    r = q(torch.randn(shape, generator=_g(2)), prec)
    xr, rr = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    u = F.instance_norm(xr, eps=1e-5) + (rr if with_res else 0)
    actf = {None: lambda t: t, "gelu": F.gelu, "leaky": lambda t: F.leaky_relu(t, 0.2)}[act]
    yr = actf(u)
    dy = q(torch.randn(shape, generator=_g(3)), prec)
    yr.backward(dy)
    xv, rv = to_var(ctx, x), to_var(ctx, r)
    code = {None: E.ACT_NONE, "gelu": E.ACT_GELU, "leaky": E.ACT_LEAKY}[act]
    yv = E.inorm(ctx, xv, act=code, res=rv if with_res else None)
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec] * (3 if prec == "fp32" else 1.5)
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol * 2
    if with_res:
        assert rel(var_grad(rv), rr.grad) < tol


@pytest.mark.parametrize("prec", PREC)
def test_inorm_into_concat_slice_and_fanout(prec):
    """IN+GELU written into channels [0:C) of a concat buffer, skip copied into [C:2C); the skip tensor is also
    consumed elsewhere (gradient fan-out accumulates)."""
    ctx = ctx_for(prec)
    x = q(torch.randn(2, 8, 6, 6, generator=_g(1)), prec)
    s = q(torch.randn(2, 8, 6, 6, generator=_g(2)), prec)
    xr, sr = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
    cat = torch.cat((F.gelu(F.instance_norm(xr)), sr), 1)
    out = cat * 1.0
    extra = F.max_pool2d(sr, 2)
    dy, de = q(torch.randn(out.shape, generator=_g(3)), prec), q(torch.randn(extra.shape, generator=_g(4)), prec)
    torch.autograd.backward([out, extra], [dy, de])
    xv, sv = to_var(ctx, x), to_var(ctx, s)
    cv = ctx.new(2, 6, 6, 16)
    E.inorm(ctx, xv, act=E.ACT_GELU, out=cv.slice(0, 8))
    E.concat_into(ctx, cv, 8, sv)
    ev = E.maxpool(ctx, sv, 2)
    set_grad(ctx, cv, dy)
    set_grad(ctx, ev, de)
    ctx.backward()
    tol = TOL[prec] * 2
    assert rel(var_data(cv), cat.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    assert rel(var_grad(sv), sr.grad) < tol


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("k", [2, 4, 16])
def test_maxpool(prec, k):
    ctx = ctx_for(prec)
    x = q(torch.randn(2, 20, 32, 16, generator=_g(1)), prec)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, k)
    dy = q(torch.randn(yr.shape, generator=_g(2)), prec)
    yr.backward(dy)
    xv = to_var(ctx, x)
    yv = E.maxpool(ctx, xv, k)
    set_grad(ctx, yv, dy)
    ctx.backward()
    assert rel(var_data(yv), yr.detach()) < 1e-6
    assert rel(var_grad(xv), xr.grad) < 1e-6


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("C,H", [(32, 16), (128, 8), (256, 3)])
def test_ca_scale(prec, C, H):
    ctx = ctx_for(prec)
    x = q(torch.randn(2, C, H, H, generator=_g(1)), prec)
    fc1 = torch.randn(C // 8, C, 1, 1, generator=_g(2)) * 0.2
    fc2 = torch.randn(C, C // 8, 1, 1, generator=_g(3)) * 0.2
    sl = torch.tensor([0.25])
    xr, f1, f2, slr = (t.clone().requires_grad_(True) for t in (x, fc1, fc2, sl))

    def mlp(v):
        return F.conv2d(F.prelu(F.conv2d(v, f1), slr), f2)
    yr = xr * torch.sigmoid(mlp(F.adaptive_avg_pool2d(xr, 1)) + mlp(F.adaptive_max_pool2d(xr, 1)))
    dy = q(torch.randn(yr.shape, generator=_g(4)), prec)
    yr.backward(dy)
    P = make_params({"fc1": fc1, "fc2": fc2, "sl": sl})
    xv = to_var(ctx, x)
    yv = E.ca_scale(ctx, xv, P["fc1"], P["sl"], P["fc2"])
    set_grad(ctx, yv, dy)
    ctx.backward()
    tol = TOL[prec] * 2
    assert rel(var_data(yv), yr.detach()) < tol
    assert rel(var_grad(xv), xr.grad) < tol
    assert rel(P["fc1"].grad.cpu(), f1.grad) < tol * 2
    assert rel(P["fc2"].grad.cpu(), f2.grad) < tol * 2
    assert rel(P["sl"].grad.cpu(), slr.grad) < tol * 2


@pytest.mark.parametrize("prec", PREC)
def test_add_n(prec):
    ctx = ctx_for(prec)
    xs = [q(torch.randn(2, 8, 4, 4, generator=_g(i)), prec) for i in range(5)]
    vs = [to_var(ctx, x) for x in xs]
    yv = E.add_n(ctx, vs)
    dy = q(torch.randn(2, 8, 4, 4, generator=_g(9)), prec)
    set_grad(ctx, yv, dy)
    ctx.backward()
    assert rel(var_data(yv), sum(xs)) < TOL[prec]
    for v in vs:
        assert rel(var_grad(v), dy) < 1e-6


def test_image_layout_roundtrip():
    ctx = ctx_for("fp32")
    img = torch.randn(2, 3, 16, 24, generator=_g(1)).cuda()
    v = E.image_to_nhwc(ctx, img)
    assert rel(var_data(v), img.cpu()) == 0
    v.g = v.t.clone()
    out = torch.ones_like(img)
    E.nhwc_grad_to_image(ctx, v, out, alpha=2.0, acc=1)
    assert rel(out.cpu(), 1 + 2 * img.cpu()) < 1e-7


def test_adam_matches_torch():
    ctx = ctx_for("fp32")
    p0 = torch.randn(1000, generator=_g(1))
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=2e-4, betas=(0.5, 0.999))
    pc, m, v = p0.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for t in range(1, 4):
        g = torch.randn(1000, generator=_g(10 + t))
        p.grad = g.clone()
        opt.step()
        gc = (g * 4).cuda()  # grad_scale 0.25 emulates the 1/world averaging
        ctx.L.adam_step(pc.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), 1000, 2e-4, 0.5, 0.999, 1e-8, t, 0.25,
                        None, ctx.stream)
    assert rel(pc.cpu(), p.detach()) < 1e-6


@pytest.mark.parametrize("nlev,C,H,W", [(4, 64, 32, 48), (3, 128, 16, 24), (2, 256, 8, 12), (4, 136, 16, 16)])
@pytest.mark.parametrize("ties", [False, True])
def test_multi_maxpool_matches_separate_pools(nlev, C, H, W, ties):
    """MaxPool2d(2..2^nlev) of one tensor in a single pass + the combined backward, against torch (first-maximum routing,
    SURVEY Q5) -- with heavily tied inputs too (small integers in bf16), where the scan-order rule decides the routing.  The
    input is a channel slice of a wider buffer and its gradient accumulates onto an existing one, like R1..R4 in the generator."""
    ctx = ctx_for("bf16")
    g = _g(7 + nlev + C)
    x = torch.randint(-3, 4, (2, C, H, W), generator=g).float() if ties else q(torch.randn(2, C, H, W, generator=g), "bf16")
    xr = x.clone().requires_grad_(True)
    outs = [F.max_pool2d(xr, 2 << l) for l in range(nlev)]
    dys = [q(torch.randn(o.shape, generator=g), "bf16") for o in outs]
    if nlev >= 3:
        dys[1] = None                      # one scale without any gradient
    torch.autograd.backward([o for o, d in zip(outs, dys) if d is not None], [d for d in dys if d is not None])
    base = q(torch.randn(2, C, H, W, generator=g), "bf16")
    wide = ctx.new(2, H, W, 2 * C)
    wide.t.zero_()
    xv = wide.slice(C, C)
    wide.t[..., C:] = x.permute(0, 2, 3, 1).to("cuda", ctx.tdtype)
    ys = E.multi_maxpool(ctx, xv, nlev)
    assert len(ys) == nlev
    for y, o in zip(ys, outs):
        assert torch.equal(var_data(y), o.detach())
    for y, d in zip(ys, dys):
        if d is not None:
            set_grad(ctx, y, d)
    # pre-existing gradient on the slice: the pooling backward must accumulate
    wide.g = torch.zeros((2, H, W, 2 * C), dtype=ctx.tdtype, device="cuda")
    wide.g[..., C:] = base.permute(0, 2, 3, 1).to("cuda", ctx.tdtype)
    ctx.backward()
    got = wide.g[..., C:].float().cpu().permute(0, 3, 1, 2)
    want = q(base + xr.grad, "bf16")
    assert rel(got, want) < 4e-3
    assert float((wide.g[..., :C].float().abs()).max()) == 0.0
    if ties:   # routing must agree element for element (sums of bf16 values: compare where a gradient landed)
        assert torch.equal((got - base) != 0, xr.grad != 0) or rel(got, want) < 1e-3
