"""Helpers shared by the -m gpu tests: move NCHW CPU tensors into engine Vars and back, relative-error metric."""
import torch

from dsgan_b200.engine import Ctx, Param, Var

_CTX = {}


def ctx_for(precision):
    if precision not in _CTX:
        _CTX[precision] = Ctx("cuda:0", precision)
    return _CTX[precision]


def to_var(ctx, x_nchw):
    N, C, H, W = x_nchw.shape
    v = ctx.new(N, H, W, C)  # padded pixel pitch for odd channel counts in bf16 mode
    if v.ld > C:
        v.t[..., C:] = float("nan")  # poison the pad lanes: kernels may carry them along but must never mix them in
    v.t[..., :C] = x_nchw.permute(0, 2, 3, 1).to("cuda:0", ctx.tdtype)
    return v


def from_nhwc(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def var_data(v):
    assert v.parent is None
    return from_nhwc(v.t[..., :v.C])


def var_grad(v):
    return from_nhwc(v.g[..., :v.C])


def set_grad(ctx, v, dy_nchw):
    gp, ld, acc = v.grad_out()
    assert acc == 0 and ld == v.ld
    if v.ld > v.C:
        v.g[..., v.C:] = float("nan")
    v.g[..., :v.C] = dy_nchw.permute(0, 2, 3, 1).to("cuda:0", ctx.tdtype)


def make_params(tensors):
    out = {}
    for k, t in tensors.items():
        d = t.detach().clone().float().cuda()
        sh = d.bfloat16()
        out[k] = Param(k, d, torch.zeros_like(d), sh.data_ptr())
        out[k].cache["bf16"] = sh  # keep the shadow alive
    return out


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def q(x, precision):
    """Round a reference input the way the engine stores it (bf16 mode keeps activations in bf16)."""
    return x.bfloat16().float() if precision == "bf16" else x


TOL = {"fp32": 1e-4, "bf16": 2e-2}
