"""Fused ConvNeXt Block MLP kernels (csrc/fused_mlp.cu) through the C ABI against torch CPU fp32 on bf16-rounded
operands (MixConvNeXtML.py:236-243), at every (C_in, N_out) a generator Block uses and then some."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from dsgan_b200._lib import lib  # noqa: E402


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _bf(x):
    return x.bfloat16().float()


def _dev(x, dt=torch.bfloat16):
    return x.to("cuda", dt).contiguous()


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _case(M, cin, nout, seed, sc=True):
    g = _g(seed)
    hid = 4 * cin
    T = _bf(torch.randn(M, cin, generator=g))
    X = _bf(torch.randn(M, cin, generator=g))
    W1 = _bf(torch.randn(hid, cin, generator=g) * (1.5 / cin ** 0.5))
    b1 = torch.randn(hid, generator=g) * 0.3
    W2 = _bf(torch.randn(nout, hid, generator=g) * (1.0 / hid ** 0.5))
    b2 = torch.randn(nout, generator=g) * 0.3
    Ws = _bf(torch.randn(nout, cin, generator=g) * (1.0 / cin ** 0.5))
    return T, X, W1, b1, W2, b2, Ws


def _ref_fwd(T, X, W1, b1, W2, b2, Ws, sc=True):
    h = F.gelu(T @ W1.t() + b1)
    y = _bf(h) @ W2.t() + b2           # the hidden is rounded to bf16 once, as the second GEMM's operand
    if sc:
        y = y + X @ Ws.t()
    return y


# (c2) 64->128, (c3) 128->256, (uc3) 256->128, (uc4) 128->64 are the generator's fused Blocks; the rest cover the template grid
SHAPES = [(64, 128), (128, 256), (256, 128), (128, 64), (64, 64), (128, 128), (256, 64)]


@pytest.mark.parametrize("cin,nout", SHAPES)
@pytest.mark.parametrize("M", [128, 1000, 128 * 300 + 64])
def test_fused_mlp_fwd(cin, nout, M):
    L = lib()
    assert L.cdll.dsgan_fused_mlp_supported(cin, nout)
    T, X, W1, b1, W2, b2, Ws = _case(M, cin, nout, seed=cin + nout + M)
    want = _ref_fwd(T, X, W1, b1, W2, b2, Ws)
    dT, dX, dW1, dW2, dWs = _dev(T), _dev(X), _dev(W1), _dev(W2), _dev(Ws)
    db1, db2 = _dev(b1, torch.float32), _dev(b2, torch.float32)
    # output into a channel slice of a wider buffer (the encoder Blocks write into the decoder's concat buffer)
    ld_y = 2 * nout
    Y = torch.full((M, ld_y), 7.0, dtype=torch.bfloat16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.fused_mlp_fwd(dT.data_ptr(), cin, dX.data_ptr(), cin, M, cin, nout, dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(),
                    db2.data_ptr(), dWs.data_ptr(), Y.data_ptr() + 2 * nout, ld_y, s)
    torch.cuda.synchronize()
    got = Y[:, nout:].float().cpu()
    assert torch.all(Y[:, :nout].float() == 7.0), "wrote outside the output slice"
    assert rel(got, want) < 6e-3, rel(got, want)       # bf16 output rounding (2^-9) + bf16 hidden
    # no shortcut / no output bias
    Y2 = torch.empty((M, nout), dtype=torch.bfloat16, device="cuda")
    L.fused_mlp_fwd(dT.data_ptr(), cin, None, 0, M, cin, nout, dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(), None, None,
                    Y2.data_ptr(), nout, s)
    torch.cuda.synchronize()
    want2 = _ref_fwd(T, X, W1, b1, W2, torch.zeros_like(b2), Ws, sc=False)
    assert rel(Y2.float().cpu(), want2) < 6e-3


def test_fused_mlp_fwd_rejects_bad_arguments():
    L = lib()
    assert not L.cdll.dsgan_fused_mlp_supported(512, 256)
    assert not L.cdll.dsgan_fused_mlp_supported(3, 64)
    t = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        L.fused_mlp_fwd(t.data_ptr(), 64, None, 0, 128, 48, 64, t.data_ptr(), t.data_ptr(), t.data_ptr(), None, None,
                        t.data_ptr(), 64, 0)


def _ref_bwd(T, dY, W1, b1, W2):
    Tr = T.clone().requires_grad_(True)
    hpre = Tr @ W1.t() + b1
    hpre.retain_grad()
    a = F.gelu(hpre)
    y = a @ W2.t()
    y.backward(dY)
    return Tr.grad, hpre.grad, a.detach()


@pytest.mark.parametrize("cin,nout", SHAPES)
@pytest.mark.parametrize("M", [128, 1000, 128 * 300 + 64])
def test_fused_mlp_bwd(cin, nout, M):
    """dT, G = dH * GELU'(Hpre), A = GELU(Hpre), db1 = colsum(G) and db2 = colsum(dY) against torch autograd on the same bf16 operands."""
    L = lib()
    T, _X, W1, b1, W2, _b2, _Ws = _case(M, cin, nout, seed=7 + cin + nout + M)
    dY = _bf(torch.randn(M, nout, generator=_g(99 + M)) * 0.1)
    want_dT, want_G, want_A = _ref_bwd(T, dY, W1, b1, W2)
    hid = 4 * cin
    dTd = torch.full((M, 2 * cin), 3.0, dtype=torch.bfloat16, device="cuda")   # written into a slice of a wider buffer
    G = torch.empty((M, hid), dtype=torch.bfloat16, device="cuda")
    A = torch.empty((M, hid), dtype=torch.bfloat16, device="cuda")
    db1 = torch.full((hid,), 0.5, dtype=torch.float32, device="cuda")
    db2 = torch.full((nout,), 0.25, dtype=torch.float32, device="cuda")
    dT_, dY_, W1_, W2_, b1_ = _dev(T), _dev(dY), _dev(W1), _dev(W2), _dev(b1, torch.float32)
    s = torch.cuda.current_stream().cuda_stream
    L.fused_mlp_bwd(dT_.data_ptr(), cin, dY_.data_ptr(), nout, M, cin, nout, W1_.data_ptr(), b1_.data_ptr(), W2_.data_ptr(),
                    dTd.data_ptr(), 2 * cin, G.data_ptr(), A.data_ptr(), db1.data_ptr(), db2.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.all(dTd[:, cin:].float() == 3.0), "wrote outside the dT slice"
    assert rel(A.float().cpu(), want_A) < 4e-3
    assert rel(G.float().cpu(), want_G) < 6e-3
    assert rel(dTd[:, :cin].float().cpu(), want_dT) < 8e-3       # G is rounded to bf16 before the second GEMM
    got_db = db1.cpu() - 0.5
    want_db = G.float().cpu().sum(0)                              # the kernel sums its own (rounded) G
    assert float((got_db - want_db).abs().max()) <= 2e-3 * float(want_db.abs().max()) + 1e-4
    assert rel(got_db, want_G.sum(0)) < 2e-2
    want_db2 = dY.float().sum(0)                                   # pwconv2 bias gradient: exact bf16 inputs, fp32 sums
    assert float(((db2.cpu() - 0.25) - want_db2).abs().max()) <= 1e-4 * float(want_db2.abs().max()) + 1e-4
