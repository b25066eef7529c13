"""GPU parity of the tcgen05/TMEM/TMA GEMM family (forward, input-gradient, weight-gradient) against torch CPU fp32
matmul on the same bf16-rounded operands, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from dsgan_b200._lib import lib  # noqa: E402


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _stream():
    return torch.cuda.current_stream().cuda_stream


SHAPES = [(128, 64, 64), (256, 128, 128), (1000, 256, 64), (4096, 512, 128), (300, 2048, 512), (16384, 64, 512),
          (777, 1024, 4096), (128, 96, 64), (5000, 128, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_forward(M, N, K):
    L = lib()
    A = (torch.randn(M, K, generator=_g(1))).bfloat16()
    W = (torch.randn(N, K, generator=_g(2)) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=_g(3))
    ref = A.float() @ W.float().t() + bias
    Ad, Wd, bd = A.cuda(), W.cuda(), bias.cuda()
    C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    assert L.cdll.dsgan_tc_gemm_supported(0, M, N, K, K, K, N) == 1
    L.tc_gemm(0, Ad.data_ptr(), K, Wd.data_ptr(), K, M, N, K, C.data_ptr(), N, bd.data_ptr(), None, 0, None, 0, 0, 0, 0,
              _stream())
    torch.cuda.synchronize()
    assert _rel(C.float().cpu(), ref) < 6e-3


def test_tc_gemm_forward_gelu_pre_and_accumulate_and_pitch():
    L = lib()
    M, N, K = 1500, 256, 128
    big = torch.randn(M, 2 * K, generator=_g(1)).bfloat16()       # A is a channel slice of a wider buffer
    A = big[:, K:]
    W = (torch.randn(N, K, generator=_g(2)) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=_g(3))
    pre_ref = A.float() @ W.float().t() + bias
    bigd, Wd, bd = big.cuda(), W.cuda(), bias.cuda()
    C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    L.tc_gemm(0, bigd.data_ptr() + 2 * K, 2 * K, Wd.data_ptr(), K, M, N, K, C.data_ptr(), N, bd.data_ptr(),
              pre.data_ptr(), N, None, 0, 3, 0, 0, _stream())
    torch.cuda.synchronize()
    assert _rel(pre.float().cpu(), pre_ref) < 6e-3
    assert _rel(C.float().cpu(), F.gelu(pre_ref)) < 8e-3
    # accumulate into an existing output (pwconv2 + shortcut, MixConvNeXtML.py:242)
    C0 = torch.randn(M, N, generator=_g(4)).bfloat16()
    C = C0.cuda()
    L.tc_gemm(0, bigd.data_ptr() + 2 * K, 2 * K, Wd.data_ptr(), K, M, N, K, C.data_ptr(), N, None, None, 0, None, 0, 0, 0,
              1, _stream())
    torch.cuda.synchronize()
    assert _rel(C.float().cpu(), pre_ref - bias + C0.float()) < 8e-3


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 64, 256), (4096, 128, 512), (333, 1024, 2048), (2048, 512, 64)])
def test_tc_gemm_dgrad(M, N, K):
    """dX[M,N=Ci] = dY[M,K=Co] . W[Co,Ci] with the GELU-derivative epilogue."""
    L = lib()
    dY = torch.randn(M, K, generator=_g(1)).bfloat16()
    W = (torch.randn(K, N, generator=_g(2)) / K ** 0.5).bfloat16()
    aux = torch.randn(M, N, generator=_g(3)).bfloat16()
    x = aux.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    ref = (dY.float() @ W.float()) * x.grad
    dYd, Wd, auxd = dY.cuda(), W.cuda(), aux.cuda()
    C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    assert L.cdll.dsgan_tc_gemm_supported(1, M, N, K, K, N, N) == 1
    L.tc_gemm(1, dYd.data_ptr(), K, Wd.data_ptr(), N, M, N, K, C.data_ptr(), N, None, None, 0, auxd.data_ptr(), N, 0, 3,
              0, _stream())
    torch.cuda.synchronize()
    assert _rel(C.float().cpu(), ref) < 8e-3


@pytest.mark.parametrize("P,Co,Ci", [(256, 128, 64), (4096, 64, 128), (10000, 256, 256), (65536, 512, 128),
                                     (1000, 128, 1024), (33000, 64, 64)])
def test_tc_wgrad(P, Co, Ci):
    L = lib()
    dY = torch.randn(P, Co, generator=_g(1)).bfloat16()
    X = torch.randn(P, Ci, generator=_g(2)).bfloat16()
    ref = dY.float().t() @ X.float()
    dW0 = torch.randn(Co, Ci, generator=_g(3))
    dW, dYd, Xd = dW0.cuda(), dY.cuda(), X.cuda()   # keep the device operands alive across the launch
    L.tc_wgrad(dYd.data_ptr(), Co, Xd.data_ptr(), Ci, P, Co, Ci, dW.data_ptr(), Ci, _stream())
    torch.cuda.synchronize()
    assert _rel(dW.cpu() - dW0, ref) < 2e-3
