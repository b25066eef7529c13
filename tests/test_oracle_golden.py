"""CPU: the oracle (oracle/dsgan_oracle.py) against fixtures produced by the reference itself
(oracle/make_golden.py, run in the build container against /root/reference)."""
import json
import os

import pytest
import torch

import dsgan_oracle as O


def _close(a, b, rtol, atol):
    return abs(a - b) <= atol + rtol * abs(b)


def _fp_close(got, want, rtol=2e-3):
    # fingerprint = [l2 norm, sum, projection]; sum/projection are compared at the scale of the norm
    scale = max(abs(want[0]), 1e-12)
    return (abs(got[0] - want[0]) <= rtol * scale and abs(got[1] - want[1]) <= rtol * scale * 50
            and abs(got[2] - want[2]) <= rtol * scale * 50)


def test_gauss_window(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "gauss_window.json")))
    got = O.gauss_window().tolist()
    assert got == pytest.approx(want, abs=1e-7)
    assert got[:3] == pytest.approx([0.0010284, 0.0075988, 0.0360008], abs=1e-6)  # SURVEY §8c anchor


def test_param_inventory():
    g = O.init_params_G()
    assert len(g) == 188
    n = sum(v.numel() for v in g.values())
    assert abs(n / 1e6 - 22.425) < 0.001, n  # 22.425 M
    d = O.init_params_D()
    assert len(d) == 10 and abs(sum(v.numel() for v in d.values()) / 1e6 - 0.696) < 0.001
    assert g["u1.model.0.weight"].shape == (1024, 512, 3, 3) and g["c1.pwconv1.weight"].shape == (12, 3)


@pytest.mark.parametrize("case", [0, 1, 4])   # case 4: 200x200 -> odd pyramid levels (padded avg_pool, MS_SSIM.py:214-216)
def test_ms_ssim_golden(golden_dir, case):
    rec = json.load(open(os.path.join(golden_dir, "ms_ssim.json")))[case]
    g = torch.Generator().manual_seed(rec["seed"])
    n, hw = rec["n"], rec["hw"]
    X = torch.rand(n, 3, hw, hw, generator=g)
    Y = torch.rand(n, 3, hw, hw, generator=g) if rec["noise"] is None else \
        (X + rec["noise"] * torch.randn(n, 3, hw, hw, generator=g)).clamp(0, 1)
    for fn in ("ssim", "ms_ssim"):
        Yr = Y.clone().requires_grad_(True)
        v = getattr(O, fn)(X, Yr, 1.0)
        v.backward()
        assert _close(float(v), rec[fn], 1e-5, 1e-6), (fn, float(v), rec[fn])
        assert _fp_close(O.fingerprint(Yr.grad), rec[fn + "_grad"], 1e-4)
        per = getattr(O, fn)(X, Y, 1.0, size_average=False).tolist()
        assert per == pytest.approx(rec[fn + "_per_image"], rel=1e-5, abs=1e-6)


def test_ssim_small_and_identity(golden_dir):
    rec = json.load(open(os.path.join(golden_dir, "ms_ssim.json")))[2]
    g = torch.Generator().manual_seed(rec["seed"])
    X = torch.rand(rec["n"], 3, rec["hw"], rec["hw"], generator=g)
    Y = (X + rec["noise"] * torch.randn(X.shape, generator=g)).clamp(0, 1)
    assert _close(float(O.ssim(X, Y, 1.0)), rec["ssim"], 1e-5, 1e-6)
    assert float(O.ssim(X, X, 1.0)) == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(AssertionError):
        O.ms_ssim(X, Y, 1.0)  # 48 <= 160, MS_SSIM.py:194-197
    Z = torch.rand(1, 3, 176, 176)
    assert float(O.ms_ssim(Z, Z, 1.0)) == pytest.approx(1.0, abs=1e-5)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_train_step_golden(golden_dir, case):
    rec = json.load(open(os.path.join(golden_dir, "train_step.json")))[case]
    torch.set_num_threads(os.cpu_count())
    bs = rec["bias_std"]
    PG, PD, PV = O.init_params_G(20, bs), O.init_params_D(20, bs), O.init_params_vgg(20, bs)
    A, B = O.synthetic_pair(rec["n"], rec["hw"], rec["hw"], seed=rec["seed_in"])
    out = O.train_step(PG, PD, PV, A, B)
    for k, want in rec["losses"].items():
        assert _close(out["losses"][k], want, 2e-4, 2e-5), (k, out["losses"][k], want)
    assert _fp_close(O.fingerprint(out["fake_B"]), rec["fake_B"], 1e-4)
    assert out["fake_B"].flatten()[:16].tolist() == pytest.approx(rec["fake_B_head"], rel=1e-3, abs=1e-4)
    bad = [k for k, want in rec["grads_D"].items() if not _fp_close(O.fingerprint(out["grads_D"][k]), want)]
    assert not bad, ("D grads", bad)
    # biases that feed an InstanceNorm have a mathematically zero gradient: what both sides hold there is
    # rounding noise, so tensors whose norm is negligible next to the largest gradient are exempt.
    gmax = max(w[0] for w in rec["grads_G"].values())
    bad = [k for k, want in rec["grads_G"].items()
           if want[0] > 2e-3 * gmax and not _fp_close(O.fingerprint(out["grads_G"][k]), want, 5e-3)]
    assert not bad, ("G grads", bad)
    # Adam's first step is lr*sign(g): tensors whose gradient is rounding noise move by +-lr at random
    bad = [k for k, want in rec["new_G"].items()
           if rec["grads_G"][k][0] > 2e-3 * gmax and not _fp_close(O.fingerprint(out["PG"][k]), want, 1e-3)]
    assert not bad, ("Adam G", bad)
    bad = [k for k, want in rec["new_D"].items() if not _fp_close(O.fingerprint(out["PD"][k]), want, 1e-3)]
    assert not bad, ("Adam D", bad)
