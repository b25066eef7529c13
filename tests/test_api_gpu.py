"""Reference-facing API on the GPU: Vgg16's 5-tuple, checkpoint interchange (incl. a file written by the reference's own
save_networks), test()/get_img_*() between training steps in CUDA-graph mode, update_learning_rate, GANLoss modes."""
import json
import os
import shutil

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import dsgan_oracle as O  # noqa: E402
from gpu_util import ctx_for, rel  # noqa: E402
from dsgan_b200 import losses  # noqa: E402
from dsgan_b200.models import create_model, networks  # noqa: E402
from dsgan_b200.models.vgg import Vgg16  # noqa: E402
from dsgan_b200.options.train_options import TrainOptions  # noqa: E402


def _model(extra=(), path="/tmp/dsgan_b200_api"):
    opt = TrainOptions().parse("/tmp/none", path, argv=list(extra), quiet=True)
    m = create_model(opt)
    m.setup(opt)
    return m, opt


def test_vgg16_returns_the_reference_5_tuple():
    """vgg.py:30-42: (relu1_2, relu2_2, relu3_3, relu4_3, relu5_3) NCHW feature maps."""
    networks.KernelNet.precision = "fp32"
    try:
        v = Vgg16().cuda()
        PV = {k: p.detach().cpu() for k, p in v.state_dict().items()}
        x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(1))
        outs = v(x.cuda())
    finally:
        networks.KernelNet.precision = "bf16"
    assert isinstance(outs, tuple) and len(outs) == 5
    want4 = O.vgg_forward(PV, x)
    h = F.max_pool2d(want4[3], 2)
    for i in (24, 26, 28):
        h = F.relu(F.conv2d(h, PV["to_relu_5_3.%d.weight" % i], PV["to_relu_5_3.%d.bias" % i], padding=1))
    for got, want in zip(outs, list(want4) + [h]):
        assert got.shape == want.shape
        assert rel(got.cpu(), want) < 1e-4
    assert outs[4].shape == (1, 512, 4, 4)


def test_save_load_networks_round_trip(tmp_path):
    """save_networks writes `<epoch>_useSE_net_X.pth` with `module.`-prefixed keys (base_model.py:92-103, Q16/Q17);
    load_networks of a fresh model restores the exact parameters."""
    m, opt = _model(path=str(tmp_path))
    with torch.no_grad():
        for net in (m.netG, m.netD):
            for p in net.parameters():
                p.add_(0.01)
    m.save_networks(3)
    for name in ("G", "D"):
        f = os.path.join(m.save_dir, "3_useSE_net_%s.pth" % name)
        sd = torch.load(f, map_location="cpu")
        assert all(k.startswith("module.") for k in sd)
    m2, _ = _model(path=str(tmp_path))
    m2.load_networks(3)
    for a, b in ((m.netG, m2.netG), (m.netD, m2.netD)):
        for (ka, pa), (kb, pb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert ka == kb and torch.equal(pa.cpu(), pb.cpu()), ka
    # continue_train path: setup() loads `which_epoch`
    opt3 = TrainOptions().parse("/tmp/none", str(tmp_path), argv=["--continue_train", "--which_epoch", "3"], quiet=True)
    m3 = create_model(opt3)
    m3.setup(opt3)
    assert torch.equal(m3.netG.state_dict()["res.weight"].cpu(), m.netG.state_dict()["res.weight"].cpu())


def test_load_checkpoint_written_by_the_reference(tmp_path, golden_dir):
    """tests/golden/ref_saved_5_useSE_net_D.pth was written by the reference's BaseModel.save_networks
    (oracle/make_golden_api.py, ndf=8).  load_networks must restore every tensor."""
    rec = json.load(open(os.path.join(golden_dir, "api.json")))
    m, opt = _model(["--ndf", "8"], path=str(tmp_path))
    os.makedirs(m.save_dir, exist_ok=True)
    shutil.copy(os.path.join(golden_dir, "ref_saved_5_useSE_net_D.pth"), os.path.join(m.save_dir, "5_useSE_net_D.pth"))
    m.model_names = ["D"]
    m.load_networks(5)
    sd = m.netD.state_dict()
    assert list(sd.keys()) == list(rec["D_fingerprints"].keys())
    for k, fp in rec["D_fingerprints"].items():
        got = O.fingerprint(sd[k].cpu())
        assert all(abs(a - b) <= 1e-6 * max(1.0, abs(b)) for a, b in zip(got, fp)), k
    # ... and the loaded discriminator runs
    y = m.netD(torch.randn(1, 6, 64, 64).cuda())
    assert y.shape == (1, 1, 6, 6) and torch.isfinite(y).all()


def test_visual_helpers_between_graph_steps_do_not_disturb_training():
    """train.py:107-118 calls get_img_tir / get_img_gen (= test(): an extra eager G forward that re-binds fake_B) /
    get_img_label after EVERY optimize_parameters().  In CUDA-graph mode the next replay must still pool and show the
    captured forward's fake_B: the loss trajectory equals the one of a run that never calls the helpers."""
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    batches = [O.synthetic_pair(2, 64, 64, seed=30 + i) for i in range(6)]
    traj = {}
    for helpers in (False, True):
        m, _ = _model(["--precision", "fp32", "--cuda_graph", "1"])
        m.netG.load_state_dict(PG)
        m.netD.load_state_dict(PD)
        m.vgg.load_state_dict(PV, strict=False)
        rows = []
        for A, B in batches:
            data = {"A": A, "B": B, "A_paths": [""], "B_paths": [""]}
            m.set_input(data)
            m.optimize_parameters()
            torch.cuda.synchronize()
            rows.append([float(m._loss[i]) for i in range(7)])
            if helpers:
                tir, gen, lab = m.get_img_tir(data), m.get_img_gen(data), m.get_img_label(data)
                assert gen.shape == tir.shape == lab.shape == (2, 3, 64, 64)
                assert float(tir.min()) >= 0.0 and float(tir.max()) <= 255.0
                vis = m.get_current_visuals()
                assert set(vis) == {"real_A", "fake_B", "real_B"}
        assert m._gs is not None and m._gs["plan"] is not None
        traj[helpers] = rows
    for step, (a, b) in enumerate(zip(traj[False], traj[True])):
        for u, v in zip(a, b):
            assert abs(u - v) <= 1e-2 * max(1.0, abs(u)), (step, a, b)   # run-to-run spread (fp32 atomics), see the prefetch test


def test_test_and_get_img_gen_match_forward():
    m, _ = _model(["--precision", "fp32"])
    A, B = O.synthetic_pair(1, 64, 64, seed=4)
    data = {"A": A, "B": B, "A_paths": ["a"], "B_paths": ["b"]}
    m.set_input(data)
    m.test()
    want = O.g_forward({k: v.detach().cpu() for k, v in m.netG.state_dict().items()}, A)
    assert rel(m.fake_B.cpu(), want.detach()) < 1e-4
    assert m.get_image_paths() == ["a"]
    gen = m.get_img_gen(data)
    assert rel(gen.cpu(), (want.detach() + 1) / 2 * 255) < 1e-4
    assert not m.ctxG.tape, "test() must not leave a tape behind"


def test_update_learning_rate_follows_the_lambda_rule(capsys):
    """networks.py:33-39 / base_model.py:68-72 with niter=10, niter_decay=10: constant for 10 epochs, then linear."""
    m, opt = _model()
    lrs = []
    for _ in range(20):
        m.update_learning_rate()
        lrs.append(m.optimizer_G.param_groups[0]["lr"])
    want = [2e-4 * (1.0 - max(0, e + 1 + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)) for e in range(1, 21)]
    assert all(abs(a - b) < 1e-12 for a, b in zip(lrs, want))
    assert m.optimizer_D.param_groups[0]["lr"] == lrs[-1]
    assert "learning rate" in capsys.readouterr().out


@pytest.mark.parametrize("mode", ["bce", "mse", "sigmoid_mse"])
@pytest.mark.parametrize("real", [True, False])
def test_gan_loss_modes(mode, real):
    """GANLoss (networks.py:143-163): BCE-with-logits (default), MSE (--no_lsgan, inverted flag Q12) and MSE on a
    sigmoid discriminator (use_sigmoid=True, networks.py:571-572): value and gradient w.r.t. the logits."""
    ctx = ctx_for("fp32")
    x = torch.randn(2, 1, 30, 30, generator=torch.Generator().manual_seed(5))
    xr = x.clone().requires_grad_(True)
    t = torch.full_like(x, 1.0 if real else 0.0)
    if mode == "bce":
        want = F.binary_cross_entropy_with_logits(xr, t)
    elif mode == "mse":
        want = F.mse_loss(xr, t)
    else:
        want = F.mse_loss(torch.sigmoid(xr), t)
    want.backward()
    from gpu_util import to_var, var_grad
    v = to_var(ctx, x)
    slot = torch.zeros(1, device="cuda")
    losses.gan_loss(ctx, v, real, slot.data_ptr(), 1.0, 1.0, use_lsgan=mode != "bce", sigmoid_d=mode == "sigmoid_mse")
    assert abs(float(slot) - float(want)) < 1e-5
    assert rel(var_grad(v), xr.grad) < 1e-4


def test_batch_metrics_reuse_the_step_outputs():
    """N4: SSIM / PSNR from the tensors of the step that just ran (no extra generator forward)."""
    import math
    from dsgan_b200._lib import lib
    m, _ = _model(["--precision", "fp32", "--cuda_graph", "0"])
    A, B = O.synthetic_pair(2, 64, 64, seed=8)
    m.set_input({"A": A, "B": B, "A_paths": [""], "B_paths": [""]})
    m.optimize_parameters()
    n0 = lib().cdll.dsgan_launch_count()
    met = m.batch_metrics()
    assert lib().cdll.dsgan_launch_count() == n0, "batch_metrics must not launch any network kernel"
    fake, real = m.fake_B.cpu(), B
    want_ssim = float(O.ssim((real + 1) / 2, (fake + 1) / 2, 1.0))
    assert abs(float(met["ssim"]) - want_ssim) < 1e-4
    mse = float((((fake.clamp(-1, 1) + 1) * 127.5 - (real + 1) * 127.5) ** 2).mean())
    assert abs(float(met["psnr"]) - 10 * math.log10(255 ** 2 / mse)) < 1e-3


def test_prefetch_input_is_equivalent_to_set_input():
    """prefetch_input() stages the next batch on a copy stream; set_input() of the same tensors takes the staged copy, any other
    batch falls back to the direct copy.  The loss trajectory (CUDA-graph mode) equals the one without prefetching."""
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    batches = [O.synthetic_pair(2, 64, 64, seed=60 + i) for i in range(5)]
    pinned = [{"A": A.pin_memory(), "B": B.pin_memory(), "A_paths": [""], "B_paths": [""]} for A, B in batches]
    traj = {}
    for prefetch in (False, True):
        m, _ = _model(["--precision", "fp32", "--cuda_graph", "1"])
        m.netG.load_state_dict(PG)
        m.netD.load_state_dict(PD)
        m.vgg.load_state_dict(PV, strict=False)
        rows = []
        for i, data in enumerate(pinned):
            m.set_input(data)
            assert torch.equal(m.real_A.cpu(), data["A"]) and torch.equal(m.real_B.cpu(), data["B"])
            m.optimize_parameters()
            if prefetch and i + 1 < len(pinned):
                # stage the wrong batch first on odd steps: set_input must then ignore it and copy directly
                m.prefetch_input(pinned[i + 1] if i % 2 == 0 else pinned[0])
            torch.cuda.synchronize()
            rows.append([float(m._loss[j]) for j in range(7)])
        traj[prefetch] = rows
    # the inputs are compared bit for bit above; the trajectories only have to agree to the run-to-run spread of two identical
    # runs (fp32 atomics in the weight gradients: a few 1e-3 on the D losses after five steps)
    for a, b in zip(traj[False], traj[True]):
        for u, v in zip(a, b):
            assert abs(u - v) <= 1e-2 * max(1.0, abs(u)), (a, b)
