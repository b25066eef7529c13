"""Generate tests/golden/*.json by running the UNMODIFIED reference (/root/reference/DSGAN) on CPU.

Runs only in the build container (the reference does not travel to the GPU box).  Nothing is
copied from the reference: it is imported, fed the oracle's deterministic weights/inputs, executed,
and its outputs are stored as small fingerprints.

Shims (none change arithmetic):
  * stub modules for import-only dependencies that are not installed (pytorch_msssim, pytorch_ssim,
    skimage.metrics) — used only by off-path helpers;
  * torchvision vgg16(pretrained=True) -> weights=None (no network); weights are then overwritten with
    the oracle's seeded VGG init;
  * Vgg16().type(torch.cuda.FloatTensor) -> .float() (hard-CUDA call, pix2pix_model.py:118).

usage:  python oracle/make_golden.py            # writes tests/golden/
"""
import argparse
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dsgan_oracle as O  # noqa: E402

REF = "/root/reference/DSGAN"


def import_reference():
    sys.path.insert(0, REF)
    for name in ("pytorch_msssim", "pytorch_ssim"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    try:
        import skimage.metrics  # noqa: F401
    except Exception:
        sk = types.ModuleType("skimage")
        skm = types.ModuleType("skimage.metrics")
        skm.peak_signal_noise_ratio = skm.structural_similarity = None
        sk.metrics = skm
        sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
    import torchvision.models as tvm
    _orig = tvm.vgg16
    tvm.vgg16 = lambda pretrained=False, **kw: _orig(weights=None)
    import models.vgg as refvgg
    refvgg.Vgg16.type = lambda self, *_a, **_k: self.float()
    import models.pix2pix_model as p2p
    import MS_SSIM as ref_msssim
    return p2p, ref_msssim


def make_opt(tmpdir):
    return argparse.Namespace(
        gpu_ids=[], isTrain=True, checkpoints_dir=tmpdir, name="golden", resize_or_crop="resize_and_crop",
        use_GAN=1, w_gan=0.01, w_vgg=1, w_tv=1, w_ss=1.25, use_condition=1, input_nc=3, output_nc=3, ngf=32,
        ndf=32, which_model_netG="MixConvNeXtML", which_model_netD="basic", n_layers_D=3, norm="instance",
        no_dropout=False, init_type="normal", no_lsgan=False, pool_size=50, lr=2e-4, beta1=0.5,
        which_direction="AtoB", lr_policy="lambda", epoch_count=1, niter=10, niter_decay=10,
        continue_train=False, verbose=False)


def run_step(p2p, n, hw, bias_std, seed_in):
    torch.manual_seed(0)
    model = p2p.Pix2PixModel()
    model.initialize(make_opt("/tmp/dsgan_golden"))
    PG, PD, PV = O.init_params_G(20, bias_std), O.init_params_D(20, bias_std), O.init_params_vgg(20, bias_std)
    assert list(model.netG.state_dict().keys()) == list(PG.keys()), "G state_dict names/order differ"
    assert list(model.netD.state_dict().keys()) == list(PD.keys()), "D state_dict names/order differ"
    model.netG.load_state_dict(PG, strict=True)
    model.netD.load_state_dict(PD, strict=True)
    missing = model.vgg.load_state_dict(PV, strict=False)
    assert not missing.unexpected_keys and all("to_relu_5_3" in k for k in missing.missing_keys), missing
    A, B = O.synthetic_pair(n, hw, hw, seed=seed_in)
    model.set_input({"A": A, "B": B, "A_paths": [""], "B_paths": [""]})
    model.optimize_parameters()
    rec = {
        "n": n, "hw": hw, "bias_std": bias_std, "seed_in": seed_in,
        "losses": {k: float(getattr(model, a)) for k, a in (
            ("G_GAN", "loss_G_GAN"), ("G_L1", "loss_G_L1"), ("vgg", "loss_vgg"), ("tv", "tv_loss"),
            ("ssim", "loss_ssim"), ("G", "loss_G"), ("D_fake", "loss_D_fake"), ("D_real", "loss_D_real"),
            ("D", "loss_D"))},
        "fake_B": O.fingerprint(model.fake_B),
        "fake_B_head": model.fake_B.flatten()[:16].tolist(),
        "grads_G": {k: O.fingerprint(p.grad) for k, p in model.netG.named_parameters()},
        "grads_D": {k: O.fingerprint(p.grad) for k, p in model.netD.named_parameters()},
        "new_G": {k: O.fingerprint(p) for k, p in model.netG.named_parameters()},
        "new_D": {k: O.fingerprint(p) for k, p in model.netD.named_parameters()},
    }
    return rec


def run_msssim(ref, n, hw, seed, noise):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, 3, hw, hw, generator=g)
    if noise is None:
        Y = torch.rand(n, 3, hw, hw, generator=g)
    else:
        Y = (X + noise * torch.randn(n, 3, hw, hw, generator=g)).clamp(0, 1)
    rec = {"n": n, "hw": hw, "seed": seed, "noise": noise}
    for fn in ("ssim", "ms_ssim"):
        if fn == "ms_ssim" and hw <= 160:
            continue
        Yr = Y.clone().requires_grad_(True)
        v = getattr(ref, fn)(X, Yr, data_range=1, size_average=True)
        v.backward()
        rec[fn] = float(v)
        rec[fn + "_grad"] = O.fingerprint(Yr.grad)
        rec[fn + "_per_image"] = getattr(ref, fn)(X, Y, data_range=1, size_average=False).tolist()
    return rec


def main():
    out = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out, exist_ok=True)
    p2p, ref_msssim = import_reference()
    torch.set_num_threads(os.cpu_count())
    if "--skip-steps" not in sys.argv:
        steps = [run_step(p2p, 2, 32, 0.05, 1), run_step(p2p, 1, 64, 0.0, 2), run_step(p2p, 1, 256, 0.05, 1),
                 run_step(p2p, 1, 256, 0.0, 1)]  # the last one is BASELINE config #1 with the reference's own init (bias 0)
        with open(os.path.join(out, "train_step.json"), "w") as f:
            json.dump(steps, f)
    ms = [run_msssim(ref_msssim, 2, 256, 3, 0.1), run_msssim(ref_msssim, 2, 256, 4, None),
          run_msssim(ref_msssim, 3, 48, 5, 0.2), run_msssim(ref_msssim, 1, 176, 6, 0.05),
          run_msssim(ref_msssim, 1, 200, 7, 0.05)]   # 200 -> 100 -> 50 -> 25 -> 13: odd pyramid levels (padded avg_pool, :214-216)
    with open(os.path.join(out, "ms_ssim.json"), "w") as f:
        json.dump(ms, f)
    win = ref_msssim._fspecial_gauss_1d(11, 1.5).flatten().tolist()
    with open(os.path.join(out, "gauss_window.json"), "w") as f:
        json.dump(win, f)
    print("golden written to", out)


if __name__ == "__main__":
    main()
