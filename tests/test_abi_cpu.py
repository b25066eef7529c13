"""CPU: the C-ABI library loads and exports every symbol include/dsgan_b200.h declares; the product path refuses to run
without a B200 (no CPU fallback); option/registry mirrors behave like the reference's."""
import ctypes
import os

import pytest
import torch

from dsgan_b200 import _lib, specs


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 40
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing
    assert cdll.dsgan_abi_version() == 1


def test_struct_mirrors_match_header_sizes():
    # dsgan_conv_desc: 17 ints (+pad) + 4 long long + 3 ints ; dsgan_tc_conv_desc: 19 + 48 + 5 ints
    assert ctypes.sizeof(_lib.ConvDesc) == 120
    assert ctypes.sizeof(_lib.TcConvDesc) == 4 * (14 + 12 + 3 + 192 + 5)
    assert ctypes.sizeof(_lib.TcWgradDesc) == 4 * 11 + 4 * 32 + 4 + 8 * 16 + 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_lib.DsganError):
        _lib.require_device()
    from dsgan_b200.models.networks import MixConvNeXtML
    net = MixConvNeXtML()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 32, 32))
    from dsgan_b200 import MS_SSIM
    with pytest.raises(RuntimeError):
        MS_SSIM.ssim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32), data_range=1)


def test_options_and_registry_mirror_reference_defaults():
    from dsgan_b200 import models
    from dsgan_b200.options.test_options import TestOptions
    from dsgan_b200.options.train_options import TrainOptions
    o = TrainOptions().parse("/data", "/tmp/dsgan_b200_opt_test", argv=[], quiet=True)
    assert (o.batchSize, o.ngf, o.ndf, o.norm, o.which_model_netG, o.which_model_netD) == (1, 32, 32, "instance", "MixConvNeXtML", "basic")
    assert (o.use_GAN, o.w_gan, o.w_vgg, o.w_tv, o.w_ss, o.use_condition) == (1, 0.01, 1, 1, 1.25, 1)
    assert (o.lr, o.beta1, o.pool_size, o.niter, o.niter_decay, o.lr_policy, o.lambda_L1) == (2e-4, 0.5, 50, 10, 10, "lambda", 100.0)
    assert o.gpu_ids == [0] and o.isTrain and o.checkpoints_dir == "/tmp/dsgan_b200_opt_test/checkpoints" and o.dataroot == "/data"
    # the reference's untyped flags arrive as strings when given on the command line (kept)
    o2 = TrainOptions().parse("/data", "/tmp/dsgan_b200_opt_test", argv=["--use_GAN", "1", "--gpu_ids", "0,1"], quiet=True)
    assert o2.use_GAN == "1" and o2.gpu_ids == [0, 1]
    t = TestOptions().parse("/data", "/tmp/dsgan_b200_opt_test", argv=[], quiet=True)
    assert not t.isTrain and not hasattr(t, "lambda_L1")
    assert models.find_model_using_name("pix2pix").__name__ == "Pix2PixModel"
    assert models.find_model_using_name("test").__name__ == "TestModel"
    with pytest.raises(ModuleNotFoundError):
        models.find_model_using_name("cycle_gan")


def test_param_tree_names_alignment_and_state_dict_roundtrip(tmp_path):
    from dsgan_b200.models import networks
    g = networks.MixConvNeXtML().init_normal()
    assert [k for k, _ in specs.generator_spec()] == list(g.state_dict().keys())
    assert all(off % 64 == 0 for off in g._offsets.values())
    sd = {("module." + k): v.clone() for k, v in g.state_dict().items()}   # DataParallel-style keys (Q16)
    torch.save(sd, tmp_path / "1_net_G.pth")
    g2 = networks.MixConvNeXtML()
    loaded = torch.load(tmp_path / "1_net_G.pth")
    g2.load_state_dict({k[7:]: v for k, v in loaded.items()})
    assert all(torch.equal(a, b) for a, b in zip(g.state_dict().values(), g2.state_dict().values()))
    assert float(g.state_dict()["local.mid32.attn.relu1.weight"]) == 0.25
    d = networks.NLayerDiscriminator(6, 32)
    assert [k for k, _ in specs.discriminator_spec()] == list(d.state_dict().keys())


def test_lambda_lr_rule():
    from dsgan_b200.models.networks import LambdaRule
    import argparse

    class Opt:
        param_groups = [{"lr": 2e-4}]
    o = argparse.Namespace(epoch_count=1, niter=10, niter_decay=10, lr_policy="lambda")
    opt = Opt()
    sch = LambdaRule(opt, o)
    lrs = []
    for _ in range(20):
        lrs.append(opt.param_groups[0]["lr"])
        sch.step()
    want = [2e-4 * (1.0 - max(0, e + 1 + 1 - 10) / 11.0) for e in range(20)]   # networks.py:35-37
    assert lrs == pytest.approx(want)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints exactly ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
