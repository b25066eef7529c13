"""Diagnostic (GPU box): end-to-end deviation of the CUDA step from the fp32 oracle, by precision and bias init."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dsgan_oracle as O  # noqa: E402
from dsgan_b200.models import create_model  # noqa: E402
from dsgan_b200.options.train_options import TrainOptions  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    torch.set_num_threads(os.cpu_count())
    n, hw = int(sys.argv[1]), int(sys.argv[2])
    for bias in (0.0, 0.05):
        PG, PD, PV = O.init_params_G(20, bias), O.init_params_D(20, bias), O.init_params_vgg(20, bias)
        A, B = O.synthetic_pair(n, hw, hw, seed=1)
        t0 = time.time()
        ref = O.train_step(PG, PD, PV, A, B)
        print("oracle step %.1fs" % (time.time() - t0), flush=True)
        for prec in ("fp32", "bf16"):
            opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_diag", argv=["--precision", prec], quiet=True)
            model = create_model(opt)
            model.setup(opt)
            model.netG.load_state_dict(PG); model.netD.load_state_dict(PD); model.vgg.load_state_dict(PV, strict=False)
            model.set_input({"A": A, "B": B, "A_paths": [""], "B_paths": [""]})
            torch.cuda.synchronize(); t0 = time.time()
            model.optimize_parameters()
            torch.cuda.synchronize(); dt = time.time() - t0
            got = {k: float(getattr(model, "tv_loss" if k == "tv" else "loss_" + k)) for k in ref["losses"]}
            print("== bias %.2f %s  step %.3fs  fake_B rel %.2e" % (bias, prec, dt, rel(model.fake_B.cpu(), ref["fake_B"])))
            print("   losses abs err:", {k: "%.1e" % abs(got[k] - ref["losses"][k]) for k in got})
            for nm, net, gr in (("D", model.netD, ref["grads_D"]), ("G", model.netG, ref["grads_G"])):
                P = net.flat_buffers()[2]
                gmax = max(float(g.norm()) for g in gr.values())
                errs = sorted(((rel(P[k].grad.cpu().reshape(g.shape), g), k) for k, g in gr.items()
                               if float(g.norm()) > 2e-3 * gmax), reverse=True)
                tot = rel(torch.cat([P[k].grad.cpu().flatten() for k in gr]), torch.cat([g.flatten() for g in gr.values()]))
                print("   %s grads: whole-net rel %.2e; worst tensors %s" % (nm, tot, [(k, "%.1e" % e) for e, k in errs[:4]]))
            del model
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
