"""Per-network error budget of one optimize_parameters() on the real kernels (test infrastructure, needs a B200).

For every (G, D, VGG) precision assignment it runs ONE training step at the given batch/size with the reference's
init and the seeded synthetic pair, and records the relative error of fake_B, of D's flat gradient and of G's flat
gradient (whole net and per parameter group) against the fp32 CPU oracle.  Output: one JSON document.

  python tests/tools/error_budget.py --batch 16 --out profiles/r2_error_budget.json
"""
import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import dsgan_oracle as O  # noqa: E402
from dsgan_b200.models import create_model  # noqa: E402
from dsgan_b200.options.train_options import TrainOptions  # noqa: E402

GROUPS = {"encoder": ("c1.", "c2.", "c3.", "c4.", "c5."), "downskips": ("down",),
          "decoder": ("u1.", "u2.", "u3.", "u4.", "uc1.", "uc2.", "uc3.", "uc4."), "local": ("local.",), "res": ("res.",)}


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cat(d, keys):
    return torch.cat([d[k].detach().flatten().float().cpu() for k in keys])


def run(prec, n, hw, ref, PG, PD, PV, A, B, extra=()):
    argv = ["--precision", "bf16", "--cuda_graph", "0", "--precision_G", prec[0], "--precision_D", prec[1],
            "--precision_vgg", prec[2]] + list(extra)
    opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_b200_budget", argv=argv, quiet=True)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
        model.setup(opt)
    model.netG.load_state_dict(PG)
    model.netD.load_state_dict(PD)
    model.vgg.load_state_dict(PV, strict=False)
    model.set_input({"A": A, "B": B, "A_paths": [""], "B_paths": [""]})
    model.optimize_parameters()
    torch.cuda.synchronize()
    PDm, PGm = model.netD.flat_buffers()[2], model.netG.flat_buffers()[2]
    gD = {k: PDm[k].grad.cpu().reshape(g.shape) for k, g in ref["grads_D"].items()}
    gG = {k: PGm[k].grad.cpu().reshape(g.shape) for k, g in ref["grads_G"].items()}
    out = {"G": prec[0], "D": prec[1], "vgg": prec[2],
           "fake_B": rel(model.fake_B.cpu(), ref["fake_B"]),
           "grads_D": rel(cat(gD, list(gD)), cat(ref["grads_D"], list(gD))),
           "grads_G": rel(cat(gG, list(gG)), cat(ref["grads_G"], list(gG))),
           "losses_abs": max(abs(float(getattr(model, "tv_loss" if k == "tv" else "loss_" + k)) - w) / max(1.0, abs(w))
                             for k, w in ref["losses"].items())}
    for gname, pre in GROUPS.items():
        keys = [k for k in gG if k.startswith(pre)]
        out["grads_G_" + gname] = rel(cat(gG, keys), cat(ref["grads_G"], keys))
        out["norm_G_" + gname] = float(cat(ref["grads_G"], keys).norm())
    for k in ("model.0.weight", "model.2.weight", "model.5.weight", "model.8.weight", "model.11.weight"):
        out["grads_D_" + k] = rel(gD[k], ref["grads_D"][k])
    if TOP:
        for tag, mine, want in (("G", gG, ref["grads_G"]), ("D", gD, ref["grads_D"])):
            tot = float(cat(want, list(want)).norm())
            rows = sorted(((float((mine[k] - want[k]).norm()), float(want[k].norm()), k) for k in want), reverse=True)
            out["top_err_" + tag] = [{"name": k, "abs_err_over_total_norm": e / tot, "rel": e / max(n, 1e-30),
                                      "norm_over_total": n / tot} for e, n, k in rows[:TOP]]
    del model
    torch.cuda.empty_cache()
    return out


TOP = 0


def main():
    global TOP
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=0, help="also list the N tensors with the largest absolute gradient error")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--out", default="")
    ap.add_argument("--configs", default="bbb,fbb,bfb,bbf,bff,fbf,ffb,fff")
    args = ap.parse_args()
    TOP = args.top
    torch.set_num_threads(os.cpu_count())
    PG, PD, PV = O.init_params_G(20, 0.0), O.init_params_D(20, 0.0), O.init_params_vgg(20, 0.0)
    A, B = O.synthetic_pair(args.batch, args.hw, args.hw, seed=1)
    ref = O.train_step(PG, PD, PV, A, B)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ac = O.train_step(PG, PD, PV, A, B, update=False)
    kD, kG = list(ref["grads_D"]), list(ref["grads_G"])
    doc = {"what": "relative error vs the fp32 CPU oracle after ONE optimize_parameters(); b = bf16 tcgen05 kernels, "
                   "f = fp32 validation kernels; one letter per network in the order G, D, VGG",
           "batch": args.batch, "hw": args.hw,
           "reference_under_torch_bf16_autocast": {
               "fake_B": rel(ac["fake_B"], ref["fake_B"]), "grads_D": rel(cat(ac["grads_D"], kD), cat(ref["grads_D"], kD)),
               "grads_G": rel(cat(ac["grads_G"], kG), cat(ref["grads_G"], kG))},
           "rows": []}
    m = {"b": "bf16", "f": "fp32"}
    for c in args.configs.split(","):
        row = run((m[c[0]], m[c[1]], m[c[2]]), args.batch, args.hw, ref, PG, PD, PV, A, B)
        row["config"] = c
        print(json.dumps({k: v for k, v in row.items() if not k.startswith("top_err")}), flush=True)
        for tag in ("G", "D"):
            for t in row.get("top_err_" + tag, []):
                print("   %s %-34s err/|g| %.2e  rel %.2e  |g_k|/|g| %.2e" % (tag, t["name"], t["abs_err_over_total_norm"],
                                                                          t["rel"], t["norm_over_total"]), flush=True)
        doc["rows"].append(row)
    if args.out:
        json.dump(doc, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
