"""CPU emulation (test infrastructure): where does bf16 STORAGE error enter the training step?

Runs the fp32 oracle with a bf16 round-trip (forward AND backward, straight-through) inserted after every op of ONE
region of the generator at a time, and reports the error of fake_B, of D's gradient and of G's gradient against the
unperturbed fp32 step.  Everything else (D, VGG, losses, the other regions) stays fp32, so each row isolates one region.
Used for profiles/r2_error_budget.json ("cpu_emulation").

  python tests/tools/bf16_emulation.py [batch]        # default 2 images of 256x256, ~1 min on 8 cores
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dsgan_oracle as O  # noqa: E402


class Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


ON = [False]


def q(x):
    return Q.apply(x) if ON[0] else x


class _F:
    pass


def install():
    fq = _F()
    for n in dir(F):
        try:
            setattr(fq, n, getattr(F, n))
        except Exception:
            pass
    for n in ("conv2d", "linear", "gelu", "instance_norm", "conv_transpose2d", "max_pool2d"):
        f = getattr(F, n)
        setattr(fq, n, (lambda f: lambda *a, **k: q(f(*a, **k)))(f))
    O.F = fq


def region(fn, name, active):
    def g(*a, **k):
        prev = ON[0]
        ON[0] = name in active
        try:
            return fn(*a, **k)
        finally:
            ON[0] = prev
    return g


ORIG = dict(block=O._block, up=O._upsample, down=O._downskip, local=O._local)


def setup(active):
    O._block = lambda P, p, x, taps=None: region(ORIG["block"], "encoder" if p.startswith("c") else "decoder", active)(P, p, x, taps)
    O._upsample = lambda P, p, x, skip, idx="model.0": region(ORIG["up"], "local" if p.startswith("local") else "decoder",
                                                              active)(P, p, x, skip, idx)
    O._downskip = region(ORIG["down"], "downskips", active)
    O._local = region(ORIG["local"], "local", active)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def cat(d, keys):
    return torch.cat([d[k].flatten() for k in keys])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.set_num_threads(os.cpu_count())
    install()
    PG, PD, PV = O.init_params_G(20, 0.0), O.init_params_D(20, 0.0), O.init_params_vgg(20, 0.0)
    A, B = O.synthetic_pair(n, 256, 256, seed=1)
    setup(set())
    ref = O.train_step(PG, PD, PV, A, B, update=False)
    kG, kD = list(ref["grads_G"]), list(ref["grads_D"])
    nz = [k for k in kG if not (k.endswith(".bias") and float(ref["grads_G"][k].norm()) < 1e-3)]
    rows = []
    for act in (["encoder"], ["downskips"], ["decoder"], ["local"], ["encoder", "downskips", "decoder", "local"]):
        setup(set(act))
        r = O.train_step(PG, PD, PV, A, B, update=False)
        rows.append({"bf16_region": "+".join(act), "fake_B": rel(r["fake_B"], ref["fake_B"]),
                     "grads_D": rel(cat(r["grads_D"], kD), cat(ref["grads_D"], kD)),
                     "grads_G": rel(cat(r["grads_G"], kG), cat(ref["grads_G"], kG)),
                     "grads_G_without_structurally_zero_biases": rel(cat(r["grads_G"], nz), cat(ref["grads_G"], nz))})
        print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    main()
