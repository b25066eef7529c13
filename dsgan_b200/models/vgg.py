"""Frozen VGG16 feature extractor shell (reference: DSGAN/models/vgg.py:5-42).  Parameters keep the reference's
names (`to_relu_1_2.0.weight` ...).  The reference downloads ImageNet weights (vgg.py:8); there is no network
here, so weights are random-initialised like torchvision's `weights=None` and can be overwritten with
load_state_dict.  The training step computes only the four taps the loss reads (relu5_3 is dead work in the reference, Q14); the public
forward() returns the reference's 5-tuple.  `load_torchvision(path)` loads a torchvision vgg16 state_dict (the file
`vgg16(pretrained=True)` downloads: keys `features.N.*`) under the reference's regrouped names."""
import math

import torch

from .. import nets, specs
from ..engine import Var, image_to_nhwc
from .networks import KernelNet


class Vgg16(KernelNet):
    def __init__(self):
        super().__init__(specs.vgg_spec(with_tail=True))
        self.frozen_hint = True  # never updated by the step: packed bf16 operands are reused until a tensor is edited
        for name, p in self.named_parameters():
            p.requires_grad = False
            if name.endswith(".bias"):
                p.data.zero_()
            else:  # kaiming_normal_(mode='fan_out', nonlinearity='relu')
                p.data.normal_(0.0, math.sqrt(2.0 / (p.shape[0] * p.shape[2] * p.shape[3])))

    def forward_var(self, x: Var, need_dx=True, with_tail=False):
        return nets.vgg_forward(self.ctx(), self.params(), x, need_dx, with_tail)

    def load_torchvision(self, path_or_state):
        """Load ImageNet weights: a torchvision vgg16 state_dict (`features.N.weight/bias`, optionally with the classifier
        keys, which are ignored) or a state_dict already under the reference's names (vgg.py:16-25)."""
        sd = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) else path_or_state
        group = {}
        for e in specs.VGG_PLAN + specs.VGG_TAIL:
            if isinstance(e, tuple):
                group[e[0].split(".")[1]] = e[0]
        out = {}
        for k, v in sd.items():
            if k.startswith("features."):
                _f, idx, leaf = k.split(".")
                if idx in group:
                    out["%s.%s" % (group[idx], leaf)] = v
            elif k.split(".")[0].startswith("to_relu_"):
                out[k] = v
        missing = [n for n, _ in self._spec if n not in out]
        if missing:
            raise KeyError("vgg weights: missing %s" % missing[:4])
        self.load_state_dict(out)
        return self

    def forward(self, x):
        """N x 3 x H x W fp32 -> the reference's 5-tuple of NCHW fp32 feature maps (relu1_2, relu2_2, relu3_3, relu4_3,
        relu5_3), vgg.py:30-42."""
        ctx = self.ctx()
        was, ctx.no_grad = ctx.no_grad, True
        try:
            taps = self.forward_var(image_to_nhwc(ctx, x.contiguous().float()), need_dx=False, with_tail=True)
        finally:
            ctx.no_grad = was
        outs = []
        for t in taps:
            o = torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32, device=x.device)
            ctx.L.nhwc_to_nchw(t.ptr, ctx.dt, t.ld, o.data_ptr(), t.N, t.C, t.H, t.W, 1.0, 0, ctx.stream)
            outs.append(o)
        return tuple(outs)
