"""Frozen VGG16 feature extractor shell (reference: DSGAN/models/vgg.py:5-42).  Parameters keep the reference's
names (`to_relu_1_2.0.weight` ...).  The reference downloads ImageNet weights (vgg.py:8); there is no network
here, so weights are random-initialised like torchvision's `weights=None` and can be overwritten with
load_state_dict.  Only the four taps the loss reads are computed (relu5_3 is dead work in the reference, Q14)."""
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

    def forward_var(self, x: Var, need_dx=True):
        return nets.vgg_forward(self.ctx(), self.params(), x, need_dx)

    def forward(self, x):
        """N x 3 x H x W fp32 -> 4 NCHW fp32 feature maps (relu1_2, relu2_2, relu3_3, relu4_3)."""
        ctx = self.ctx()
        was, ctx.no_grad = ctx.no_grad, True
        try:
            taps = self.forward_var(image_to_nhwc(ctx, x.contiguous().float()), need_dx=False)
        finally:
            ctx.no_grad = was
        outs = []
        for t in taps:
            o = torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32, device=x.device)
            ctx.L.nhwc_to_nchw(t.ptr, ctx.dt, t.ld, o.data_ptr(), t.N, t.C, t.H, t.W, 1.0, 0, ctx.stream)
            outs.append(o)
        return tuple(outs)
