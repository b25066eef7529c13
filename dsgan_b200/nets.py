"""The three networks of the hot path as kernel graphs over the engine, plus the nn.Module shells that hold
their parameters under the reference's names.

G: MixConvNeXtML (models/model/MixConvNeXtML.py:428-494)   D: NLayerDiscriminator (networks.py:533-579)
VGG16 taps (models/vgg.py:5-42).  The graphs are hand-scheduled: forward issues the kernels, and records on
the engine tape the backward kernels each layer needs.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import specs
import os

from .engine import (ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, Ctx, Param, Var, add_n, block_mlp, ca_scale, concat_into,
                     conv2d, conv_transpose2d, dwconv, dwconv_multi, fused_mlp_ok, image_to_nhwc, inorm, maxpool, multi_maxpool)


class ParamTree(nn.Module):
    """nn.Module whose parameters carry the reference's dotted names and live as views of one flat fp32 buffer
    (so Adam, the bf16 shadow copy and the NCCL all-reduce each touch a single contiguous range)."""

    def __init__(self, spec):
        super().__init__()
        self._spec = list(spec)
        # every tensor starts on a 64-element boundary of the flat buffers: its bf16 shadow is then 128-byte aligned,
        # which TMA (16 B) and the vectorised kernels need.  The padding stays zero (zero gradient, zero Adam update).
        self._offsets, off = {}, 0
        for name, shape in self._spec:
            self._offsets[name] = off
            off += (math.prod(shape) + 63) // 64 * 64
        self._numel = off
        flat = torch.zeros(self._numel)
        for name, shape in self._spec:
            n = math.prod(shape)
            off = self._offsets[name]
            mod = self
            *path, leaf = name.split(".")
            for part in path:
                if not hasattr(mod, part):
                    mod.add_module(part, nn.Module())
                mod = getattr(mod, part)
            mod.register_parameter(leaf, nn.Parameter(flat[off:off + n].view(shape)))
        self._flat = flat
        self._flat_grad = None
        self._flat_bf16 = None
        self._plist = None
        self.pack_epoch = 0      # bumped whenever the bf16 shadows are re-derived
        self._sig = None         # (tensor versions) at the last refresh; lets frozen networks skip the re-pack

    def init_normal(self, gain=0.02):
        """networks.init_weights('normal') (networks.py:49-70): Conv*/Linear weights N(0,gain), biases 0,
        PReLU slope untouched (0.25)."""
        for name, p in self.named_parameters():
            if name.endswith("relu1.weight"):
                p.data.fill_(0.25)
            elif name.endswith(".bias"):
                p.data.zero_()
            else:
                p.data.normal_(0.0, gain)
        return self

    def flat_buffers(self):
        """(flat params, flat grads, {name: Param}); re-flattens if .to()/.cuda() replaced the storages."""
        params = dict(self.named_parameters())
        first = params[self._spec[0][0]]
        dev = first.device
        ok = self._flat.device == dev and self._flat_grad is not None and self._flat_grad.device == dev
        if ok:
            for name, shape in self._spec:
                if params[name].data_ptr() != self._flat.data_ptr() + 4 * self._offsets[name]:
                    ok = False
                    break
        if not ok:
            flat = torch.zeros(self._numel, dtype=torch.float32, device=dev)
            grad = torch.zeros(self._numel, dtype=torch.float32, device=dev)
            for name, shape in self._spec:
                n, off = math.prod(shape), self._offsets[name]
                flat[off:off + n].copy_(params[name].data.reshape(-1).float())
                params[name].data = flat[off:off + n].view(shape)
                params[name].grad = grad[off:off + n].view(shape)
            self._flat, self._flat_grad, self._plist = flat, grad, None
            self._flat_bf16 = torch.empty(self._numel, dtype=torch.bfloat16, device=dev)
        if self._plist is None:
            base32, base16 = self._flat.data_ptr(), self._flat_bf16.data_ptr()
            self._plist = {n: Param(n, p.data, p.grad, base16 + (p.data_ptr() - base32) // 2, owner=self)
                           for n, p in params.items()}
            self._sig = None
        return self._flat, self._flat_grad, self._plist

    def refresh_bf16(self, ctx):
        """Re-derive the bf16 GEMM operands from the fp32 masters (one pass over the flat buffer).  Called at the
        start of every forward so load_state_dict / in-place edits / optimizer steps can never leave them stale."""
        flat, _g, _p = self.flat_buffers()
        sig = None
        if getattr(self, "frozen_hint", False):
            # network that no optimizer of ours ever touches (VGG, vgg.py:27-28): torch bumps `_version` on every in-place
            # edit (load_state_dict, .copy_, init); identical versions => identical masters => packed copies still valid.
            # (NOT keyed on requires_grad: D is frozen in the G step but was just updated by the fused Adam kernel.)
            sig = tuple(p._version for p in self.parameters())
            if sig == self._sig:
                return
        self._sig = sig
        ctx.L.pack_bf16(flat.data_ptr(), self._flat_bf16.data_ptr(), flat.numel(), ctx.stream)
        self.pack_epoch += 1  # conv-weight slabs cached on the Params: first use packs lazily, later steps in one launch
        ctx.repack_slabs(self)


# ------------------------------------------------------------------------------------------------
# Generator graph
# ------------------------------------------------------------------------------------------------

def _block(ctx: Ctx, P, p, x: Var, out: Var = None, need_dx=True):
    """ConvNeXt Block (MixConvNeXtML.py:230-243): dw7x7 -> IN -> Linear(C,4C) -> GELU -> Linear(4C,P) (+) 1x1 shortcut."""
    W1, b1, W2, b2, Ws = (P[p + s] for s in (".pwconv1.weight", ".pwconv1.bias", ".pwconv2.weight", ".pwconv2.bias",
                                             ".shortcut.weight"))

    def branch():
        t = dwconv(ctx, x, P[p + ".dwconv.weight"], P[p + ".dwconv.bias"], 7, need_dx=need_dx, bias_via_in=True)
        return inorm(ctx, t)
    y_probe = out if out is not None else None
    plans = Ws.data.shape[0]
    if y_probe is None:
        y_probe = ctx.new(x.N, x.H, x.W, plans)
    if fused_mlp_ok(ctx, x, x, y_probe, (W1, W2, Ws)):
        # fused tcgen05 MLP: the 4C hidden never reaches HBM in the forward pass, is recomputed in the backward pass
        return block_mlp(ctx, x, branch, W1, b1, W2, b2, Ws, out=y_probe, need_dx=need_dx)
    # the shortcut writes y first and pwconv2 accumulates into it, so that in the backward pass (reverse order) the depthwise
    # input-gradient OVERWRITES dL/dx and the shortcut's GEMM epilogue does the fan-in add (cheaper than a depthwise RMW)
    y = conv2d(ctx, x, Ws, None, 1, out=y_probe, need_dx=need_dx)
    t = branch()
    h = conv2d(ctx, t, W1, b1, 1, act=ACT_GELU)
    conv2d(ctx, h, W2, b2, 1, out=y, acc=1)
    return y


def _upsample(ctx: Ctx, P, p, x: Var, skip: Var):
    """upSample (MixConvNeXtML.py:60-66): ConvT -> IN -> GELU written straight into the concat buffer."""
    t = conv_transpose2d(ctx, x, P[p + ".weight"], P[p + ".bias"], bias_via_in=True)
    par = skip.parent
    if par is not None and par.parent is None and skip.coff == t.C and par.C == t.C + skip.C:
        cat = par        # the skip tensor was produced in place in channels [C, 2C) of this concat buffer: nothing to copy
        inorm(ctx, t, act=ACT_GELU, out=cat.slice(0, t.C))
        return cat
    cat = ctx.new(t.N, t.H, t.W, t.C + skip.C)
    inorm(ctx, t, act=ACT_GELU, out=cat.slice(0, t.C))
    concat_into(ctx, cat, t.C, skip)
    return cat


def _downskip(ctx: Ctx, P, name, x: Var, k, pooled: Var = None):
    """MaxPool(k) -> 1x1 (no bias) -> IN -> GELU (MixConvNeXtML.py:328-426).  `pooled`: MaxPool(k)(x) if the caller already
    has it (the k = 2 branch pools the same tensor as the encoder's downSample: one kernel, gradients fan in)."""
    px = pooled if pooled is not None else maxpool(ctx, x, k)
    return inorm(ctx, conv2d(ctx, px, P[name + ".1.weight"], None, 1), act=ACT_GELU)


def _midmlka(ctx: Ctx, P, p, x: Var):
    """MidMLKA (MixConvNeXtML.py:109-117)."""
    cat = ctx.new(x.N, x.H, x.W, x.C)
    # The five biases of this module sit in front of an InstanceNorm (through the CA gate): their gradients are near-zero
    # sums that must not be taken over rounded bf16 tensors -> formed from fp32 plane statistics in ca_scale's backward.
    if os.environ.get("DSGAN_DW_MULTI", "1") != "0":
        dwconv_multi(ctx, x, [(P["%s.X%d.weight" % (p, k)], P["%s.X%d.bias" % (p, k)], k) for k in (3, 5, 7, 9)], out=cat,
                     bias_grad=False)
    else:   # one launch per branch (A/B measurements)
        q = x.C // 4
        for i, k in enumerate((3, 5, 7, 9)):
            dwconv(ctx, x.slice(i * q, q), P["%s.X%d.weight" % (p, k)], P["%s.X%d.bias" % (p, k)], k,
                   out=cat.slice(i * q, q), bias_grad=False)
    o = conv2d(ctx, cat, P[p + ".conv.weight"], P[p + ".conv.bias"], 1, bias_grad=False)
    holder = {}

    def sums():
        holder["S"] = ctx.zeros_f32(x.N, x.C)
        return holder["S"]
    o = ca_scale(ctx, o, P[p + ".attn.fc1.weight"], P[p + ".attn.relu1.weight"], P[p + ".attn.fc2.weight"],
                 mid_bias=(holder, P[p + ".conv.weight"], P[p + ".conv.bias"],
                           [P["%s.X%d.bias" % (p, k)] for k in (3, 5, 7, 9)]))
    return inorm(ctx, o, act=ACT_GELU, res=x, dsum_nc=sums)  # IN, then += x, then GELU (Q3)


def _local(ctx: Ctx, P, x: Var):
    """OriginMLKA (MixConvNeXtML.py:161-189)."""
    L = "local."
    d1 = conv2d(ctx, x, P[L + "to32.weight"], None, 1, need_dx=False)
    d2 = _midmlka(ctx, P, L + "mid32", maxpool(ctx, d1, 2))
    d3 = conv2d(ctx, d2, P[L + "to64.weight"], None, 1)
    d4 = _midmlka(ctx, P, L + "mid64", maxpool(ctx, d3, 2))
    d5 = conv2d(ctx, d4, P[L + "to128.weight"], None, 1)
    d6 = _midmlka(ctx, P, L + "mid128", maxpool(ctx, d5, 2))
    d7 = conv2d(ctx, d6, P[L + "to256.weight"], None, 1)
    d8 = _midmlka(ctx, P, L + "mid256", maxpool(ctx, d7, 2))
    u1 = _midmlka(ctx, P, L + "upc1.1", conv2d(ctx, _upsample(ctx, P, L + "up1.model.0", d8, d6),
                                               P[L + "upc1.0.weight"], None, 1))
    u2 = _midmlka(ctx, P, L + "upc2", _upsample(ctx, P, L + "up2.model.0", u1, d4))
    u3 = _midmlka(ctx, P, L + "upc3", _upsample(ctx, P, L + "up3.model.0", u2, d3))
    u4 = inorm(ctx, conv_transpose2d(ctx, u3, P[L + "up4.0.weight"], P[L + "up4.0.bias"], bias_via_in=True))
    sc = conv2d(ctx, x, P[L + "shortcut.0.weight"], None, 1, need_dx=False)
    return inorm(ctx, sc, act=ACT_GELU, res=u4)  # GELU(IN(up4) + IN(shortcut))


def generator_forward(ctx: Ctx, P, x: Var) -> Var:
    """MixConvNeXtML.forward (MixConvNeXtML.py:461-494) -> NHWC 3-channel output Var."""
    if x.H % 16 or x.W % 16:
        raise ValueError("MixConvNeXtML needs H and W to be multiples of 16, got %dx%d" % (x.H, x.W))
    fk = ctx.fork()          # the local (OriginMLKA) branch only shares the input with the main U-Net
    with fk:
        loc = _local(ctx, P, x)
    R, pools, t = [], [], x
    for i, (name, _cin, cout) in enumerate(specs.ENC):
        out = None
        if i < 4:   # R1..R4 are the decoder's skip tensors: write them straight into channels [C, 2C) of its concat buffer
            out = ctx.new(x.N, x.H >> i, x.W >> i, 2 * cout).slice(cout, cout)
        if i > 0:
            t = pools[-1][0]    # MaxPool2d(2)(R_i): the encoder's downSample, also the input of R_i's first down-skip branch
        t = _block(ctx, P, name, t, out=out, need_dx=i > 0)
        R.append(t)
        if i < 4:   # every pooled view of R_{i+1} the down-skips need (k = 2 .. 2^(4-i)) in ONE pass over it
            pools.append(multi_maxpool(ctx, t, 4 - i))
    R1, R2, R3, R4, R5 = R
    # pyramid[level] collects the down-skip tensors landing on that decoder level
    lvl = {16: [R5], 8: [], 4: [], 2: []}
    for (mod, _cin, branches), src, pl in zip(specs.SKIPS, (R1, R2, R3, R4), pools):
        for br, k, _cout in branches:
            scale = (x.H // src.H) * k          # total down-sampling w.r.t. the input
            lvl[scale].append(_downskip(ctx, P, "%s.%s" % (mod, br), src, k, pooled=pl[k.bit_length() - 2]))
    # (every addend below -- Block outputs and down-skip branches -- is consumed by its sum only: gradients are shared)
    o = add_n(ctx, lvl[16], share_grad=True)
    for (up, blk, _cin, _cout), skip, s in zip(specs.DEC, (R4, R3, R2, R1), (8, 4, 2, None)):
        o = _block(ctx, P, blk, _upsample(ctx, P, up + ".model.0", o, skip))
        if s is not None:
            o = add_n(ctx, [o] + lvl[s], share_grad=True)
    ctx.join(fk, keep=(x, loc))
    return conv2d(ctx, add_n(ctx, [o, loc], share_grad=True), P["res.weight"], P["res.bias"], 3, pad=1)


# ------------------------------------------------------------------------------------------------
# Discriminator / VGG graphs
# ------------------------------------------------------------------------------------------------

def discriminator_forward(ctx: Ctx, P, x: Var, need_dx=True) -> Var:
    """NLayerDiscriminator (networks.py:543-579): logits N x 30 x 30 x 1 (NHWC)."""
    h = conv2d(ctx, x, P["model.0.weight"], P["model.0.bias"], 4, 2, 1, act=ACT_LEAKY, need_dx=need_dx)
    for m, st in specs.D_LAYERS[1:4]:
        h = conv2d(ctx, h, P["model.%d.weight" % m], P["model.%d.bias" % m], 4, st, 1, bias_via_in=True)
        h = inorm(ctx, h, act=ACT_LEAKY)
    return conv2d(ctx, h, P["model.11.weight"], P["model.11.bias"], 4, 1, 1)


def vgg_forward(ctx: Ctx, P, x: Var, need_dx=True, with_tail=False):
    """Vgg16 taps relu1_2, relu2_2, relu3_3, relu4_3 (vgg.py:30-38); weights are frozen (vgg.py:27-28).  `with_tail`
    also runs the relu5_3 block (vgg.py:39-41): the training loss never reads it (pix2pix_model.py:182-186), the public
    Vgg16.forward returns it like the reference."""
    taps, h, first = [], x, True
    plan = specs.VGG_PLAN + (("P",) + specs.VGG_TAIL + ("T",) if with_tail else ())
    for e in plan:
        if e == "T":
            taps.append(h)
        elif e == "P":
            h = maxpool(ctx, h, 2)
        else:
            h = conv2d(ctx, h, P[e[0] + ".weight"], P[e[0] + ".bias"], 3, 1, 1, act=ACT_RELU,
                       need_dx=need_dx or not first)
            first = False
    return taps
