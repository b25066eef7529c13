"""CPU oracle for the DS-GAN adversarial training step.  TEST INFRASTRUCTURE ONLY.

This file is a functional, CPU-only restatement (plain ``torch.nn.functional`` in
fp32/fp64) of the reference's hot path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``dsgan_b200`` never does.

Parity pin: the reference ships no tests/golden vectors, so this oracle is
pinned against the *reference itself* executed in the build container:
``oracle/make_golden.py`` imports ``/root/reference/DSGAN`` (with import stubs only), loads the
weights produced by :func:`init_params_G` / ``_D`` / ``_vgg`` into the reference modules,
runs the reference and stores the fingerprints under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this file against those fixtures.

Each function cites the reference file:line (under /root/reference/DSGAN) it restates.
The arithmetic below PyTorch's operator boundary (conv, instance_norm, gelu, max_pool,
adam ...) is third-party: torch 2.11.0 / torchvision 0.26.0 as installed (the reference pins only
``torch>=0.4.0``, requirements.txt:1-2).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Parameter inventories (names/shapes == reference state_dict, SURVEY.md §5 "Checkpoint")
# --------------------------------------------------------------------------------------


def _convnext_block(p, dim, plans):
    # models/model/MixConvNeXtML.py:213-224 (registration order: shortcut, dwconv, pwconv1, pwconv2)
    return [
        (p + ".shortcut.weight", (plans, dim, 1, 1)),
        (p + ".dwconv.weight", (dim, 1, 7, 7)),
        (p + ".dwconv.bias", (dim,)),
        (p + ".pwconv1.weight", (4 * dim, dim)),
        (p + ".pwconv1.bias", (4 * dim,)),
        (p + ".pwconv2.weight", (plans, 4 * dim)),
        (p + ".pwconv2.bias", (plans,)),
    ]


def _up(p, cin, cout, idx="model.0"):
    # MixConvNeXtML.py:53 ConvTranspose2d weight is (C_in, C_out, 3, 3)
    return [(f"{p}.{idx}.weight", (cin, cout, 3, 3)), (f"{p}.{idx}.bias", (cout,))]


def _mlka(p, dim):
    # MixConvNeXtML.py:76-97 (registration order: conv, attn.fc1, attn.relu1, attn.fc2, X3, X5, X7, X9)
    q = dim // 4
    out = [
        (p + ".conv.weight", (dim, dim, 1, 1)),
        (p + ".conv.bias", (dim,)),
        (p + ".attn.fc1.weight", (dim // 8, dim, 1, 1)),
        (p + ".attn.relu1.weight", (1,)),
        (p + ".attn.fc2.weight", (dim, dim // 8, 1, 1)),
    ]
    for k in (3, 5, 7, 9):
        out += [(f"{p}.X{k}.weight", (q, 1, k, k)), (f"{p}.X{k}.bias", (q,))]
    return out


def g_param_spec():
    """MixConvNeXtML parameters in state_dict order (MixConvNeXtML.py:428-459)."""
    s = []
    s += _convnext_block("c1", 3, 64) + _convnext_block("c2", 64, 128) + _convnext_block("c3", 128, 256)
    s += _convnext_block("c4", 256, 512) + _convnext_block("c5", 512, 1024)
    s += _up("u1", 1024, 512) + _convnext_block("uc1", 1024, 512)
    s += _up("u2", 512, 256) + _convnext_block("uc2", 512, 256)
    s += _up("u3", 256, 128) + _convnext_block("uc3", 256, 128)
    s += _up("u4", 128, 64) + _convnext_block("uc4", 128, 64)
    for cin, name, outs in (
        (64, "down64", (("to2", 128), ("to4", 256), ("to8", 512), ("to16", 1024))),
        (128, "down128", (("to4", 256), ("to8", 512), ("to16", 1024))),
        (256, "down256", (("to8", 512), ("to16", 1024))),
        (512, "down512", (("to16", 1024),)),
    ):
        for br, cout in outs:  # MixConvNeXtML.py:328-426, conv is index 1 of each Sequential
            s.append((f"{name}.{br}.1.weight", (cout, cin, 1, 1)))
    # OriginMLKA, MixConvNeXtML.py:119-159
    s += [("local.to32.weight", (32, 3, 1, 1))] + _mlka("local.mid32", 32)
    s += [("local.to64.weight", (64, 32, 1, 1))] + _mlka("local.mid64", 64)
    s += [("local.to128.weight", (128, 64, 1, 1))] + _mlka("local.mid128", 128)
    s += [("local.to256.weight", (256, 128, 1, 1))] + _mlka("local.mid256", 256)
    s += _up("local.up1", 256, 128)
    s += [("local.upc1.0.weight", (128, 256, 1, 1))] + _mlka("local.upc1.1", 128)
    s += _up("local.up2", 128, 64) + _mlka("local.upc2", 128)
    s += _up("local.up3", 128, 64) + _mlka("local.upc3", 128)
    s += _up("local.up4", 128, 64, idx="0")
    s += [("local.shortcut.0.weight", (64, 3, 1, 1))]
    s += [("res.weight", (3, 64, 3, 3)), ("res.bias", (3,))]
    return s


def d_param_spec(input_nc=6, ndf=32):
    """NLayerDiscriminator(n_layers=3) parameters (models/networks.py:533-579)."""
    chans = [input_nc, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    idx = [0, 2, 5, 8, 11]
    s = []
    for i, m in enumerate(idx):
        s += [(f"model.{m}.weight", (chans[i + 1], chans[i], 4, 4)), (f"model.{m}.bias", (chans[i + 1],))]
    return s


VGG_CFG = [  # (state_dict prefix in models/vgg.py:16-25, cin, cout); "P" = MaxPool2d(2)
    ("to_relu_1_2.0", 3, 64), ("to_relu_1_2.2", 64, 64), "TAP", "P",
    ("to_relu_2_2.5", 64, 128), ("to_relu_2_2.7", 128, 128), "TAP", "P",
    ("to_relu_3_3.10", 128, 256), ("to_relu_3_3.12", 256, 256), ("to_relu_3_3.14", 256, 256), "TAP", "P",
    ("to_relu_4_3.17", 256, 512), ("to_relu_4_3.19", 512, 512), ("to_relu_4_3.21", 512, 512), "TAP",
]


def vgg_param_spec():
    """VGG16 features up to relu4_3.  The relu5_3 block (vgg.py:39-40) is computed by the reference
    but never read by the loss (pix2pix_model.py:182-186), so it is not part of the oracle."""
    s = []
    for e in VGG_CFG:
        if isinstance(e, tuple):
            s += [(e[0] + ".weight", (e[2], e[1], 3, 3)), (e[0] + ".bias", (e[2],))]
    return s


def _gen_for(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def _init(spec, seed, bias_std, dtype, vgg=False):
    out = OrderedDict()
    for name, shape in spec:
        g = _gen_for(name, seed)
        if name.endswith("relu1.weight"):
            t = torch.full(shape, 0.25)  # nn.PReLU default, untouched by networks.py:49-70
        elif name.endswith(".bias"):
            t = torch.randn(shape, generator=g) * bias_std  # reference: 0 (networks.py:63-64)
        elif vgg:
            # torchvision vgg16(weights=None): kaiming_normal_(fan_out, relu)
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
        else:
            t = torch.randn(shape, generator=g) * 0.02  # networks.py:53-54, init_type 'normal'
        out[name] = t.to(dtype)
    return out


def init_params_G(seed=20, bias_std=0.0, dtype=torch.float32):
    """N(0,0.02) weights, zero (or N(0,bias_std)) biases, PReLU slope 0.25 — networks.py:49-79.
    Per-tensor seeded by name so any subset can be regenerated on any box."""
    return _init(g_param_spec(), seed, bias_std, dtype)


def init_params_D(seed=20, bias_std=0.0, dtype=torch.float32):
    return _init(d_param_spec(), seed + 1, bias_std, dtype)


def init_params_vgg(seed=20, bias_std=0.0, dtype=torch.float32):
    return _init(vgg_param_spec(), seed + 2, bias_std, dtype, vgg=True)


def synthetic_pair(n, h=256, w=256, seed=1, dtype=torch.float32):
    """Synthetic TIR/RGB pair, SURVEY.md §8(d): A grey in [-1,1], B correlated with A."""
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(n, 1, h, w, generator=g) * 2 - 1).expand(n, 3, h, w).contiguous()
    b = torch.clamp(0.5 * a + 0.5 * (torch.rand(n, 3, h, w, generator=g) * 2 - 1), -1, 1)
    return a.to(dtype), b.to(dtype)


# --------------------------------------------------------------------------------------
# Generator (models/model/MixConvNeXtML.py)
# --------------------------------------------------------------------------------------


def _inorm(x):
    # nn.InstanceNorm2d(affine=False, track_running_stats=False, eps=1e-5): networks.py:25
    return F.instance_norm(x, eps=1e-5)


def _block(P, p, x, taps=None):
    """ConvNeXt Block.forward, MixConvNeXtML.py:230-243."""
    h = F.conv2d(x, P[p + ".dwconv.weight"], P[p + ".dwconv.bias"], padding=3, groups=x.shape[1])
    h = _inorm(h).permute(0, 2, 3, 1)
    h = F.linear(h, P[p + ".pwconv1.weight"], P[p + ".pwconv1.bias"])
    h = F.gelu(h)  # exact erf GELU (:223)
    h = F.linear(h, P[p + ".pwconv2.weight"], P[p + ".pwconv2.bias"]).permute(0, 3, 1, 2)
    y = F.conv2d(x, P[p + ".shortcut.weight"]) + h
    if taps is not None:
        taps[p] = y
    return y


def _upsample(P, p, x, skip, idx="model.0"):
    """upSample.forward, MixConvNeXtML.py:60-66: ConvT(k3,s2,p1,op1) -> IN -> GELU -> cat(skip)."""
    y = F.conv_transpose2d(x, P[f"{p}.{idx}.weight"], P[f"{p}.{idx}.bias"], stride=2, padding=1, output_padding=1)
    y = F.gelu(_inorm(y))
    return torch.cat((y, skip), 1)


def _downskip(P, name, x, k):
    """MaxPool2d(k) -> 1x1 conv (no bias) -> IN -> GELU, MixConvNeXtML.py:328-426."""
    return F.gelu(_inorm(F.conv2d(F.max_pool2d(x, k), P[name + ".1.weight"])))


def _ca(P, p, x):
    """CA.forward, MixConvNeXtML.py:17-22 (shared fc1/PReLU/fc2 on avg- and max-pooled vectors)."""
    def mlp(v):
        return F.conv2d(F.prelu(F.conv2d(v, P[p + ".fc1.weight"]), P[p + ".relu1.weight"]), P[p + ".fc2.weight"])
    return torch.sigmoid(mlp(F.adaptive_avg_pool2d(x, 1)) + mlp(F.adaptive_max_pool2d(x, 1)))


def _midmlka(P, p, x):
    """MidMLKA.forward, MixConvNeXtML.py:109-117."""
    q = x.shape[1] // 4
    parts = []
    for i, k in enumerate((3, 5, 7, 9)):
        xi = x[:, i * q:(i + 1) * q]
        parts.append(F.conv2d(xi, P[f"{p}.X{k}.weight"], P[f"{p}.X{k}.bias"], padding=k // 2, groups=q))
    out = F.conv2d(torch.cat(parts, 1), P[p + ".conv.weight"], P[p + ".conv.bias"])
    out = out * _ca(P, p + ".attn", out)
    out = _inorm(out) + x  # residual is added AFTER the norm (:113-114)
    return F.gelu(out)


def _local(P, x, taps=None):
    """OriginMLKA.forward, MixConvNeXtML.py:161-189."""
    L = "local."
    d1 = F.conv2d(x, P[L + "to32.weight"])
    d2 = _midmlka(P, L + "mid32", F.max_pool2d(d1, 2))
    d3 = F.conv2d(d2, P[L + "to64.weight"])
    d4 = _midmlka(P, L + "mid64", F.max_pool2d(d3, 2))
    d5 = F.conv2d(d4, P[L + "to128.weight"])
    d6 = _midmlka(P, L + "mid128", F.max_pool2d(d5, 2))
    d7 = F.conv2d(d6, P[L + "to256.weight"])
    d8 = _midmlka(P, L + "mid256", F.max_pool2d(d7, 2))
    u1 = _midmlka(P, L + "upc1.1", F.conv2d(_upsample(P, L + "up1", d8, d6), P[L + "upc1.0.weight"]))
    u2 = _midmlka(P, L + "upc2", _upsample(P, L + "up2", u1, d4))
    u3 = _midmlka(P, L + "upc3", _upsample(P, L + "up3", u2, d3))
    u4 = _inorm(F.conv_transpose2d(u3, P[L + "up4.0.weight"], P[L + "up4.0.bias"], stride=2, padding=1,
                                   output_padding=1))
    out = F.gelu(u4 + _inorm(F.conv2d(x, P[L + "shortcut.0.weight"])))
    if taps is not None:
        taps.update({"local.d2": d2, "local.d8": d8, "local.u1": u1, "local.u3": u3, "local": out})
    return out


def g_forward(P, x, taps=None):
    """MixConvNeXtML.forward, MixConvNeXtML.py:461-494.  x: N×3×H×W, H,W % 16 == 0 (Q19)."""
    if x.shape[2] % 16 or x.shape[3] % 16:
        raise ValueError("MixConvNeXtML needs H and W to be multiples of 16, got %s" % (tuple(x.shape),))
    R1 = _block(P, "c1", x, taps)
    R2 = _block(P, "c2", F.max_pool2d(R1, 2), taps)
    R3 = _block(P, "c3", F.max_pool2d(R2, 2), taps)
    R4 = _block(P, "c4", F.max_pool2d(R3, 2), taps)
    R5 = _block(P, "c5", F.max_pool2d(R4, 2), taps)
    d64 = [_downskip(P, "down64." + n, R1, k) for n, k in (("to2", 2), ("to4", 4), ("to8", 8), ("to16", 16))]
    d128 = [_downskip(P, "down128." + n, R2, k) for n, k in (("to4", 2), ("to8", 4), ("to16", 8))]
    d256 = [_downskip(P, "down256." + n, R3, k) for n, k in (("to8", 2), ("to16", 4))]
    d512 = [_downskip(P, "down512.to16", R4, 2)]
    O1 = _block(P, "uc1", _upsample(P, "u1", R5 + d64[3] + d128[2] + d256[1] + d512[0], R4), taps)
    O2 = _block(P, "uc2", _upsample(P, "u2", O1 + d64[2] + d128[1] + d256[0], R3), taps)
    O3 = _block(P, "uc3", _upsample(P, "u3", O2 + d64[1] + d128[0], R2), taps)
    O4 = _block(P, "uc4", _upsample(P, "u4", O3 + d64[0], R1), taps)
    loc = _local(P, x, taps)
    return F.conv2d(O4 + loc, P["res.weight"], P["res.bias"], padding=1)  # no tanh (:492)


# --------------------------------------------------------------------------------------
# Discriminator, VGG, losses
# --------------------------------------------------------------------------------------


def d_forward(P, x, use_sigmoid=False, taps=None):
    """NLayerDiscriminator(n_layers=3, InstanceNorm) forward, networks.py:543-579."""
    h = F.leaky_relu(F.conv2d(x, P["model.0.weight"], P["model.0.bias"], stride=2, padding=1), 0.2)
    for m, s in ((2, 2), (5, 2), (8, 1)):
        h = F.conv2d(h, P[f"model.{m}.weight"], P[f"model.{m}.bias"], stride=s, padding=1)
        h = F.leaky_relu(_inorm(h), 0.2)
        if taps is not None:
            taps[f"d{m}"] = h
    h = F.conv2d(h, P["model.11.weight"], P["model.11.bias"], stride=1, padding=1)
    return torch.sigmoid(h) if use_sigmoid else h


def vgg_forward(P, x):
    """Vgg16.forward taps relu1_2, relu2_2, relu3_3, relu4_3 (models/vgg.py:30-38)."""
    taps, h = [], x
    for e in VGG_CFG:
        if e == "TAP":
            taps.append(h)
        elif e == "P":
            h = F.max_pool2d(h, 2)
        else:
            h = F.relu(F.conv2d(h, P[e[0] + ".weight"], P[e[0] + ".bias"], padding=1))
    return taps


def gan_loss(pred, target_is_real, use_lsgan=False):
    """GANLoss.__call__, networks.py:154-163 (default: BCEWithLogits vs constant 1.0/0.0)."""
    t = torch.full_like(pred, 1.0 if target_is_real else 0.0)
    return F.mse_loss(pred, t) if use_lsgan else F.binary_cross_entropy_with_logits(pred, t)


def tv_loss(x):
    """pix2pix_model.py:189-191: batch SUM of |dx|+|dy| over the literal 320*256."""
    di = (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().sum()
    dj = (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().sum()
    return (di + dj) / (320 * 256)


def gauss_window(size=11, sigma=1.5, dtype=torch.float32):
    """_fspecial_gauss_1d, MS_SSIM.py:9-23."""
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).to(dtype)


def _blur(x, win):
    """gaussian_filter, MS_SSIM.py:26-52: valid separable conv, along H then along W."""
    c = x.shape[1]
    k = win.to(x.dtype).view(1, 1, -1).repeat(c, 1, 1)
    x = F.conv2d(x, k.unsqueeze(-1), groups=c)   # (C,1,11,1): along H
    return F.conv2d(x, k.unsqueeze(-2), groups=c)  # (C,1,1,11): along W


def ssim_maps(X, Y, data_range=1.0, win=None, K=(0.01, 0.03)):
    """_ssim, MS_SSIM.py:55-92 -> per-(N,C) means of ssim_map and cs_map."""
    win = gauss_window() if win is None else win
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _blur(X, win), _blur(Y, win)
    s1 = _blur(X * X, win) - mu1 * mu1
    s2 = _blur(Y * Y, win) - mu2 * mu2
    s12 = _blur(X * Y, win) - mu1 * mu2
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    sm = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs
    return sm.flatten(2).mean(-1), cs.flatten(2).mean(-1)


def ssim(X, Y, data_range=1.0, size_average=True):
    """ssim, MS_SSIM.py:95-150 (nonnegative_ssim=False)."""
    s, _ = ssim_maps(X, Y, data_range)
    return s.mean() if size_average else s.mean(1)


MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # MS_SSIM.py:200


def ms_ssim(X, Y, data_range=1.0, size_average=True):
    """ms_ssim, MS_SSIM.py:153-225."""
    if min(X.shape[-2:]) <= 160:
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    vals = []
    for lvl in range(5):
        s, cs = ssim_maps(X, Y, data_range)
        if lvl < 4:
            vals.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X, Y = F.avg_pool2d(X, 2, padding=pad), F.avg_pool2d(Y, 2, padding=pad)
    vals.append(torch.relu(s))
    w = torch.tensor(MS_WEIGHTS, dtype=X.dtype).view(-1, 1, 1)
    v = torch.prod(torch.stack(vals, 0) ** w, 0)
    return v.mean() if size_average else v.mean(1)


# --------------------------------------------------------------------------------------
# The training step (models/pix2pix_model.py:129-217)
# --------------------------------------------------------------------------------------

DEFAULT_W = dict(w_gan=0.01, w_vgg=1.0, w_tv=1.0, w_ss=1.25)  # options/base_options.py:64-68


def g_losses(PG, PD, PV, A, B, fake, w=DEFAULT_W, tv_scale=1.0):
    """backward_G's loss, pix2pix_model.py:164-197 (lambda_L1 is parsed but unused, Q11)."""
    out = OrderedDict()
    out["G_GAN"] = gan_loss(d_forward(PD, torch.cat((A, fake), 1)), True)
    out["G_L1"] = F.l1_loss(fake, B)
    fr, ff = vgg_forward(PV, B), vgg_forward(PV, fake)
    out["vgg"] = sum(F.l1_loss(f, r) for f, r in zip(ff, fr))
    out["tv"] = tv_loss(fake)
    out["ssim"] = 1 - ssim((B + 1) / 2, (fake + 1) / 2, 1.0)
    out["G"] = (out["G_GAN"] * w["w_gan"] + out["G_L1"] + out["vgg"] * w["w_vgg"]
                + out["tv"] * w["w_tv"] * tv_scale + out["ssim"] * w["w_ss"])
    return out


def adam_step(params, grads, state, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
    """torch.optim.Adam semantics (pix2pix_model.py:122-125): no weight decay, eps after the
    bias-corrected sqrt."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k, p in params.items():
        g = grads[k]
        m = state.setdefault("m." + k, torch.zeros_like(p))
        v = state.setdefault("v." + k, torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = (v.sqrt() / math.sqrt(1 - betas[1] ** t)).add_(eps)
        p.data.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** t))


def train_step(PG, PD, PV, A, B, opt_state=None, lr=2e-4, update=True, tv_scale=1.0, pooled_fake_AB=None):
    """One optimize_parameters() (pix2pix_model.py:201-217).  Returns losses (floats), fake_B and the
    D/G gradients of this step.  ImagePool is the identity for the first 50 images (image_pool.py:18-21);
    pass ``pooled_fake_AB`` to emulate a later swap."""
    opt_state = {} if opt_state is None else opt_state
    PGr = {k: v.detach().requires_grad_(True) for k, v in PG.items()}
    PDr = {k: v.detach().requires_grad_(True) for k, v in PD.items()}
    fake = g_forward(PGr, A)
    # ---- D step (backward_D :141-162)
    fake_AB = torch.cat((A, fake.detach()), 1) if pooled_fake_AB is None else pooled_fake_AB
    l_fake = gan_loss(d_forward(PDr, fake_AB), False)
    l_real = gan_loss(d_forward(PDr, torch.cat((A, B), 1)), True)
    l_d = 0.5 * (l_fake + l_real)
    gD = dict(zip(PDr.keys(), torch.autograd.grad(l_d, list(PDr.values()))))
    PD_new = {k: v.detach().clone() for k, v in PD.items()}
    if update:
        adam_step(PD_new, gD, opt_state.setdefault("D", {}), lr)
    # ---- G step (backward_G :164-199) uses the UPDATED D (Q13)
    PDf = {k: v.detach() for k, v in PD_new.items()}
    L = g_losses(PGr, PDf, PV, A, B, fake, tv_scale=tv_scale)
    gfake, *gG = torch.autograd.grad(L["G"], [fake] + list(PGr.values()))
    gG = dict(zip(PGr.keys(), gG))
    PG_new = {k: v.detach().clone() for k, v in PG.items()}
    if update:
        adam_step(PG_new, gG, opt_state.setdefault("G", {}), lr)
    losses = OrderedDict((k, float(v)) for k, v in L.items())
    losses["D_fake"], losses["D_real"], losses["D"] = float(l_fake), float(l_real), float(l_d)
    return dict(losses=losses, fake_B=fake.detach(), grad_fake_B=gfake, grads_D=gD, grads_G=gG,
                PD=PD_new, PG=PG_new, opt_state=opt_state)


def fingerprint(t):
    """(l2 norm, sum, fixed pseudo-random projection) — compact golden record for a big tensor."""
    t = t.detach().double().flatten()
    g = torch.Generator().manual_seed(12345 + t.numel() % 9973)
    r = torch.randn(min(t.numel(), 4096), generator=g, dtype=torch.float64)
    idx = torch.linspace(0, t.numel() - 1, r.numel()).long()
    return [float(t.norm()), float(t.sum()), float((t[idx] * r).sum())]
