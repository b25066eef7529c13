"""Network factory with the reference's signatures (DSGAN/models/networks.py:21-163): get_norm_layer, get_scheduler,
define_G, define_D, GANLoss.  The returned networks are nn.Module shells (ParamTree) whose forward runs the sm_100a
kernel graph; there is no DataParallel wrapper — data parallelism is one process per GPU (dsgan_b200.parallel)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import nets, specs
from ..engine import Ctx, Var, image_to_nhwc

_CTX = {}


def get_ctx(device, precision="bf16") -> Ctx:
    key = (str(torch.device(device)), precision)
    if key not in _CTX:
        _CTX[key] = Ctx(device, precision)
    return _CTX[key]


def get_norm_layer(norm_type="instance"):
    if norm_type in ("instance", "none"):
        return norm_type
    if norm_type == "batch":
        raise NotImplementedError("normalization layer [batch] is off the default path and not built (SURVEY §0.2)")
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


class LambdaRule:
    """lr_policy 'lambda' (networks.py:33-39): lr * (1 - max(0, epoch + 1 + epoch_count - niter)/(niter_decay+1)),
    stepped once per epoch by BaseModel.update_learning_rate."""

    def __init__(self, optimizer, opt):
        self.optimizer, self.opt, self.epoch = optimizer, opt, 0
        self.base_lr = optimizer.param_groups[0]["lr"]
        self._apply()

    def _apply(self):
        o = self.opt
        f = 1.0 - max(0, self.epoch + 1 + o.epoch_count - o.niter) / float(o.niter_decay + 1)
        self.optimizer.param_groups[0]["lr"] = self.base_lr * f

    def step(self):
        self.epoch += 1
        self._apply()


def get_scheduler(optimizer, opt):
    if opt.lr_policy == "lambda":
        return LambdaRule(optimizer, opt)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


class KernelNet(nets.ParamTree):
    """Base of G/D/VGG shells: parameters under reference names + an engine context chosen at first use."""
    precision = "bf16"

    def ctx(self) -> Ctx:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dsgan_b200 networks run on a B200 only (no CPU fallback); move the module to cuda")
        return get_ctx(dev, self.precision)

    def params(self):
        """{name: Param}; in bf16 mode the packed bf16 operands are refreshed first."""
        P = self.flat_buffers()[2]
        ctx = self.ctx()
        if ctx.precision == "bf16":
            self.refresh_bf16(ctx)
        return P


class MixConvNeXtML(KernelNet):
    """Generator shell (reference: models/model/MixConvNeXtML.py:428-494)."""

    def __init__(self):
        super().__init__(specs.generator_spec())

    def forward_var(self, real_A: torch.Tensor) -> Var:
        ctx = self.ctx()
        return nets.generator_forward(ctx, self.params(), image_to_nhwc(ctx, real_A.contiguous().float()))

    def forward(self, real_A):
        """N x 3 x H x W fp32 (NCHW) -> N x 3 x H x W fp32.  The autograd tape is left on the context."""
        ctx = self.ctx()
        y = self.forward_var(real_A)
        out = torch.empty((y.N, y.C, y.H, y.W), dtype=torch.float32, device=real_A.device)
        ctx.L.nhwc_to_nchw(y.ptr, ctx.dt, y.ld, out.data_ptr(), y.N, y.C, y.H, y.W, 1.0, 0, ctx.stream)
        self.last_output = y
        return out


class NLayerDiscriminator(KernelNet):
    """PatchGAN shell (reference: networks.py:533-579) for n_layers=3, InstanceNorm."""

    def __init__(self, input_nc, ndf=32, n_layers=3, use_sigmoid=False):
        if n_layers != 3:
            raise NotImplementedError("only the reference default n_layers=3 ('basic') discriminator is built")
        super().__init__(specs.discriminator_spec(input_nc, ndf))
        self.use_sigmoid = use_sigmoid

    def forward_var(self, x: Var, need_dx=True) -> Var:
        return nets.discriminator_forward(self.ctx(), self.params(), x, need_dx)

    def forward(self, x):
        """N x C x H x W fp32 -> N x 1 x 30 x 30 logits (fp32)."""
        ctx = self.ctx()
        y = self.forward_var(image_to_nhwc(ctx, x.contiguous().float()), need_dx=False)
        out = torch.empty((y.N, 1, y.H, y.W), dtype=torch.float32, device=x.device)
        ctx.L.nhwc_to_nchw(y.ptr, ctx.dt, y.ld, out.data_ptr(), y.N, 1, y.H, y.W, 1.0, 0, ctx.stream)
        return torch.sigmoid(out) if self.use_sigmoid else out


def init_net(net, init_type="normal", gpu_ids=()):
    """networks.py:73-79 without the DataParallel wrapper."""
    if init_type != "normal":
        raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
    print("initialize network with %s" % init_type)
    net.init_normal(0.02)
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(torch.device("cuda:%d" % gpu_ids[0]))
    return net


def define_G(input_nc, output_nc, ngf, which_model_netG, norm="batch", use_dropout=False, init_type="normal",
             gpu_ids=()):
    """networks.py:81-113.  As in the reference, MixConvNeXtML ignores input_nc/output_nc/ngf/norm/use_dropout
    (it is constructed with no arguments, :108-109)."""
    get_norm_layer(norm)
    if which_model_netG != "MixConvNeXtML":
        raise NotImplementedError("Generator model name [%s] is not recognized" % which_model_netG)
    return init_net(MixConvNeXtML(), init_type, gpu_ids)


def define_D(input_nc, ndf, which_model_netD, n_layers_D=3, norm="batch", use_sigmoid=False, init_type="normal",
             gpu_ids=()):
    """networks.py:115-131 ('basic' and 'n_layers' with 3 layers)."""
    if get_norm_layer(norm) != "instance":
        raise NotImplementedError("the built discriminator uses InstanceNorm (reference default)")
    if which_model_netD == "basic" or (which_model_netD == "n_layers" and n_layers_D == 3):
        return init_net(NLayerDiscriminator(input_nc, ndf, 3, use_sigmoid), init_type, gpu_ids)
    raise NotImplementedError("Discriminator model name [%s] is not recognized" % which_model_netD)


class GANLoss(nn.Module):
    """networks.py:143-163: use_lsgan=True -> MSELoss, else BCEWithLogitsLoss, against a constant label."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.use_lsgan = use_lsgan

    def __call__(self, input, target_is_real):
        """input: CUDA tensor of predictions -> 0-d fp32 loss tensor (value only; the training step uses the fused
        loss+gradient kernel directly)."""
        ctx = get_ctx(input.device, "fp32")
        x = input.contiguous().float()
        val = torch.zeros(1, dtype=torch.float32, device=input.device)
        t = float(self.real_label if target_is_real else self.fake_label)
        ctx.L.gan_loss(x.data_ptr(), 0, x.numel(), 1, t, 1 if self.use_lsgan else 0, 1.0, val.data_ptr(), 0.0, None,
                       ctx.stream)
        return val[0]
