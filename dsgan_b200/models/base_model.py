"""BaseModel with the reference's surface (DSGAN/models/base_model.py:7-177): initialize, setup, eval, test,
update_learning_rate, get_current_visuals/losses, save/load_networks, set_requires_grad."""
import os
from collections import OrderedDict

import torch

from . import networks


class BaseModel:
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def name(self):
        return "BaseModel"

    def initialize(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        if not self.gpu_ids:
            raise RuntimeError("dsgan_b200 runs on B200 GPUs only: gpu_ids=-1 (CPU) is the reference's path, "
                               "not this one (no CPU fallback)")
        self.device = torch.device("cuda:{}".format(self.gpu_ids[0]))
        if len(self.gpu_ids) > 1:
            import warnings
            warnings.warn("gpu_ids=%s: only %s is used by this process; multi-GPU training is one process per GPU "
                          "(torchrun), not nn.DataParallel (networks.py:77)" % (self.gpu_ids, self.device))
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.visual_names, self.image_paths = [], [], [], []

    def set_input(self, input):
        self.input = input

    def forward(self):
        pass

    def setup(self, opt, parser=None):
        if self.isTrain:
            self.schedulers = [networks.get_scheduler(optimizer, opt) for optimizer in self.optimizers]
        if not self.isTrain or opt.continue_train:
            self.load_networks(opt.which_epoch)
        self.print_networks(opt.verbose)

    def eval(self):
        pass  # no BatchNorm/Dropout on the built path: train and eval are numerically identical

    def test(self):
        ctx = self.netG.ctx()
        was, ctx.no_grad = ctx.no_grad, True
        try:
            self.forward()
        finally:
            ctx.no_grad = was
            ctx.clear()

    def get_image_paths(self):
        return self.image_paths

    def optimize_parameters(self):
        pass

    def update_learning_rate(self):
        for scheduler in self.schedulers:
            scheduler.step()
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str))

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n))) for n in self.loss_names if isinstance(n, str))

    def save_networks(self, which_epoch):
        """<epoch>_useSE_net_<name>.pth holding net.state_dict() with DataParallel's `module.` prefix, exactly the
        file the reference writes on GPU (base_model.py:92-103, Q16/Q17).  Under one-process-per-GPU data parallelism
        only rank 0 writes (tmp file + atomic rename) and every rank waits for it."""
        from .. import parallel
        if parallel.rank() == 0:
            os.makedirs(self.save_dir, exist_ok=True)
            for name in self.model_names:
                net = getattr(self, "net" + name)
                sd = OrderedDict(("module." + k, v.detach().cpu()) for k, v in net.state_dict().items())
                path = os.path.join(self.save_dir, "%s_useSE_net_%s.pth" % (which_epoch, name))
                torch.save(sd, path + ".tmp")
                os.replace(path + ".tmp", path)
        parallel.barrier()

    def load_networks(self, which_epoch):
        """Reads <epoch>_net_<name>.pth like the reference (base_model.py:116-148) and, because the reference cannot
        re-read its own `_useSE_` files (Q17), falls back to that name.  Accepts bare or `module.`-prefixed keys."""
        for name in self.model_names:
            path = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, name))
            if not os.path.exists(path):
                path = os.path.join(self.save_dir, "%s_useSE_net_%s.pth" % (which_epoch, name))
            print("loading the model from %s" % path)
            sd = torch.load(path, map_location="cpu")
            sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
            getattr(self, "net" + name).load_state_dict(sd, strict=False)

    def print_networks(self, verbose):
        print("---------- Networks initialized -------------")
        for name in self.model_names:
            net = getattr(self, "net" + name)
            n = sum(p.numel() for p in net.parameters())
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (name, n / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):
        for net in nets if isinstance(nets, list) else [nets]:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad
