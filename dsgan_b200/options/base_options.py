"""Option system with the reference's flags, defaults and two-pass model hook (DSGAN/options/base_options.py:12-141).
Table-driven; `parse(dataset_path, path, argv=None)` keeps the reference's non-standard signature and adds an
optional argv (the reference always reads sys.argv)."""
import argparse
import os

import torch

INF = float("inf")
# (flag, kwargs) — defaults are the reference's (base_options.py:16-69).  The six loss/GAN switches have no
# `type=` in the reference, so values given on the command line arrive as strings there; only defaults behave
# (SURVEY §5 "Config / flags").  The same (untyped) definition is kept.
BASE_FLAGS = [
    ("--dataroot", dict(type=str, default="/root/dataset/256x256")),
    ("--batchSize", dict(type=int, default=1)),
    ("--loadSize_w", dict(type=int, default=256)), ("--fineSize_w", dict(type=int, default=256)),
    ("--loadSize_h", dict(type=int, default=256)), ("--fineSize_h", dict(type=int, default=256)),
    ("--input_nc", dict(type=int, default=3)), ("--output_nc", dict(type=int, default=3)),
    ("--ngf", dict(type=int, default=32)), ("--ndf", dict(type=int, default=32)),
    ("--which_model_netD", dict(type=str, default="basic")),
    ("--which_model_netG", dict(type=str, default="MixConvNeXtML")),
    ("--n_layers_D", dict(type=int, default=3)),
    ("--gpu_ids", dict(type=str, default="0")),
    ("--name", dict(type=str, default="experiment_name")),
    ("--dataset_mode", dict(type=str, default="aligned")),
    ("--model", dict(type=str, default="pix2pix")),
    ("--which_direction", dict(type=str, default="AtoB")),
    ("--nThreads", dict(type=int, default=4)),
    ("--checkpoints_dir", dict(type=str, default="./checkpoints/")),
    ("--norm", dict(type=str, default="instance")),
    ("--serial_batches", dict(action="store_true")),
    ("--display_winsize", dict(type=int, default=256)), ("--display_id", dict(type=int, default=1)),
    ("--display_server", dict(type=str, default="http://localhost")), ("--display_port", dict(type=int, default=8097)),
    ("--no_dropout", dict(action="store_true")),
    ("--max_dataset_size", dict(type=int, default=INF)),
    ("--resize_or_crop", dict(type=str, default="resize_and_crop")),
    ("--no_flip", dict(action="store_true")),
    ("--init_type", dict(type=str, default="normal")),
    ("--verbose", dict(action="store_true")),
    ("--suffix", dict(type=str, default="")),
    ("--use_GAN", dict(default=1)), ("--w_gan", dict(default=0.01)), ("--w_vgg", dict(default=1)),
    ("--w_tv", dict(default=1)), ("--w_ss", dict(default=1.25)), ("--use_condition", dict(default=1)),
    # extension (not in the reference): activation precision of the sm_100a kernels
    ("--precision", dict(type=str, default="bf16", choices=["bf16", "fp32"])),
    # extension: per-network override of --precision (the networks only meet at NCHW fp32 images)
    ("--precision_G", dict(type=str, default="", choices=["", "bf16", "fp32"])),
    ("--precision_D", dict(type=str, default="", choices=["", "bf16", "fp32"])),
    ("--precision_vgg", dict(type=str, default="", choices=["", "bf16", "fp32"])),
    # extension: ImageNet VGG16 weights for the perceptual loss (torchvision vgg16 state_dict).  The reference downloads them
    # (vgg.py:8); without a file the VGG is random-initialised and a warning is printed.
    ("--vgg_weights", dict(type=str, default="")),
    # extension: replay optimize_parameters() as CUDA graphs after two eager warm-up steps (0 = always eager)
    ("--cuda_graph", dict(type=int, default=1)),
]


class BaseOptions:
    isTrain = True
    EXTRA_FLAGS = []

    def initialize(self, parser):
        for flag, kw in BASE_FLAGS + self.EXTRA_FLAGS:
            parser.add_argument(flag, **kw)
        return parser

    def gather_options(self, argv=None):
        from .. import models
        parser = self.initialize(argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter))
        opt, _unknown = parser.parse_known_args(argv)
        parser = models.get_option_setter(opt.model)(parser, self.isTrain)  # model-specific flags (--lambda_L1)
        self.parser = parser
        return parser.parse_args(argv)

    def print_options(self, opt):
        lines = ["----------------- Options ---------------"]
        for k, v in sorted(vars(opt).items()):
            default = self.parser.get_default(k)
            note = "\t[default: %s]" % str(default) if v != default else ""
            lines.append("{:>25}: {:<30}{}".format(str(k), str(v), note))
        lines.append("----------------- End -------------------")
        message = "\n".join(lines)
        print(message)
        expr_dir = os.path.join(opt.checkpoints_dir, opt.name)
        os.makedirs(expr_dir, exist_ok=True)
        with open(os.path.join(expr_dir, "opt.txt"), "wt") as f:
            f.write(message + "\n")

    def parse(self, dataset_path, path, argv=None, quiet=False):
        opt = self.gather_options(argv)
        opt.isTrain = self.isTrain
        opt.checkpoints_dir = os.path.join(path, "checkpoints")
        opt.dataroot = dataset_path
        if opt.suffix:
            opt.name = opt.name + "_" + opt.suffix.format(**vars(opt))
        if not quiet:
            self.print_options(opt)
        opt.gpu_ids = [int(s) for s in opt.gpu_ids.split(",") if int(s) >= 0]
        if opt.gpu_ids and torch.cuda.is_available():
            torch.cuda.set_device(opt.gpu_ids[0])
        self.opt = opt
        return opt
