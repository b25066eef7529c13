"""Training flags (DSGAN/options/train_options.py:5-28), same names and defaults."""
from .base_options import BaseOptions


class TrainOptions(BaseOptions):
    isTrain = True
    EXTRA_FLAGS = [
        ("--display_freq", dict(type=int, default=100)), ("--display_ncols", dict(type=int, default=4)),
        ("--update_html_freq", dict(type=int, default=1000)), ("--print_freq", dict(type=int, default=100)),
        ("--save_latest_freq", dict(type=int, default=5000)), ("--save_epoch_freq", dict(type=int, default=50)),
        ("--continue_train", dict(action="store_true", default=False)),
        ("--epoch_count", dict(type=int, default=1)),
        ("--phase", dict(type=str, default="train_all/")),
        ("--which_epoch", dict(type=str, default="1")),
        ("--niter", dict(type=int, default=10)), ("--niter_decay", dict(type=int, default=10)),
        ("--beta1", dict(type=float, default=0.5)), ("--lr", dict(type=float, default=0.0002)),
        ("--no_lsgan", dict(action="store_true")),
        ("--pool_size", dict(type=int, default=50)),
        ("--no_html", dict(action="store_true")),
        ("--lr_policy", dict(type=str, default="lambda")),
        ("--lr_decay_iters", dict(type=int, default=50)),
    ]
