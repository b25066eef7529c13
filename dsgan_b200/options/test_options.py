"""Test flags (DSGAN/options/test_options.py:5-14), same names and defaults."""
from .base_options import INF, BaseOptions


class TestOptions(BaseOptions):
    isTrain = False
    EXTRA_FLAGS = [
        ("--ntest", dict(type=int, default=INF)),
        ("--results_dir", dict(type=str, default="epoch_8_result_original/")),
        ("--aspect_ratio", dict(type=float, default=1.0)),
        ("--phase", dict(type=str, default="test_all/")),
        ("--which_epoch", dict(type=str, default="1")),
        ("--how_many", dict(type=int, default=1000)),
    ]
