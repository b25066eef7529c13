"""Input pipeline for the training step (reference: DSGAN/data/__init__.py:31-63, data/aligned_dataset.py:37-90)."""
from .pipeline import DeviceInputPipeline, ShardedBatchSampler, draw_augment  # noqa: F401
