"""Data-parallel plumbing (one process per GPU, torch.distributed; replaces nn.DataParallel, networks.py:73-79).

The path shards by samples: every rank holds full replicas of G, D, VGG and both Adam states, runs the whole step on its
slice of the batch, and exchanges only gradients — one all-reduce (SUM) over each network's flat gradient buffer, averaged
by the fused Adam (`grad_scale = 1/world`).  InstanceNorm is per-sample, so no statistics cross ranks.

Keeping R-rank training identical to single-GPU big-batch math needs one correction: every loss of the step is a batch
MEAN except the TV term, which is a batch SUM over a constant divisor (pix2pix_model.py:189-191).  After gradient
averaging a sum-type term would be divided by R, so its gradient is pre-multiplied by R (`tv_grad_scale`)."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_batch(batch, r=None, world=None):
    """Rank r's contiguous slice [r*B/R, (r+1)*B/R) of a {'A','B','A_paths','B_paths'} batch dict."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    n = batch["A"].shape[0]
    if n % world:
        raise ValueError("global batch %d is not divisible by the world size %d" % (n, world))
    lo, hi = r * n // world, (r + 1) * n // world
    out = {}
    for k, v in batch.items():
        out[k] = v[lo:hi] if isinstance(v, (torch.Tensor, list, tuple)) and len(v) == n else v
    return out


def tv_grad_scale(world=None):
    return float(world_size() if world is None else world)


def adam_grad_scale(world=None):
    return 1.0 / float(world_size() if world is None else world)


def allreduce_grads(flat_grad):
    """SUM all-reduce of one network's flat gradient buffer (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    if world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def broadcast_params(flat_params, src=0):
    if world_size() > 1:
        dist.broadcast(flat_params, src)
    return flat_params


def barrier():
    if world_size() > 1:
        dist.barrier()
