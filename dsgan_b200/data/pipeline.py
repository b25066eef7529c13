"""GPU-side input pipeline.

The reference feeds the step from a 4-worker PIL DataLoader that does ToTensor / crop / Normalize / flip per image on
the host in fp32 (data/aligned_dataset.py:53-76) and ships 12 bytes per pixel pair.  At several thousand images/s per
box that loader is the bottleneck, so the per-pixel work moves to the device:

  host:   decoded uint8 HWC images (what PIL / cv2 hand over) -> one pinned staging buffer per batch (double buffered)
  copy:   H2D on a side stream, 3 bytes per pixel pair instead of 24
  device: dsgan_preprocess_u8 = ToTensor -> crop -> Normalize(0.5, 0.5) -> flip (-> gray), bit-exact with the reference's ops,
          straight into the fp32 NCHW tensors `Pix2PixModel.set_input` takes.

`ShardedBatchSampler` replaces DataLoader(shuffle=not serial_batches) under one-process-per-GPU data parallelism: every
rank draws the same epoch permutation and takes its contiguous slice of each global batch (parallel.shard_batch layout).
Decoding files stays with the caller (any loader that yields uint8 arrays works); nothing here touches the disk."""
import random

import torch

from .._lib import lib


def draw_augment(opt, n, rng=random):
    """Per-image crop offsets and flip decisions with the reference's draws, in its order (aligned_dataset.py:56-74):
    w_offset, h_offset = randint(0, max(0, load - fine - 1)); flip = (not no_flip) and random() < 0.5."""
    h_off, w_off, flip = [], [], []
    for _ in range(n):
        w_off.append(rng.randint(0, max(0, opt.loadSize_w - opt.fineSize_w - 1)))
        h_off.append(rng.randint(0, max(0, opt.loadSize_h - opt.fineSize_h - 1)))
        flip.append(int((not opt.no_flip) and rng.random() < 0.5))
    return h_off, w_off, flip


class ShardedBatchSampler:
    """Index batches for rank `rank` of `world`: one shared permutation per epoch (seed + epoch), global batches of
    `batch_size * world` consecutive entries, rank r takes entries [r*batch_size, (r+1)*batch_size) of each.  The tail that
    does not fill a global batch is wrapped around (every rank runs the same number of steps)."""

    def __init__(self, n_items, batch_size, rank=0, world=1, shuffle=True, seed=20, max_items=float("inf")):
        if n_items < 1 or batch_size < 1 or not (0 <= rank < world):
            raise ValueError("bad sampler arguments")
        self.n = int(min(n_items, max_items))
        self.batch_size, self.rank, self.world, self.shuffle, self.seed = batch_size, rank, world, shuffle, seed
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        g = self.batch_size * self.world
        return (self.n + g - 1) // g

    def __iter__(self):
        if self.shuffle:
            gen = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=gen).tolist()
        else:
            order = list(range(self.n))
        g = self.batch_size * self.world
        total = len(self) * g
        order = (order * ((total + self.n - 1) // self.n))[:total]
        for b in range(len(self)):
            lo = b * g + self.rank * self.batch_size
            yield order[lo:lo + self.batch_size]


class DeviceInputPipeline:
    """uint8 batches -> the {'A', 'B', 'A_paths', 'B_paths'} dict of device tensors set_input() takes.

    pipe = DeviceInputPipeline(opt, device)
    batch = pipe(A_u8, B_u8, paths_A, paths_B)        # A_u8, B_u8: uint8 [N, loadSize_h, loadSize_w, 3] (numpy or torch, host)
    model.set_input(batch)
    """

    def __init__(self, opt, device, rng=random):
        self.opt, self.device, self.rng = opt, torch.device(device), rng
        self.stream = torch.cuda.Stream(self.device)
        self._slots, self._i = [None, None], 0
        self.in_nc = opt.output_nc if opt.which_direction == "BtoA" else opt.input_nc
        self.out_nc = opt.input_nc if opt.which_direction == "BtoA" else opt.output_nc

    def _slot(self, shape):
        s = self._slots[self._i]
        if s is None or s["shape"] != shape:
            n, h, w, _ = shape
            s = {"shape": shape,
                 "host": torch.empty((2,) + shape, dtype=torch.uint8).pin_memory(),
                 "meta_host": torch.empty((3, n), dtype=torch.int32).pin_memory(),
                 "dev": torch.empty((2,) + shape, dtype=torch.uint8, device=self.device),
                 "meta": torch.empty((3, n), dtype=torch.int32, device=self.device),
                 "done": torch.cuda.Event()}
            self._slots[self._i] = s
        self._i ^= 1
        return s

    def __call__(self, A_u8, B_u8, A_paths=None, B_paths=None, augment=None):
        A_u8, B_u8 = torch.as_tensor(A_u8), torch.as_tensor(B_u8)
        if A_u8.dtype != torch.uint8 or A_u8.dim() != 4 or A_u8.shape[-1] != 3 or A_u8.shape != B_u8.shape:
            raise ValueError("expected two uint8 [N, H, W, 3] batches of equal shape, got %s / %s %s" %
                             (tuple(A_u8.shape), tuple(B_u8.shape), A_u8.dtype))
        o = self.opt
        n, hs, ws, _ = A_u8.shape
        H, W = min(o.fineSize_h, hs), min(o.fineSize_w, ws)
        h_off, w_off, flip = augment if augment is not None else draw_augment(o, n, self.rng)
        if max(h_off) + H > hs or max(w_off) + W > ws:
            raise ValueError("crop window leaves the %dx%d source" % (hs, ws))
        s = self._slot(tuple(A_u8.shape))
        s["done"].synchronize()                       # the slot's previous batch has left the staging buffers
        s["host"][0].copy_(A_u8)
        s["host"][1].copy_(B_u8)
        s["meta_host"].copy_(torch.tensor([h_off, w_off, flip], dtype=torch.int32))
        L = lib()
        out = {"A": torch.empty((n, self.in_nc, H, W), dtype=torch.float32, device=self.device),
               "B": torch.empty((n, self.out_nc, H, W), dtype=torch.float32, device=self.device),
               "A_paths": list(A_paths) if A_paths is not None else [""] * n,
               "B_paths": list(B_paths) if B_paths is not None else [""] * n}
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)                  # output tensors were allocated on the current stream
        with torch.cuda.stream(self.stream):
            s["dev"].copy_(s["host"], non_blocking=True)
            s["meta"].copy_(s["meta_host"], non_blocking=True)
            m = s["meta"]
            for k, (key, nc) in enumerate((("A", self.in_nc), ("B", self.out_nc))):
                L.preprocess_u8(s["dev"][k].data_ptr(), n, hs, ws, m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(),
                                out[key].data_ptr(), nc, H, W, self.stream.cuda_stream)
            s["done"].record(self.stream)
        cur.wait_stream(self.stream)                  # consumers on the current stream see finished tensors
        return out
