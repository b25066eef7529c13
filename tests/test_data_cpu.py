"""Host logic of the input pipeline (no GPU): the sharded sampler and the reference's augmentation draws."""
import random
from types import SimpleNamespace

from dsgan_b200.data import ShardedBatchSampler, draw_augment


def test_sampler_shards_cover_every_global_batch_once():
    n, bs, world = 103, 4, 4
    per_rank = [list(ShardedBatchSampler(n, bs, r, world, shuffle=True, seed=7)) for r in range(world)]
    assert len({len(p) for p in per_rank}) == 1 and len(per_rank[0]) == (n + bs * world - 1) // (bs * world)
    seen = []
    for step in range(len(per_rank[0])):
        glob = [i for r in range(world) for i in per_rank[r][step]]
        assert len(glob) == bs * world
        seen += glob
    assert set(seen) == set(range(n))                        # every image is visited
    assert len(seen) - n < bs * world                        # only the wrapped tail repeats
    # deterministic per epoch, different across epochs, identical permutation on every rank
    s = ShardedBatchSampler(n, bs, 1, world, seed=7)
    assert list(s) == per_rank[1]
    s.set_epoch(1)
    assert list(s) != per_rank[1]


def test_sampler_serial_batches_and_max_dataset_size():
    s = ShardedBatchSampler(10, 2, 0, 1, shuffle=False)
    assert list(s) == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]
    assert len(ShardedBatchSampler(1000, 16, 0, 1, max_items=100)) == 7
    r1 = list(ShardedBatchSampler(8, 2, 1, 2, shuffle=False))
    assert r1 == [[2, 3], [6, 7]]


def test_draw_augment_follows_the_reference_draw_order():
    """aligned_dataset.py:56-74: per image w_offset, h_offset (randint, inclusive) then the flip draw."""
    opt = SimpleNamespace(loadSize_w=286, fineSize_w=256, loadSize_h=270, fineSize_h=256, no_flip=False)
    random.seed(5)
    want = []
    for _ in range(6):
        w = random.randint(0, max(0, 286 - 256 - 1))
        h = random.randint(0, max(0, 270 - 256 - 1))
        f = int(random.random() < 0.5)
        want.append((h, w, f))
    random.seed(5)
    h_off, w_off, flip = draw_augment(opt, 6)
    assert list(zip(h_off, w_off, flip)) == want
    assert max(w_off) <= 29 and max(h_off) <= 13
    opt.no_flip = True
    assert draw_augment(opt, 4)[2] == [0, 0, 0, 0]
    opt = SimpleNamespace(loadSize_w=256, fineSize_w=256, loadSize_h=256, fineSize_h=256, no_flip=True)
    assert draw_augment(opt, 3) == ([0, 0, 0], [0, 0, 0], [0, 0, 0])
