"""Host-side API parity that needs no GPU: ImagePool against a trace recorded from the reference's own
util/image_pool.py (tests/golden/api.json, oracle/make_golden_api.py), state_dict inventories, VGG weight loading."""
import json
import os
import random

import torch

from dsgan_b200 import specs
from dsgan_b200.models.vgg import Vgg16
from dsgan_b200.util.image_pool import ImagePool


def _api(golden_dir):
    return json.load(open(os.path.join(golden_dir, "api.json")))


def test_image_pool_matches_reference_trace(golden_dir):
    """Seeded python `random`: identical returned/stored images, including the swap phase after 50 images
    (util/image_pool.py:12-32)."""
    rec = _api(golden_dir)["image_pool"]
    random.seed(rec["seed"])
    pool = ImagePool(rec["pool_size"])
    nxt = 0
    for bs, want in zip(rec["batches"], rec["returned"]):
        imgs = torch.stack([torch.full((6, 2, 2), float(nxt + i)) for i in range(bs)])
        nxt += bs
        got = pool.query(imgs)
        assert got.shape == imgs.shape
        assert [int(v) for v in got[:, 0, 0, 0].tolist()] == want
    assert [int(t[0, 0, 0, 0]) for t in pool.images] == rec["stored"]
    assert pool.num_imgs == rec["pool_size"]
    assert any(w != list(range(s, s + len(w))) for w, s in zip(rec["returned"][4:], (64, 80, 87, 88, 104))), \
        "trace never left the identity phase"


def test_image_pool_size_zero_is_identity():
    x = torch.randn(3, 6, 4, 4)
    assert ImagePool(0).query(x) is x


def test_state_dict_inventories_match_reference(golden_dir):
    """(name, shape) of the generator / discriminator state_dicts as the reference's own modules report them."""
    rec = _api(golden_dir)
    assert [[n, list(s)] for n, s in specs.generator_spec()] == rec["G_keys"]
    assert [[n, list(s)] for n, s in specs.discriminator_spec(6, 8)] == rec["D_keys"]


def test_vgg_loads_torchvision_names():
    """`--vgg_weights`: a torchvision vgg16 state_dict (features.N.*) lands under the reference's regrouped names
    (vgg.py:16-25); classifier keys are ignored, a missing conv raises."""
    v = Vgg16()
    g = torch.Generator().manual_seed(3)
    idx = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
    shapes = [(64, 3), (64, 64), (128, 64), (128, 128), (256, 128), (256, 256), (256, 256), (512, 256), (512, 512),
              (512, 512), (512, 512), (512, 512), (512, 512)]
    sd = {}
    for i, (o, c) in zip(idx, shapes):
        sd["features.%d.weight" % i] = torch.randn(o, c, 3, 3, generator=g)
        sd["features.%d.bias" % i] = torch.randn(o, generator=g)
    sd["classifier.0.weight"] = torch.zeros(4, 4)
    v.load_torchvision(sd)
    got = v.state_dict()
    assert torch.equal(got["to_relu_1_2.0.weight"], sd["features.0.weight"])
    assert torch.equal(got["to_relu_3_3.14.bias"], sd["features.14.bias"])
    assert torch.equal(got["to_relu_5_3.28.weight"], sd["features.28.weight"])
    del sd["features.7.weight"]
    try:
        Vgg16().load_torchvision(sd)
    except KeyError:
        pass
    else:
        raise AssertionError("missing conv was accepted")
