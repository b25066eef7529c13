"""Generate the API-level fixtures of tests/golden/ by running the UNMODIFIED reference (build container only):

  * ref_saved_5_useSE_net_D.pth  -- written by the reference's own BaseModel.save_networks (base_model.py:92-103) for a
    small discriminator (ndf=8, 44 k parameters) holding seeded weights; api.json stores their fingerprints.  Proves that
    dsgan_b200's load_networks reads a reference-written checkpoint.
  * api.json["image_pool"]       -- the ids returned by the reference's ImagePool.query (util/image_pool.py:12-32) for a
    seeded python `random` stream, well past the 50-image fill phase.
  * api.json["G_keys"] / ["D_keys"] -- (name, shape) of the reference networks' state_dicts.

usage:  python oracle/make_golden_api.py
"""
import json
import os
import random
import shutil
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dsgan_oracle as O  # noqa: E402
from make_golden import import_reference, make_opt  # noqa: E402


def main():
    out = os.path.join(os.path.dirname(HERE), "tests", "golden")
    p2p, _ = import_reference()
    from util.image_pool import ImagePool
    rec = {}
    # ---- checkpoint written by the reference ----
    tmp = "/tmp/dsgan_golden_api"
    shutil.rmtree(tmp, ignore_errors=True)
    opt = make_opt(tmp)
    opt.ndf = 8
    torch.manual_seed(7)
    model = p2p.Pix2PixModel()
    model.initialize(opt)
    rec["G_keys"] = [[k, list(v.shape)] for k, v in model.netG.state_dict().items()]
    rec["D_keys"] = [[k, list(v.shape)] for k, v in model.netD.state_dict().items()]
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in model.netD.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    rec["D_fingerprints"] = {k: O.fingerprint(v) for k, v in model.netD.state_dict().items()}
    os.makedirs(model.save_dir, exist_ok=True)
    model.model_names = ["D"]
    model.save_networks(5)
    shutil.copy(os.path.join(model.save_dir, "5_useSE_net_D.pth"), os.path.join(out, "ref_saved_5_useSE_net_D.pth"))
    # ---- ImagePool trace ----
    random.seed(1234)
    pool = ImagePool(50)
    trace, nxt = [], 0
    for bs in [16, 16, 16, 16, 16, 7, 1, 16, 16]:
        imgs = torch.stack([torch.full((6, 2, 2), float(nxt + i)) for i in range(bs)])
        nxt += bs
        res = pool.query(imgs)
        trace.append([int(v) for v in res[:, 0, 0, 0].tolist()])
    rec["image_pool"] = {"seed": 1234, "pool_size": 50, "batches": [16, 16, 16, 16, 16, 7, 1, 16, 16], "returned": trace,
                         "stored": [int(t[0, 0, 0, 0]) for t in pool.images]}
    with open(os.path.join(out, "api.json"), "w") as f:
        json.dump(rec, f)
    print("api golden written:", os.path.getsize(os.path.join(out, "ref_saved_5_useSE_net_D.pth")), "bytes checkpoint")


if __name__ == "__main__":
    main()
