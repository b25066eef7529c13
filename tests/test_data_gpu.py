"""Device-side input pipeline against the reference's CPU transforms (data/aligned_dataset.py:53-90), bit for bit."""
import random
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from dsgan_b200.data import DeviceInputPipeline, draw_augment  # noqa: E402


def _reference_item(img_u8, h_off, w_off, flip, fine_h, fine_w, nc):
    """ToTensor -> crop -> Normalize(0.5, 0.5) -> flip -> gray, with plain torch ops in the reference's order."""
    t = img_u8.permute(2, 0, 1).float().div(255)
    t = t[:, h_off:h_off + fine_h, w_off:w_off + fine_w]
    t = t.sub(0.5).div(0.5)
    if flip:
        t = t.index_select(2, torch.arange(t.size(2) - 1, -1, -1))
    if nc == 1:
        t = (t[0] * 0.299 + t[1] * 0.587 + t[2] * 0.114).unsqueeze(0)
    return t


@pytest.mark.parametrize("load,fine,nc_in", [((286, 300), (256, 256), 3), ((256, 256), (256, 256), 3), ((70, 90), (64, 80), 1)])
def test_device_pipeline_matches_reference_transforms(load, fine, nc_in):
    opt = SimpleNamespace(loadSize_h=load[0], loadSize_w=load[1], fineSize_h=fine[0], fineSize_w=fine[1], no_flip=False,
                          which_direction="AtoB", input_nc=nc_in, output_nc=3)
    g = torch.Generator().manual_seed(3)
    n = 5
    A = torch.randint(0, 256, (n, load[0], load[1], 3), dtype=torch.uint8, generator=g)
    B = torch.randint(0, 256, (n, load[0], load[1], 3), dtype=torch.uint8, generator=g)
    pipe = DeviceInputPipeline(opt, "cuda:0", rng=random.Random(11))
    aug = draw_augment(opt, n, random.Random(11))
    for _rep in range(3):      # exercises both staging slots and their reuse
        out = pipe(A.numpy(), B.numpy(), ["a%d" % i for i in range(n)], ["b%d" % i for i in range(n)], augment=aug)
        torch.cuda.synchronize()
        assert out["A"].shape == (n, nc_in, fine[0], fine[1]) and out["B"].shape == (n, 3, fine[0], fine[1])
        assert out["A_paths"][2] == "a2"
        for i in range(n):
            wa = _reference_item(A[i], aug[0][i], aug[1][i], aug[2][i], fine[0], fine[1], nc_in)
            wb = _reference_item(B[i], aug[0][i], aug[1][i], aug[2][i], fine[0], fine[1], 3)
            assert torch.equal(out["A"][i].cpu(), wa), "A differs from ToTensor/crop/Normalize/flip"
            assert torch.equal(out["B"][i].cpu(), wb)
    assert any(aug[2]) and not all(aug[2])
    with pytest.raises(ValueError):
        pipe(A.float().numpy(), B.numpy())


def test_pipeline_feeds_the_training_step():
    from dsgan_b200.models import create_model
    from dsgan_b200.options.train_options import TrainOptions
    opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_b200_data", argv=["--loadSize_w", "72", "--loadSize_h", "72",
                                                                         "--fineSize_w", "64", "--fineSize_h", "64"], quiet=True)
    m = create_model(opt)
    m.setup(opt)
    pipe = DeviceInputPipeline(opt, m.device)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):
        A = torch.randint(0, 256, (2, 72, 72, 3), dtype=torch.uint8, generator=g)
        B = torch.randint(0, 256, (2, 72, 72, 3), dtype=torch.uint8, generator=g)
        m.set_input(pipe(A, B))
        m.optimize_parameters()
    losses = m.get_current_losses()
    assert m.real_A.shape == (2, 3, 64, 64) and float(m.real_A.min()) >= -1 and float(m.real_A.max()) <= 1
    assert all(torch.isfinite(torch.tensor(list(losses.values()))))
