"""Host-side launch time vs device time of one optimize_parameters() (is the step launch-bound?)."""
import contextlib, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsgan_b200.models import create_model
from dsgan_b200.options.train_options import TrainOptions

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_b200_bench", argv=["--precision", "bf16", "--gpu_ids", "0", "--batchSize", str(b)], quiet=True)
torch.manual_seed(20)
with contextlib.redirect_stdout(io.StringIO()):
    model = create_model(opt); model.setup(opt)
g = torch.Generator().manual_seed(1)
A = (torch.rand(b, 1, 256, 256, generator=g) * 2 - 1).expand(b, 3, 256, 256).contiguous()
B = torch.clamp(0.5 * A + 0.5 * (torch.rand(b, 3, 256, 256, generator=g) * 2 - 1), -1, 1)
model.set_input({"A": A, "B": B, "A_paths": [""] * b, "B_paths": [""] * b})
for _ in range(3):
    model.optimize_parameters()
torch.cuda.synchronize()
for streams in (True, False):
    model.ctx.use_streams = streams
    model.optimize_parameters(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model.optimize_parameters()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("batch %d streams=%s: host issue %.2f ms/step, host+drain %.2f ms/step" % (b, streams, (t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
