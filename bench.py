#!/usr/bin/env python
"""bench.py — DS-GAN adversarial training step on B200 (BASELINE.json metric: train img/s at 256x256).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]

One "step" = one full optimize_parameters() (G forward, D step + Adam, G step through D/VGG/L1/TV/SSIM + Adam,
pix2pix_model.py:201-217) on a per-GPU batch of 16 synthetic 256x256 TIR/RGB pairs (BASELINE configs[1]); weak
scaling (16 images per GPU, global batch 128 at 8 GPUs = configs[2]).  Prints ONE JSON line (rank 0).
`--impl reference` times the CPU oracle (a port of the reference's own CPU path: /root/reference does not exist
on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train img/s at 256x256 (G+D optimize_parameters, device-timed)"
H = W = 256


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """CPU arm: the oracle's train_step (port of pix2pix_model.py:201-217) with all host threads."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dsgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = args.cpu_batch
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    A, B = O.synthetic_pair(b, H, W, seed=1)
    st = {}
    for _ in range(args.warmup):
        O.train_step(PG, PD, PV, A, B, st)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.train_step(PG, PD, PV, A, B, st)
    dt = time.perf_counter() - t0
    v = b * args.steps / dt
    sample = "%d steps of %d images (256x256) of the same G+D training step, fp32, torch CPU" % (args.steps, b)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DS-GAN training step (MixConvNeXtML G + PatchGAN D + VGG/L1/TV/SSIM losses, 2x Adam), "
                               "per-GPU batch %d, 256x256, random-init weights" % b,
                   "per_gpu_batch": b, "global_batch": b, "parallelism": "cpu",
                   "note": "CPU oracle (port of the reference's torch-CPU path, fp32) on rank 0's host cores; one step = the "
                           "same %d-image batch the GPU arm runs per GPU" % b},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(batch=16, steps=2):
    """The oracle (CPU port of the reference's own torch path) on the SAME per-step batch as the GPU arm, all host
    threads, bounded to 1 warm-up + `steps` timed steps (~25 s on 16 cores)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dsgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    A, B = O.synthetic_pair(batch, H, W, seed=1)
    st = {}
    O.train_step(PG, PD, PV, A, B, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(PG, PD, PV, A, B, st)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": "%d steps of %d images (256x256) of the same G+D training step, fp32, torch CPU" % (steps, batch)}


def extra_configs(ctx, hbm, tf_burst):
    """BASELINE configs[3] (MS-SSIM fwd+bwd sweep, 64x3x256x256, HBM GB/s) and configs[4] (generator-only inference,
    32x512x512) measured in the same run: CUDA events, 3 warm-ups, median of 5, L2 flushed between iterations."""
    import torch
    from dsgan_b200 import losses
    from dsgan_b200.models import networks
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def med(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    out = {}
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.rand(64, 3, 256, 256, device="cuda", generator=g)
    Y = (X + 0.1 * torch.randn(X.shape, device="cuda", generator=g)).clamp(0, 1)
    val, dY = torch.zeros(1, device="cuda"), torch.zeros_like(X)
    ms = med(lambda: losses.ssim_value_and_grad(ctx, X, Y, val.data_ptr(), 1.0, dY, 1.0, multiscale=True))
    nb = 5 * X.numel() * 4
    out["ms_ssim_fwd_bwd_64x3x256x256"] = {"ms": ms, "algorithmic_bytes": nb, "achieved_gbs": nb / (ms * 1e-3) / 1e9,
                                           "peak_gbs": hbm, "frac": nb / (ms * 1e-3) / 1e9 / hbm, "dtype": "f32"}
    del X, Y, dY
    was = networks.KernelNet.precision
    networks.KernelNet.precision = "bf16"
    G = networks.MixConvNeXtML().init_normal().cuda()
    networks.KernelNet.precision = was
    xin = torch.rand(32, 3, 512, 512, device="cuda", generator=g) * 2 - 1
    gctx = G.ctx()

    def infer():
        gctx.no_grad = True
        try:
            G(xin)
        finally:
            gctx.no_grad = False
            gctx.clear()
    ms = med(infer, iters=3, warm=2)
    fl = 2 * 40.77e9 * 4 * 32
    out["generator_inference_32x3x512x512"] = {"ms": ms, "img_per_s": 32 / (ms * 1e-3), "algorithmic_flops": fl,
                                               "achieved_tflops": fl / (ms * 1e-3) / 1e12, "peak_tflops": tf_burst,
                                               "frac": fl / (ms * 1e-3) / 1e12 / tf_burst, "dtype": "bf16"}
    return out


def traffic_note():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r2_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), beside its algorithmic bytes; None until a capture is committed."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from dsgan_b200 import engine
    from dsgan_b200._lib import lib
    from dsgan_b200.models import create_model
    from dsgan_b200.options.train_options import TrainOptions

    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch %d is not divisible by %d GPUs" % (args.global_batch, world))
        args.batch = args.global_batch // world
    opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_b200_bench",
                               argv=["--precision", args.precision, "--gpu_ids", str(local), "--batchSize", str(args.batch),
                                     "--cuda_graph", "0" if args.no_graph else "1"],
                               quiet=True)
    torch.manual_seed(20)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
        model.setup(opt)
    g = torch.Generator().manual_seed(1 + rank)
    b = args.batch
    A = (torch.rand(b, 1, H, W, generator=g) * 2 - 1).expand(b, 3, H, W).contiguous()
    B = torch.clamp(0.5 * A + 0.5 * (torch.rand(b, 3, H, W, generator=g) * 2 - 1), -1, 1)
    hostA, hostB = A.pin_memory(), B.pin_memory()
    batch = {"A": hostA, "B": hostB, "A_paths": [""] * b, "B_paths": [""] * b}
    L = lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident arm ("value") ---------------------------------------------------------
    model.set_input(batch)
    for _ in range(max(args.warmup, 0 if args.no_graph else 3)):   # (graph mode: two eager steps + the capture step are never timed)
        model.optimize_parameters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.cdll.dsgan_launch_count()
    ms = timed(model.optimize_parameters, args.steps)
    launches = L.cdll.dsgan_launch_count() - n0
    gs = getattr(model, "_gs", None)
    graphed = gs is not None and gs.get("plan") is not None
    if graphed:   # replayed steps launch the kernels recorded at capture; the library's counter only sees eager launches
        launches += args.steps * gs["kernels_per_replay"]
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: host batch in, losses out, every step -------------------------------------
    def e2e_step():
        model.set_input(batch)             # (takes the copy prefetch_input() started during the previous step, if any)
        model.optimize_parameters()
        model.prefetch_input(batch)        # next batch: pinned host -> device on the copy stream, under this step's kernels
        model.get_current_losses()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # ---- roofline of the dominant kernel family (dense conv / GEMM), timed live with CUDA events ----
    ctx = model.ctx
    # (one extra step with every launch bracketed by CUDA events on its stream; the side stream is disabled for it so that
    # each kernel is timed alone instead of sharing the SMs with the concurrent branch)
    ctx.profile = engine.Profile()
    streams, ctx.use_streams = ctx.use_streams, False
    model.optimize_parameters()
    torch.cuda.synchronize()
    ctx.use_streams = streams
    prof = ctx.profile.summary()
    if rank == 0 and args.detail:
        for ms_, n_, tf_, name_ in ctx.profile.detail(400):
            print("%9.3f ms  x%-4d %8.1f TF/s  %s" % (ms_, n_, tf_, name_), file=sys.stderr)
    ctx.profile = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sust, which = peaks()
    imgs = b * world
    dense = prof["dense"]
    achieved = dense["flops"] / (dense["ms"] * 1e-3) / 1e12 if dense["ms"] > 0 else 0.0
    out = {
        "metric": METRIC, "value": imgs * args.steps / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "DS-GAN training step (MixConvNeXtML G + PatchGAN D + VGG/L1/TV/SSIM losses, 2x Adam), "
                               "per-GPU batch %d, 256x256, random-init weights" % b,
                   "per_gpu_batch": b, "global_batch": imgs, "parallelism": "dp%d" % world,
                   "l2_note": "per-step working set (activations, >5 GB) far exceeds the 126 MB L2",
                   "launch": "CUDA-graph replay of the step (image pool / NCCL eager between segments)" if graphed
                             else "eager launches"},
        "e2e": {"value": imgs * args.steps / (ms_e2e * 1e-3), "unit": "img/s",
                "h2d_bytes_per_step": int(hostA.numel() * 4 + hostB.numel() * 4), "d2h_bytes_per_step": 16},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": dense["name"], "achieved": achieved, "peak": tf_sust,
                     "unit": "TFLOP/s", "frac": achieved / tf_sust, "peak_source": which + " (sustained bf16)",
                     "traffic": traffic_note(), "launches_per_step": dense["n"], "ms_per_step": dense["ms"],
                     "share_of_step": dense["ms"] / (ms / args.steps),
                     "families": {k: {"ms": v["ms"], "n": v["n"]} for k, v in prof.items()}},
    }
    out["parity"] = {"mode": "bf16 activations / fp32 accumulate, statistics, losses and master weights",
                     "meets": "losses <= 1e-3, fake_B <= 3e-2, G gradients <= 3e-2 (whole-net, configs[1])",
                     "open": "D gradients 6-8e-2 end to end: fake_B's bf16 error x the ~3-20x conditioning of the "
                             "real/fake cancellation in D's loss (profiles/r2_error_budget.json)"}
    if world == 1 and not args.no_extra:
        out["extra_configs"] = extra_configs(ctx, hbm, tf_burst)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(b if b <= 16 else 16)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under torchrun), so the
    # process-level stdout is pointed at stderr for the whole run and the result line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    for l in out.getvalue().splitlines():
        if not l.startswith("{"):
            print(l, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (BASELINE configs[1] = 16)")
    ap.add_argument("--cpu-batch", dest="cpu_batch", type=int, default=16,
                    help="images per CPU reference step (default: the GPU arm's per-GPU batch)")
    ap.add_argument("--global-batch", dest="global_batch", type=int, default=0,
                    help="strong scaling: fixed global batch split over the GPUs (BASELINE configs[2]: 128)")
    ap.add_argument("--no-extra", dest="no_extra", action="store_true", help="skip the configs[3]/[4] side measurements")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--detail", action="store_true", help="print the per-kernel profile to stderr")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
