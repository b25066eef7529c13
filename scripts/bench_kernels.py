"""Per-kernel rooflines on one B200 (run under gpurun): CUDA-event timing on torch's current stream, 3 warm-ups, every
timed launch works on tensors far larger than the 126 MB L2 or is preceded by an L2 flush.  Prints one JSON line per case.

  python scripts/bench_kernels.py [--out gpurun_out/kernels.jsonl]

Covers BASELINE configs #4 (MS-SSIM fwd+bwd, 64x3x256x256) and #5 (generator inference 32x512x512), the InstanceNorm /
depthwise bandwidth kernels and the tcgen05 GEMM / implicit-conv kernels at the layer shapes of the training step."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsgan_b200 import MS_SSIM, engine as E, losses  # noqa: E402
from dsgan_b200.engine import Ctx, Param  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


ITERS, WARM = None, None


def timeit(fn, iters=10, warm=3, flush=True):
    iters = ITERS if ITERS is not None else iters
    warm = WARM if WARM is not None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]   # median: robust against a one-off stall (allocator growth, lazy module load)


def graphed(fn):
    """Capture fn into a CUDA graph (after eager warm-up) so the timing excludes host launch overhead."""
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def emit(out, name, ms, bytes_=None, flops=None, **kw):
    rec = {"kernel": name, "ms": round(ms, 4)}
    if bytes_ is not None:
        gbs = bytes_ / (ms * 1e-3) / 1e9
        rec.update(bound="hbm", algorithmic_bytes=int(bytes_), achieved_gbs=round(gbs, 1), peak_gbs=PEAK["hbm_gbs"],
                   frac=round(gbs / PEAK["hbm_gbs"], 3))
    if flops is not None:
        tf = flops / (ms * 1e-3) / 1e12
        rec.update(bound="tensor", algorithmic_flops=float(flops), achieved_tflops=round(tf, 1), peak_tflops=PEAK["bf16_tflops"],
                   frac=round(tf / PEAK["bf16_tflops"], 3))
    rec.update(kw)
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")


def params(shapes):
    P = {}
    for k, shp in shapes.items():
        d = (torch.randn(shp, device="cuda") * 0.05)
        sh = d.bfloat16()
        P[k] = Param(k, d, torch.zeros_like(d), sh.data_ptr())
        P[k].cache["bf16"] = sh
    return P


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.jsonl"))
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=None, help="override timed iterations (1 for ncu captures)")
    ap.add_argument("--warm", type=int, default=None)
    args = ap.parse_args()
    global ITERS, WARM
    ITERS, WARM = args.iters, args.warm
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    ctx = Ctx("cuda:0", "bf16")
    g = torch.Generator(device="cuda").manual_seed(1)
    want = lambda tag: (not args.only) or args.only in tag

    # ---- config #4: MS-SSIM / SSIM fwd + bwd (w.r.t. Y), 64x3x256x256 fp32 ------------------------------------
    if want("ssim"):
        X = torch.rand(64, 3, 256, 256, device="cuda", generator=g)
        Y = (X + 0.1 * torch.randn(X.shape, device="cuda", generator=g)).clamp(0, 1)
        nbytes = X.numel() * 4
        for multi, name in ((True, "ms_ssim"), (False, "ssim")):
            val = torch.zeros(1, device="cuda")
            dY = torch.zeros_like(Y)

            def fwd_bwd():
                losses.ssim_value_and_grad(ctx, X, Y, val.data_ptr(), 1.0, dY, 1.0, multiscale=multi)

            def fwd():
                losses.ssim_value_and_grad(ctx, X, Y, val.data_ptr(), 1.0, None, 0.0, multiscale=multi)
            emit(out, name + " fwd+bwd 64x3x256x256 fp32 (config #4)" if multi else name + " fwd+bwd 64x3x256x256 fp32",
                 timeit(graphed(fwd_bwd)), bytes_=5 * nbytes,
                 note="algorithmic bytes = 5*N*C*H*W*4 (SURVEY 8d); CUDA-graph replay of the whole call (value + dY)")
            emit(out, name + " fwd 64x3x256x256 fp32", timeit(graphed(fwd)), bytes_=2 * nbytes)
            emit(out, name + " fwd+bwd, eager host launches", timeit(fwd_bwd), bytes_=5 * nbytes)
        # the two level-0 kernels on their own (ABI calls)
        L, st = ctx.L, ctx.stream
        NC, Hh, Ww = 64 * 3, 256, 256
        sums = torch.zeros(NC, 2, device="cuda")
        coef = torch.full((NC, 2), 1e-5, device="cuda")
        mom = torch.empty(5, NC, Hh - 10, Ww - 10, device="cuda")
        dY = torch.zeros_like(Y)
        emit(out, "ssim_fwd kernel level 0 (no moments)", timeit(lambda: L.ssim_fwd(X.data_ptr(), Y.data_ptr(), NC, Hh, Ww, 1e-4, 9e-4, sums.data_ptr(), None, st)), bytes_=2 * nbytes)
        emit(out, "ssim_fwd kernel level 0 (+ moments stored)", timeit(lambda: L.ssim_fwd(X.data_ptr(), Y.data_ptr(), NC, Hh, Ww, 1e-4, 9e-4, sums.data_ptr(), mom.data_ptr(), st)), bytes_=2 * nbytes)
        emit(out, "ssim_bwd kernel level 0 (from moments)", timeit(lambda: L.ssim_bwd(X.data_ptr(), Y.data_ptr(), NC, Hh, Ww, 1e-4, 9e-4, coef.data_ptr(), mom.data_ptr(), dY.data_ptr(), 0, None, st)), bytes_=3 * nbytes)
        emit(out, "ssim_bwd kernel level 0 (tiled, moments recomputed)", timeit(lambda: L.ssim_bwd(X.data_ptr(), Y.data_ptr(), NC, Hh, Ww, 1e-4, 9e-4, coef.data_ptr(), None, dY.data_ptr(), 0, None, st)), bytes_=3 * nbytes)

    # ---- InstanceNorm family on the uc4 / u4 tensor: 16 x 256 x 256 x 128 bf16 ---------------------------------
    if want("inorm"):
        x = ctx.new(16, 256, 256, 128)
        x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g))
        nb = x.t.numel() * 2
        stats = ctx.f32(16, 128, 3)
        bst = ctx.f32(16, 128, 2)
        y, dy, dx = ctx.new(16, 256, 256, 128), ctx.new(16, 256, 256, 128), ctx.new(16, 256, 256, 128)
        dy.t.copy_(torch.randn(dy.t.shape, device="cuda", generator=g))
        L, s, HW = ctx.L, ctx.stream, 256 * 256
        emit(out, "inorm_stats 16x256x256x128 bf16", timeit(lambda: L.inorm_stats(x.ptr, 128, 1, 16, HW, 128, stats.data_ptr(), s)), bytes_=nb)
        emit(out, "inorm_apply+GELU", timeit(lambda: L.inorm_apply(x.ptr, 128, stats.data_ptr(), None, 0, y.ptr, 128, 1, 16, HW, 128, 3, s)), bytes_=2 * nb)
        emit(out, "inorm_bwd_stats (GELU)", timeit(lambda: L.inorm_bwd_stats(x.ptr, 128, stats.data_ptr(), None, 0, dy.ptr, 128, 1, 16, HW, 128, 3, bst.data_ptr(), s)), bytes_=2 * nb)
        emit(out, "inorm_bwd_apply (GELU)", timeit(lambda: L.inorm_bwd_apply(x.ptr, 128, stats.data_ptr(), None, 0, dy.ptr, 128, bst.data_ptr(), dx.ptr, 128, 0, None, 0, 0, 1, 16, HW, 128, 3, None, None, s)), bytes_=3 * nb)
        del x, y, dy, dx

    # ---- depthwise convolutions at the Block shapes (bf16 NHWC) -----------------------------------------------
    if want("dwconv"):
        L, s = ctx.L, ctx.stream
        for (Nn, Hh, Cc, kk) in ((16, 256, 128, 7), (16, 128, 256, 7), (16, 64, 512, 7), (16, 128, 32, 9)):
            x, y, dy, dx = (ctx.new(Nn, Hh, Hh, Cc) for _ in range(4))
            x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g))
            dy.t.copy_(torch.randn(dy.t.shape, device="cuda", generator=g))
            nb = x.t.numel() * 2
            w = torch.randn(Cc, 1, kk, kk, device="cuda") * 0.1
            b = torch.zeros(Cc, device="cuda")
            dw, db = torch.zeros_like(w), torch.zeros_like(b)
            tag = "dwconv%d %dx%dx%dx%d bf16 " % (kk, Nn, Hh, Hh, Cc)
            emit(out, tag + "fwd", timeit(lambda: L.dwconv_fwd(x.ptr, Cc, w.data_ptr(), b.data_ptr(), y.ptr, Cc, 1, Nn, Hh, Hh, Cc, kk, 0, 0, s)), bytes_=2 * nb)
            emit(out, tag + "dgrad (flip, overwrite)", timeit(lambda: L.dwconv_fwd(dy.ptr, Cc, w.data_ptr(), None, dx.ptr, Cc, 1, Nn, Hh, Hh, Cc, kk, 1, 0, s)), bytes_=2 * nb)
            emit(out, tag + "dgrad (flip, accumulate)", timeit(lambda: L.dwconv_fwd(dy.ptr, Cc, w.data_ptr(), None, dx.ptr, Cc, 1, Nn, Hh, Hh, Cc, kk, 1, 1, s)), bytes_=3 * nb)
            emit(out, tag + "wgrad", timeit(lambda: L.dwconv_wgrad(x.ptr, Cc, dy.ptr, Cc, dw.data_ptr(), db.data_ptr(), 1, Nn, Hh, Hh, Cc, kk, s)), bytes_=2 * nb)
            del x, y, dy, dx

    # ---- tcgen05 pointwise GEMMs at the Block shapes (N=16) ---------------------------------------------------
    if want("gemm"):
        for tag, M, K, N, act in (("uc4.pwconv1 fwd (+bias+GELU, pre)", 1048576, 128, 512, 3), ("uc4.pwconv2 fwd", 1048576, 512, 64, 0),
                                  ("uc3.pwconv1 fwd (+bias+GELU, pre)", 262144, 256, 1024, 3), ("uc1.pwconv1 fwd (+bias+GELU, pre)", 16384, 1024, 4096, 3),
                                  ("uc1.pwconv2 fwd", 16384, 4096, 512, 0), ("c5.pwconv1 fwd (+bias+GELU, pre)", 4096, 512, 2048, 3),
                                  ("plain GEMM 16384x4096x4096", 16384, 4096, 4096, 0)):
            A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
            Wt = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
            bias = torch.zeros(N, device="cuda")
            C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
            pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if act else None
            fn = lambda: ctx.L.tc_gemm(0, A.data_ptr(), K, Wt.data_ptr(), K, M, N, K, C.data_ptr(), N, bias.data_ptr(),
                                       pre.data_ptr() if act else None, N, None, 0, act, 0, 0, ctx.stream)
            ms = timeit(fn)
            io = (M * K + N * K + M * N * (2 if act else 1)) * 2
            emit(out, "tc_gemm " + tag + " M%d K%d N%d" % (M, K, N), ms, flops=2.0 * M * N * K,
                 hbm_bytes=io, hbm_bound_ms=round(io / PEAK["hbm_gbs"] / 1e6, 4))
            dW = torch.zeros(N, K, device="cuda")
            ms = timeit(lambda: ctx.L.tc_wgrad(C.data_ptr(), N, A.data_ptr(), K, M, N, K, dW.data_ptr(), K, ctx.stream))
            emit(out, "tc_wgrad " + tag.split(" ")[0] + " P%d Co%d Ci%d" % (M, N, K), ms, flops=2.0 * M * N * K)
            del A, Wt, C, pre, dW

    # ---- tcgen05 implicit-GEMM 3x3 convs at the VGG shapes (N=16) ---------------------------------------------
    if want("conv"):
        for Ci, Co, HW_ in ((64, 64, 256), (128, 128, 128), (256, 256, 64), (512, 512, 32)):
            x = ctx.new(16, HW_, HW_, Ci)
            x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g))
            P = params({"w": (Co, Ci, 3, 3), "b": (Co,)})
            ctx.no_grad = True
            fn = lambda: E.conv2d(ctx, x, P["w"], P["b"], 3, 1, 1, act=E.ACT_RELU)
            ms = timeit(fn)
            ctx.no_grad = False
            emit(out, "tc_conv 3x3 %d->%d @%dx%d N=16 (+bias+ReLU)" % (Ci, Co, HW_, HW_), ms,
                 flops=2.0 * 16 * HW_ * HW_ * Ci * Co * 9)
            del x

    # ---- small-channel layers (CUDA-core sc_conv backend): forward, and forward + backward (dgrad + wgrad + bias) ------
    if want("smallch"):
        for Ci, Co, k, st, pd, HW_, N_ in ((3, 64, 3, 1, 1, 256, 16), (64, 3, 3, 1, 1, 256, 16), (6, 32, 4, 2, 1, 256, 16),
                                           (3, 12, 1, 1, 0, 256, 16), (12, 64, 1, 1, 0, 256, 16), (3, 64, 1, 1, 0, 256, 16)):
            x = ctx.new(N_, HW_, HW_, Ci)
            x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g))
            P = params({"w": (Co, Ci, k, k) if k > 1 else (Co, Ci), "b": (Co,)})
            Ho = (HW_ + 2 * pd - k) // st + 1
            ctx.no_grad = True
            ms = timeit(lambda: E.conv2d(ctx, x, P["w"], P["b"], k, st, pd))
            ctx.no_grad = False
            io = (N_ * HW_ * HW_ * x.ld + N_ * Ho * Ho * ((Co + 7) // 8 * 8)) * 2
            emit(out, "sc_conv k%ds%d %d->%d @%dx%d N=%d fwd" % (k, st, Ci, Co, HW_, HW_, N_), ms, bytes_=io)

            def fwd_bwd():
                x.g = None
                y = E.conv2d(ctx, x, P["w"], P["b"], k, st, pd)
                y.grad_out()
                ctx.backward()
            ms2 = timeit(fwd_bwd)
            emit(out, "sc_conv k%ds%d %d->%d fwd+dgrad+wgrad+colsum" % (k, st, Ci, Co), ms2, bytes_=3 * io)
            del x

    # ---- fused ConvNeXt Block MLP (csrc/fused_mlp.cu) at the four generator Blocks that use it, N = 16 -------------------
    if want("mlp"):
        L, st = ctx.L, ctx.stream
        for name, M, cin, nout in (("uc4", 1048576, 128, 64), ("uc3", 262144, 256, 128), ("c2", 262144, 64, 128),
                                   ("c3", 65536, 128, 256)):
            hid = 4 * cin
            T = torch.randn(M, cin, device="cuda", generator=g).bfloat16()
            X = torch.randn(M, cin, device="cuda", generator=g).bfloat16()
            dY = (torch.randn(M, nout, device="cuda", generator=g) * 0.1).bfloat16()
            W1 = (torch.randn(hid, cin, device="cuda", generator=g) / cin ** 0.5).bfloat16()
            W2 = (torch.randn(nout, hid, device="cuda", generator=g) / hid ** 0.5).bfloat16()
            Ws = (torch.randn(nout, cin, device="cuda", generator=g) / cin ** 0.5).bfloat16()
            b1 = torch.randn(hid, device="cuda", generator=g) * 0.1
            b2 = torch.randn(nout, device="cuda", generator=g) * 0.1
            Y = torch.empty(M, nout, device="cuda", dtype=torch.bfloat16)
            dT = torch.empty(M, cin, device="cuda", dtype=torch.bfloat16)
            G = torch.empty(M, hid, device="cuda", dtype=torch.bfloat16)
            A = torch.empty(M, hid, device="cuda", dtype=torch.bfloat16)
            db1 = torch.zeros(hid, device="cuda")
            fl_f = 2.0 * M * (cin * hid + hid * nout + cin * nout)
            fl_b = 2.0 * M * (hid * nout + hid * cin)
            ms = timeit(lambda: L.fused_mlp_fwd(T.data_ptr(), cin, X.data_ptr(), cin, M, cin, nout, W1.data_ptr(), b1.data_ptr(),
                                                W2.data_ptr(), b2.data_ptr(), Ws.data_ptr(), Y.data_ptr(), nout, st))
            emit(out, "fused_mlp_fwd %s M%d %d->%d->%d (+shortcut)" % (name, M, cin, hid, nout), ms, flops=fl_f,
                 hbm_bytes=2 * M * (2 * cin + nout), hbm_bound_ms=round(2 * M * (2 * cin + nout) / PEAK["hbm_gbs"] / 1e6, 4))
            ms = timeit(lambda: L.fused_mlp_bwd(T.data_ptr(), cin, dY.data_ptr(), nout, M, cin, nout, W1.data_ptr(), b1.data_ptr(),
                                                W2.data_ptr(), dT.data_ptr(), cin, G.data_ptr(), A.data_ptr(), db1.data_ptr(), None, st))
            nb = 2 * M * (2 * cin + nout + 2 * hid)
            emit(out, "fused_mlp_bwd %s M%d %d->%d->%d (dT, G, A, db1; hidden recomputed)" % (name, M, cin, hid, nout), ms,
                 flops=fl_b, hbm_bytes=nb, hbm_bound_ms=round(nb / PEAK["hbm_gbs"] / 1e6, 4),
                 note="algorithmic FLOPs = the two input-gradient GEMMs; the recomputed pwconv1 earns no credit")
            del T, X, dY, Y, dT, G, A

    # ---- config #5: generator-only inference, 32 x 512 x 512, bf16 ---------------------------------------------
    if want("infer"):
        from dsgan_b200.models import networks
        networks.KernelNet.precision = "bf16"
        G = networks.MixConvNeXtML().init_normal().cuda()
        xin = torch.rand(32, 3, 512, 512, device="cuda", generator=g) * 2 - 1
        gctx = G.ctx()

        def infer():
            gctx.no_grad = True
            try:
                G(xin)
            finally:
                gctx.no_grad = False
                gctx.clear()
        ms = timeit(infer, iters=5, warm=2, flush=False)
        rec = {"kernel": "config #5: MixConvNeXtML inference 32x3x512x512 bf16", "ms": round(ms, 3), "img_per_s": round(32 / (ms * 1e-3), 1),
               "algorithmic_flops": 2 * 40.77e9 * 4 * 32, "achieved_tflops": round(2 * 40.77e9 * 4 * 32 / (ms * 1e-3) / 1e12, 1)}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
