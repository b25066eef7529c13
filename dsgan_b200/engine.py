"""Host-side execution engine: NHWC device tensors, a reverse-mode tape, and thin wrappers that hand raw
pointers to the C ABI (include/dsgan_b200.h).  PyTorch is used for device memory and streams only — no
torch operator computes anything on the hot path.

Gradient protocol: every backward kernel can either overwrite or accumulate (`accumulate` flag of the ABI).
`Var.grad_out()` returns the gradient buffer plus that flag, so fan-out (a tensor consumed by several ops, e.g.
the encoder skips R1..R4, MixConvNeXtML.py:476-491) costs no extra pass.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import ConvDesc, DwBranch, PackJob, TcConvDesc, TcWgradDesc, lib, require_device

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4


class Var:
    """An NHWC activation (or a channel slice of one) that can carry a gradient."""
    __slots__ = ("t", "ptr", "N", "H", "W", "C", "ld", "es", "g", "parent", "coff", "fused_act", "bias_gptr", "bias_claimed")

    def __init__(self, t, N, H, W, C, ld=None, ptr=None, parent=None, coff=0):
        self.t, self.N, self.H, self.W, self.C = t, N, H, W, C
        self.ld = C if ld is None else ld
        self.es = t.element_size()
        self.ptr = t.data_ptr() if ptr is None else ptr
        self.g = None          # dense gradient tensor [N,H,W,C] (same dtype as t)
        self.parent, self.coff = parent, coff
        self.fused_act = None  # (act, aux Var): this tensor is act(aux) and its grad is stored w.r.t. aux
        # gradient address of the bias that produced this tensor when its ONLY consumer is an InstanceNorm: the norm's
        # backward then delivers the bias gradient (fp32 sum of its dx, a structural zero) and the producer skips its colsum
        self.bias_gptr, self.bias_claimed = None, False

    @property
    def npix(self):
        return self.N * self.H * self.W

    def slice(self, c0, c):
        return Var(self.t, self.N, self.H, self.W, c, ld=self.ld, ptr=self.ptr + c0 * self.es, parent=self, coff=c0)

    # -- gradient access -------------------------------------------------------------------
    def grad_out(self):
        """-> (ptr, ld, accumulate) of the buffer a backward kernel must write this Var's gradient into."""
        if self.parent is not None:
            par = self.parent
            if par.parent is None and par.g is None:
                # a slice is written before its whole parent: start the parent's gradient at zero so that
                # sibling slices (and later whole-tensor writers) can all accumulate
                par.g = torch.zeros((par.N, par.H, par.W, par.ld), dtype=par.t.dtype, device=par.t.device)
            p, ld, _acc = par.grad_out()
            return p + self.coff * self.es, ld, 1
        if self.g is None:
            self.g = torch.empty((self.N, self.H, self.W, self.ld), dtype=self.t.dtype, device=self.t.device)
            return self.g.data_ptr(), self.ld, 0
        return self.g.data_ptr(), self.ld, 1

    def grad_in(self):
        """-> (ptr, ld) of the accumulated gradient, or None if nothing flowed here."""
        if self.parent is not None:
            r = self.parent.grad_in()
            return None if r is None else (r[0] + self.coff * self.es, r[1])
        return None if self.g is None else (self.g.data_ptr(), self.ld)


class Param:
    """fp32 master parameter + fp32 gradient, both views into per-network flat buffers."""
    __slots__ = ("name", "data", "grad", "cache", "bf16_ptr", "owner")

    def __init__(self, name, data, grad, bf16_ptr=0, owner=None):
        self.name, self.data, self.grad, self.cache = name, data, grad, {}
        self.bf16_ptr = bf16_ptr  # address of this tensor inside the network's flat bf16 shadow copy (0 = none)
        self.owner = owner        # ParamTree whose `pack_epoch` says when the fp32 masters may have changed

    @property
    def epoch(self):
        return self.owner.pack_epoch if self.owner is not None else 0

    @property
    def ptr(self):
        return self.data.data_ptr()

    @property
    def gptr(self):
        return self.grad.data_ptr()


FAMILY = {"fused_mlp_fwd": "dense", "fused_mlp_bwd": "dense", "mid_bias_grads": "ca", "conv_fwd": "dense", "conv_wgrad": "dense", "tc_gemm": "dense", "tc_wgrad": "dense", "tc_conv": "dense",
          "tc_conv_wgrad": "dense", "pack_conv_weight": "optim", "pack_conv_weights": "optim",
          "colsum": "reduce", "dwconv_fwd": "dwconv", "dwconv_wgrad": "dwconv", "dwconv_multi_fwd": "dwconv", "dwconv_multi_wgrad": "dwconv",
          "inorm_stats": "norm", "inorm_apply": "norm", "inorm_bwd_stats": "norm", "inorm_bwd_apply": "norm",
          "maxpool_fwd": "pool", "maxpool_bwd": "pool", "multipool_fwd": "pool", "multipool_bwd": "pool", "ca_fwd": "ca", "ca_bwd": "ca", "scale_nc_fwd": "ca",
          "scale_nc_bwd_reduce": "ca", "scale_nc_bwd_apply": "ca",
          "gan_loss": "loss", "l1_loss": "loss", "tv_loss": "loss", "ssim_fwd": "ssim", "ssim_bwd": "ssim",
          "avgpool2_fwd": "ssim", "avgpool2_bwd": "ssim", "msssim_combine": "ssim", "ssim_combine": "ssim",
          "adam_step": "optim", "pack_transpose_bf16": "optim", "pack_bf16": "optim"}


class Profile:
    """CUDA-event timing of every ABI call on the launching stream, grouped into kernel families; the dense
    (conv / GEMM) family also carries its algorithmic FLOPs (2 x true MACs) for the roofline in bench.py."""

    def __init__(self):
        self.items = []
        self.pending_flops = 0.0
        self.label = ""

    def begin(self, name):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fl, self.pending_flops = self.pending_flops, 0.0
        lab, self.label = self.label, ""
        return (name + (" " + lab if lab else ""), fl, e0)

    def detail(self, top=25):
        """[(ms, launches, TFLOP/s, 'abi-name label')] sorted by time."""
        torch.cuda.synchronize()
        agg = {}
        for name, fl, e0, e1 in self.items:
            d = agg.setdefault(name, [0.0, 0, 0.0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
            d[2] += fl
        rows = sorted(((v[0], v[1], (v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0), k) for k, v in agg.items()),
                      reverse=True)
        return rows[:top]

    def end(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.items.append(tok + (e1,))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, fl, e0, e1 in self.items:
            fam = FAMILY.get(name.split(" ")[0], "elementwise")
            d = out.setdefault(fam, {"ms": 0.0, "n": 0, "flops": 0.0, "name": fam})
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
            d["flops"] += fl
        out.setdefault("dense", {"ms": 0.0, "n": 0, "flops": 0.0, "name": "dense"})
        out["dense"]["name"] = "dense conv/GEMM family (conv_fwd + conv_wgrad)"
        return out


class _NoFork:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Fork:
    def __init__(self, ctx, side, main):
        self.ctx, self.side, self.main = ctx, side, main

    def __enter__(self):
        self._cm = torch.cuda.stream(self.side)
        self._cm.__enter__()
        self._prev, self.ctx._side = self.ctx._side, self.side
        return self

    def __exit__(self, *a):
        self.ctx._side = self._prev
        return self._cm.__exit__(*a)


class Ctx:
    """One engine context per device: dtype mode, tape, scratch, kernel wrappers."""

    def __init__(self, device="cuda:0", precision="bf16"):
        require_device()
        self.L = lib()
        self.device = torch.device(device)
        assert precision in ("bf16", "fp32")
        self.precision = precision
        self.dt = BF16 if precision == "bf16" else F32
        self.tdtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.tape = []
        self.param_grads = True   # False while D is frozen in the G step (pix2pix_model.py:214)
        self.no_grad = False
        self.use_tc = True        # bf16 mode: route eligible GEMMs to the tcgen05 kernels
        self.use_streams = True   # run independent branches on a side stream (fork/join)
        self._side, self._side_stream = None, None
        self.pack_epoch = 0       # bumped whenever a network's fp32 masters may have changed (ParamTree.refresh_bf16)

    @property
    def profile(self):
        return self.L.profiler

    @profile.setter
    def profile(self, p):
        self.L.profiler = p

    def _flops(self, geom, transposed_geom=False):
        """Algorithmic FLOPs = 2 x true MACs of a conv described by geom (strided-transposed ops count the MACs
        of the equivalent forward conv: every (small-grid pixel, tap, ci, co) once)."""
        if self.L.profiler is not None:
            N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
            small = min(Hi * Wi, Ho * Wo) if s > 1 else Ho * Wo
            self.L.profiler.pending_flops = 2.0 * N * small * Ci * Co * k * k
            self.L.profiler.label = "k%ds%d %d->%d @%dx%d->%dx%d" % (k, s, Ci, Co, Hi, Wi, Ho, Wo)

    # ---- memory ---------------------------------------------------------------------------
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def new(self, N, H, W, C):
        """NHWC activation.  In bf16 mode the pixel pitch of odd channel counts (3, 6, 12, 1) is padded to a multiple
        of 8 elements (16 bytes) so the tensor can be a TMA source; the pad lanes are never read."""
        ld = C if (self.dt == F32 or C % 8 == 0) else (C + 7) // 8 * 8
        return Var(torch.empty((N, H, W, ld), dtype=self.tdtype, device=self.device), N, H, W, C, ld=ld)

    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def zeros_f32(self, *shape):
        t = self.f32(*shape)
        self.L.memset(t.data_ptr(), 0, t.numel() * 4, self.stream)
        return t

    def zero_(self, t):
        self.L.memset(t.data_ptr(), 0, t.numel() * t.element_size(), self.stream)

    def record(self, fn):
        if not self.no_grad:
            self.tape.append(fn if self._side is None else (fn, self._side))

    # ---- stream-level concurrency -----------------------------------------------------------------------------------
    # Independent sub-graphs (the OriginMLKA branch next to the main U-Net, D(fake) next to D(real)) are made of many
    # 10-40 us kernels that cannot fill 148 SMs on their own.  fork()/join() put such a branch on a side stream; the
    # same markers replayed in reverse give the backward pass the mirrored concurrency.
    def fork(self):
        """-> context manager: everything launched (and recorded) inside runs on a side stream that first waits for
        the work already queued on the current stream."""
        if not self.use_streams:
            return _NoFork()
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(self.device)
        side = self._side_stream
        side.wait_stream(main)
        if not self.no_grad:
            self.tape.append(("fork", side, main, []))
        return _Fork(self, side, main)

    def join(self, fk, keep=()):
        """Make the current stream wait for the forked branch.  `keep`: Vars that cross the two streams; they are held
        until the mirrored join of the backward pass so the caching allocator cannot hand their memory to one stream
        while the other still uses it."""
        if isinstance(fk, _NoFork):
            return
        fk.main.wait_stream(fk.side)
        if not self.no_grad:
            self.tape.append(("join", fk.side, fk.main, list(keep)))

    def backward(self, tape=None):
        """Run (and drop) a tape in reverse; default: the context's current tape."""
        if tape is None:
            tape, self.tape = self.tape, []
        held = []
        while tape:  # popping lets each layer's activations and gradients be freed as soon as it is done
            item = tape.pop()
            if not isinstance(item, tuple):
                item()
            elif item[0] == "join":     # forward join == backward fork: the side branch may start once we got here
                _kind, side, main, keep = item
                side.wait_stream(main)
                held.append(keep)
            elif item[0] == "fork":     # forward fork == backward join
                _kind, side, main, _keep = item
                main.wait_stream(side)
                held.clear()
            else:
                fn, side = item
                with torch.cuda.stream(side):
                    fn()

    def take_tape(self):
        """Detach the recorded tape (e.g. keep G's graph alive across the D step)."""
        tape, self.tape = self.tape, []
        return tape

    def clear(self):
        self.tape = []

    # ---- raw kernel wrappers (pointer level) ---------------------------------------------------
    def _desc(self, N, Hi, Wi, Ci, Ho, Wo, Co, k, stride, pad, transposed, ld_in, ld_out, wst, act=0, dact=0, acc=0,
              ld_aux=0, ld_pre=0):
        d = ConvDesc()
        d.dtype = self.dt
        d.N, d.Hi, d.Wi, d.Ci, d.Ho, d.Wo, d.Co = N, Hi, Wi, Ci, Ho, Wo, Co
        d.kh = d.kw = k
        d.stride, d.pad, d.transposed = stride, pad, int(transposed)
        d.ld_in, d.ld_out, d.ld_aux, d.ld_pre = ld_in, ld_out, ld_aux, ld_pre
        d.w_sco, d.w_sci, d.w_sky, d.w_skx = wst
        d.act, d.dact, d.accumulate = act, dact, acc
        return d

    def _slabs(self, w, O, I, k, wst):
        """bf16 [k*k][O_pad][I_pad] zero-padded K-major slabs of a conv weight for the tcgen05 implicit GEMM, cached per
        pack epoch.  -> (ptr, O_pad, I_pad)"""
        Op, Ip = (O + 31) // 32 * 32, (I + 63) // 64 * 64
        key = ("slabs", O, I, wst)
        hit = w.cache.get(key)
        cur = torch.cuda.current_stream(self.device)
        sid = cur.cuda_stream
        cap = torch.cuda.is_current_stream_capturing()
        # Cross-stream ordering is only needed INSIDE one step (a forked branch sharing weights with the main stream): fork /
        # join already order consecutive steps.  An event recorded outside a graph capture is never waited on from inside one.
        if hit is None or hit[0] != w.epoch:
            buf = hit[1] if hit is not None else torch.empty(k * k * Op * Ip, dtype=torch.bfloat16, device=self.device)
            self.L.pack_conv_weight(w.ptr, buf.data_ptr(), O, I, Op, Ip, k, k, wst[0], wst[1], wst[2], wst[3], 0,
                                    self.stream)
            ev = torch.cuda.Event()
            ev.record(cur)
            w.cache[key] = hit = (w.epoch, buf, ev, sid, cap, (O, I, Op, Ip, k, wst))
        elif hit[3] != sid and not cap and not hit[4]:
            cur.wait_event(hit[2])      # packed on another stream in this step (forked branch sharing the same weights)
        # (while capturing, every slab was re-packed by repack_slabs() on the forking stream before any branch started, so
        #  fork()'s wait_stream already orders it; events of another capture must not be waited on)
        return hit[1].data_ptr(), Op, Ip

    def repack_slabs(self, tree):
        """Re-pack every conv-weight slab a network has used so far in ONE launch (called right after the network's bf16
        shadow was refreshed, on the stream that forks the branches): 48 tiny per-layer launches per step otherwise."""
        entries = [(p, key) for p in tree._plist.values() for key in p.cache if isinstance(key, tuple) and key[0] == "slabs"]
        if not entries:
            return
        sig = tuple((p.name, key) for p, key in entries)
        st = tree.__dict__.get("_slab_jobs")
        if st is None or st[0] != sig:
            jobs = (PackJob * len(entries))()
            nblocks = 0
            for j, (p, key) in zip(jobs, entries):
                O, I, Op, Ip, k, wst = p.cache[key][5]
                j.src, j.dst = p.ptr, p.cache[key][1].data_ptr()
                j.O, j.I, j.O_pad, j.I_pad, j.kh, j.kw = O, I, Op, Ip, k, k
                j.s_o, j.s_i, j.s_ky, j.s_kx = wst
                j.block0 = nblocks
                nblocks += (k * k * Op * Ip + 1023) // 1024
            host = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
            st = (sig, host.to(self.device), len(entries), nblocks)
            tree.__dict__["_slab_jobs"] = st
        cur = torch.cuda.current_stream(self.device)
        self.L.pack_conv_weights(st[1].data_ptr(), st[2], st[3], self.stream)
        ev = torch.cuda.Event()
        ev.record(cur)
        cap = torch.cuda.is_current_stream_capturing()
        for p, key in entries:
            old = p.cache[key]
            p.cache[key] = (p.epoch, old[1], ev, cur.cuda_stream, cap, old[5])

    def _tc_conv(self, geom, xin, w, wst, bias_ptr, out, act, dact, acc, aux, pre, transposed):
        """tcgen05 implicit-GEMM path of conv_raw; returns False when the shape is not eligible."""
        N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
        if self.dt != BF16 or not self.use_tc or k * k > 16 or s not in (1, 2) or not hasattr(w, "cache"):
            return False
        if not self.L.cdll.dsgan_tc_conv_supported(Ci, Co, xin[1], out[1]):
            return False
        if xin[0] % 16:
            return False
        slabs, co_pad, ci_pad = self._slabs(w, Co, Ci, k, wst)
        d = TcConvDesc()
        d.N, d.Hi, d.Wi, d.Ci, d.ld_in = N, Hi, Wi, Ci, xin[1]
        d.Ho, d.Wo, d.Co, d.ld_out = Ho, Wo, Co, out[1]
        d.nslabs, d.co_pad, d.ci_pad = k * k, co_pad, ci_pad
        d.ld_aux, d.ld_pre = (aux[1] if aux else 0), (pre[1] if pre else 0)
        d.act, d.dact, d.accumulate = act, dact, acc
        self._flops(geom)
        launches = []
        if not transposed:      # out[o] = sum_k in[o*s - p + k]
            launches.append((Ho, Wo, s, 1, 0, 0, [(ky - p, kx - p, ky * k + kx) for ky in range(k) for kx in range(k)]))
        elif s == 1:            # out[o] = sum_k in[o + p - k]
            launches.append((Ho, Wo, 1, 1, 0, 0, [(p - ky, p - kx, ky * k + kx) for ky in range(k) for kx in range(k)]))
        else:                   # out[o] = sum_k in[(o + p - k)/2]: one launch per output parity class
            for py in range(2):
                for px in range(2):
                    taps = [((py + p - ky) // 2, (px + p - kx) // 2, ky * k + kx) for ky in range(k) for kx in range(k)
                            if (py + p - ky) % 2 == 0 and (px + p - kx) % 2 == 0]
                    launches.append(((Ho - py + 1) // 2, (Wo - px + 1) // 2, 1, 2, py, px, taps))
        launches = [l for l in launches if l[6] and l[0] > 0 and l[1] > 0]
        # the four output-parity classes of a stride-2 transposed op share one launch when their grids coincide
        groups = [launches] if len({l[:4] for l in launches}) == 1 else [[l] for l in launches]
        for grp in groups:
            Hg, Wg, is_, os_ = grp[0][:4]
            d.Hg, d.Wg, d.in_stride, d.out_stride, d.nclass = Hg, Wg, is_, os_, len(grp)
            for c, (_h, _w, _i, _o, oy0, ox0, taps) in enumerate(grp):
                d.oy0[c], d.ox0[c], d.ntaps[c] = oy0, ox0, len(taps)
                for i, (dy, dx, sl) in enumerate(taps):
                    d.dy[16 * c + i], d.dx[16 * c + i], d.slab[16 * c + i] = dy, dx, sl
            self.L.tc_conv(ctypes.byref(d), xin[0], slabs, bias_ptr, out[0], pre[0] if pre else None,
                           aux[0] if aux else None, self.stream)
        return True

    def conv_raw(self, geom, xin, w, wst, bias_ptr, out, act=0, dact=0, acc=0, aux=None, pre=None,
                 transposed=False):
        """geom = (N,Hi,Wi,Ci,Ho,Wo,Co,k,stride,pad); xin/out/aux/pre = (ptr, ld); w = Param (or raw fp32 pointer)."""
        assert not (acc and dact and dact != ACT_RELU), "accumulate+dact is only exact for the idempotent ReLU mask"
        N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
        if self._tc_conv(geom, xin, w, wst, bias_ptr, out, act, dact, acc, aux, pre, transposed):
            return
        w_ptr = w.ptr if hasattr(w, "ptr") else w
        d = self._desc(N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p, transposed, xin[1], out[1], wst, act, dact, acc,
                       aux[1] if aux else 0, pre[1] if pre else 0)
        self._flops(geom)
        self.L.conv_fwd(ctypes.byref(d), xin[0], w_ptr, bias_ptr, out[0], pre[0] if pre else None,
                        aux[0] if aux else None, self.stream)

    def _tc_wgrad_conv(self, geom, xin, dout, dw_ptr, wst):
        N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
        if self.dt != BF16 or not self.use_tc or k * k > 16 or s not in (1, 2):
            return False
        if xin[0] % 16 or dout[0] % 16 or not self.L.cdll.dsgan_tc_conv_wgrad_supported(Co, Ci, dout[1], xin[1]):
            return False
        d = TcWgradDesc()
        d.N, d.Hg, d.Wg, d.Cg, d.ld_g = N, Ho, Wo, Co, dout[1]
        d.Hx, d.Wx, d.Cx, d.ld_x = Hi, Wi, Ci, xin[1]
        d.x_stride, d.ntaps = s, k * k
        for ky in range(k):
            for kx in range(k):
                t = ky * k + kx
                d.dy[t], d.dx[t], d.tap_off[t] = ky - p, kx - p, ky * wst[2] + kx * wst[3]
        d.s_g, d.s_x = wst[0], wst[1]
        self._flops(geom)
        self.L.tc_conv_wgrad(ctypes.byref(d), dout[0], xin[0], dw_ptr, self.stream)
        return True

    def wgrad_raw(self, geom, xin, dout, dw_ptr, wst):
        N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
        if self._tc_wgrad_conv(geom, xin, dout, dw_ptr, wst):
            return
        d = self._desc(N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p, False, xin[1], dout[1], wst)
        self._flops(geom)
        self.L.conv_wgrad(ctypes.byref(d), xin[0], dout[0], dw_ptr, self.stream)

    # ---- tensor-core (tcgen05) pointwise GEMMs: used in bf16 mode whenever the shape is eligible -------------
    def tc_ok(self, mode, M, N, K, lda, ldb, ldc, *ptrs):
        if self.dt != BF16 or not self.use_tc or any(p % 16 for p in ptrs if p):
            return False
        return bool(self.L.cdll.dsgan_tc_gemm_supported(mode, M, N, K, lda, ldb, ldc))

    def tc_gemm(self, mode, a, w_ptr, ldb, M, N, K, out, bias_ptr=None, pre=None, aux=None, act=0, dact=0, acc=0):
        assert not (acc and dact and dact != ACT_RELU)
        if self.L.profiler is not None:
            self.L.profiler.pending_flops = 2.0 * M * N * K
            self.L.profiler.label = "mode%d M%d N%d K%d" % (mode, M, N, K)
        self.L.tc_gemm(mode, a[0], a[1], w_ptr, ldb, M, N, K, out[0], out[1], bias_ptr, pre[0] if pre else None,
                       pre[1] if pre else 0, aux[0] if aux else None, aux[1] if aux else 0, act, dact, acc, self.stream)

    def tc_wgrad(self, dy, x, P, Co, Ci, dw_ptr):
        if self.L.profiler is not None:
            self.L.profiler.pending_flops = 2.0 * P * Co * Ci
            self.L.profiler.label = "P%d Co%d Ci%d" % (P, Co, Ci)
        self.L.tc_wgrad(dy[0], dy[1], x[0], x[1], P, Co, Ci, dw_ptr, Ci, self.stream)

    def colsum(self, x, npix, C, out_ptr):
        self.L.colsum(x[0], self.dt, x[1], npix, C, out_ptr, self.stream)

    def copy_channels(self, src, dst, npix, C, acc=0):
        self.L.copy_channels(src[0], src[1], dst[0], dst[1], self.dt, npix, C, acc, self.stream)


# ------------------------------------------------------------------------------------------------
# weight-stride helpers (element strides of the fp32 master weight for (out ch, in ch, ky, kx))
# ------------------------------------------------------------------------------------------------

def wst_conv(Co, Ci, k):          # nn.Conv2d / nn.Linear weight [Co, Ci, k, k]
    return (Ci * k * k, k * k, k, 1)


def wst_conv_T(Co, Ci, k):        # same weight seen from the input-gradient side: out ch = Ci, in ch = Co
    return (k * k, Ci * k * k, k, 1)


def wst_convT(Ci, Co, k):         # nn.ConvTranspose2d weight [Ci, Co, k, k], forward: out ch = Co
    return (k * k, Co * k * k, k, 1)


def wst_convT_T(Ci, Co, k):       # ConvTranspose2d input-gradient / weight-gradient view: out ch = Ci, in ch = Co
    return (Co * k * k, k * k, k, 1)


# ------------------------------------------------------------------------------------------------
# Differentiable ops
# ------------------------------------------------------------------------------------------------

def conv2d(ctx: Ctx, x: Var, w: Param, b, k, stride=1, pad=0, act=ACT_NONE, out: Var = None, acc=0,
           need_dx=True, keep_pre=False, bias_via_in=False, bias_grad=True):
    """nn.Conv2d / nn.Linear (k=1) forward with fused bias + activation.  If `act` is set the returned Var is
    marked `fused_act`: consumers deliver its gradient already multiplied by act' (see conv dgrad / maxpool)."""
    Co, Ci = w.data.shape[0], w.data.shape[1]
    assert Ci == x.C, (w.name, Ci, x.C)
    Ho = (x.H + 2 * pad - k) // stride + 1
    Wo = (x.W + 2 * pad - k) // stride + 1
    y = out if out is not None else ctx.new(x.N, Ho, Wo, Co)
    assert (y.H, y.W, y.C) == (Ho, Wo, Co)
    geom = (x.N, x.H, x.W, Ci, Ho, Wo, Co, k, stride, pad)
    pre = None
    if act == ACT_GELU or keep_pre:
        pre = ctx.new(x.N, Ho, Wo, Co)
    M = x.npix
    pointwise = k == 1 and stride == 1 and pad == 0 and w.bf16_ptr
    prep = (pre.ptr, pre.ld) if pre is not None else None
    bptr = b.ptr if b is not None else None
    if pointwise and ctx.tc_ok(0, M, Co, Ci, x.ld, Ci, y.ld, x.ptr, y.ptr, w.bf16_ptr, pre.ptr if pre else 0):
        ctx.tc_gemm(0, (x.ptr, x.ld), w.bf16_ptr, Ci, M, Co, Ci, (y.ptr, y.ld), bptr, prep, None, act, 0, acc)
    else:
        ctx.conv_raw(geom, (x.ptr, x.ld), w, wst_conv(Co, Ci, k), bptr, (y.ptr, y.ld), act=act, acc=acc, pre=prep)
    if act != ACT_NONE:
        y.fused_act = (act, pre if act == ACT_GELU else y)
    train_w = ctx.param_grads
    if bias_via_in and b is not None and train_w and not ctx.no_grad:
        assert act == ACT_NONE and not acc
        y.bias_gptr = b.gptr

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        if train_w:
            if pointwise and ctx.tc_ok(2, M, Ci, Co, gi[1], x.ld, Ci, gi[0], x.ptr):
                ctx.tc_wgrad(gi, (x.ptr, x.ld), M, Co, Ci, w.gptr)
            else:
                ctx.wgrad_raw(geom, (x.ptr, x.ld), gi, w.gptr, wst_conv(Co, Ci, k))
            if b is not None and y.bias_gptr is None and bias_grad:
                ctx.colsum(gi, y.npix, Co, b.gptr)
            assert y.bias_gptr is None or y.bias_claimed, "bias_via_in: no InstanceNorm consumed %s" % w.name
        if need_dx:
            conv2d_dgrad(ctx, x, gi, w, geom, pointwise)
    ctx.record(bwd)
    return y


def conv2d_dgrad(ctx, x: Var, gi, w: Param, geom, pointwise=False):
    """dx (=|+=) conv-transpose of the output gradient; applies x.fused_act' in the epilogue."""
    N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p = geom
    gp, gld, gacc = x.grad_out()
    dact, aux = (0, None)
    if x.fused_act is not None:
        dact, av = x.fused_act
        aux = (av.ptr, av.ld)
    M = N * Hi * Wi
    if pointwise and ctx.tc_ok(1, M, Ci, Co, gi[1], Ci, gld, gi[0], gp, w.bf16_ptr, aux[0] if aux else 0):
        ctx.tc_gemm(1, gi, w.bf16_ptr, Ci, M, Ci, Co, (gp, gld), None, None, aux, 0, dact, gacc)
        return
    g2 = (N, Ho, Wo, Co, Hi, Wi, Ci, k, s, p)
    ctx.conv_raw(g2, gi, w, wst_conv_T(Co, Ci, k), None, (gp, gld), dact=dact, acc=gacc, aux=aux, transposed=True)


def fused_mlp_ok(ctx: Ctx, x: Var, t: Var, y: Var, params):
    """The fused Block-MLP kernels (csrc/fused_mlp.cu) apply in bf16 mode to C_in in {64,128,256}, N_out in {64,128,256}."""
    if ctx.dt != BF16 or not ctx.use_tc or not getattr(ctx, "use_fused_mlp", True):
        return False
    Cin, Nout = x.C, y.C
    if not ctx.L.cdll.dsgan_fused_mlp_supported(Cin, Nout):
        return False
    if any(p.bf16_ptr == 0 or p.bf16_ptr % 16 for p in params):
        return False
    return not (x.ld % 8 or t.ld % 8 or y.ld % 16 or x.ptr % 16 or t.ptr % 16 or y.ptr % 32)


def block_mlp(ctx: Ctx, x: Var, t_fn, w1: Param, b1: Param, w2: Param, b2: Param, ws: Param, out: Var = None,
              need_dx=True):
    """y = shortcut(x) + pwconv2(GELU(pwconv1(t))) with t = t_fn() (the depthwise + norm branch of a ConvNeXt Block,
    MixConvNeXtML.py:230-243) through the fused tcgen05 kernels: the 4C hidden stays on chip in the forward pass and is
    recomputed in the backward pass.  Tape order: the shortcut's backward is recorded BEFORE t_fn's ops, so that in the
    reverse pass the depthwise input-gradient overwrites dL/dx and the shortcut's GEMM epilogue does the fan-in add."""
    Cin, Nout, H4 = x.C, ws.data.shape[0], w1.data.shape[0]
    M = x.npix
    y = out if out is not None else ctx.new(x.N, x.H, x.W, Nout)
    train_w = ctx.param_grads
    geom = (x.N, x.H, x.W, Cin, x.H, x.W, Nout, 1, 1, 0)

    def bwd_shortcut():
        gi = y.grad_in()
        if gi is None:
            return
        if train_w:
            ctx.tc_wgrad(gi, (x.ptr, x.ld), M, Nout, Cin, ws.gptr)
        if need_dx:
            conv2d_dgrad(ctx, x, gi, ws, geom, True)
    ctx.record(bwd_shortcut)
    t = t_fn()
    L = ctx.L
    if L.profiler is not None:
        L.profiler.pending_flops = 2.0 * M * (Cin * H4 + H4 * Nout + Cin * Nout)
        L.profiler.label = "M%d %d->%d->%d" % (M, Cin, H4, Nout)
    L.fused_mlp_fwd(t.ptr, t.ld, x.ptr, x.ld, M, Cin, Nout, w1.bf16_ptr, b1.ptr, w2.bf16_ptr, b2.ptr, ws.bf16_ptr, y.ptr,
                    y.ld, ctx.stream)

    def bwd_mlp():
        gi = y.grad_in()
        if gi is None:
            return
        G = torch.empty((M, H4), dtype=torch.bfloat16, device=ctx.device)
        A = torch.empty((M, H4), dtype=torch.bfloat16, device=ctx.device)
        gp, gld, gacc = t.grad_out()
        assert gacc == 0 and t.fused_act is None
        if L.profiler is not None:   # algorithmic: the two input-gradient GEMMs (the recomputed pwconv1 earns no credit)
            L.profiler.pending_flops = 2.0 * M * (H4 * Nout + H4 * Cin)
            L.profiler.label = "M%d %d->%d->%d" % (M, Cin, H4, Nout)
        L.fused_mlp_bwd(t.ptr, t.ld, gi[0], gi[1], M, Cin, Nout, w1.bf16_ptr, b1.ptr, w2.bf16_ptr, gp, gld, G.data_ptr(),
                        A.data_ptr(), b1.gptr if train_w else None, b2.gptr if train_w else None, ctx.stream)
        if train_w:
            ctx.tc_wgrad((G.data_ptr(), H4), (t.ptr, t.ld), M, H4, Cin, w1.gptr)
            ctx.tc_wgrad(gi, (A.data_ptr(), H4), M, Nout, H4, w2.gptr)
    ctx.record(bwd_mlp)
    return y


def conv_transpose2d(ctx: Ctx, x: Var, w: Param, b: Param, bias_via_in=False):
    """nn.ConvTranspose2d(k=3, s=2, p=1, output_padding=1) forward (MixConvNeXtML.py:53,150)."""
    Ci, Co, k = w.data.shape[0], w.data.shape[1], 3
    assert Ci == x.C
    Ho, Wo = 2 * x.H, 2 * x.W
    y = ctx.new(x.N, Ho, Wo, Co)
    gf = (x.N, x.H, x.W, Ci, Ho, Wo, Co, k, 2, 1)
    ctx.conv_raw(gf, (x.ptr, x.ld), w, wst_convT(Ci, Co, k), b.ptr, (y.ptr, y.ld), transposed=True)
    train_w = ctx.param_grads
    if bias_via_in and train_w and not ctx.no_grad:
        y.bias_gptr = b.gptr

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        # seen as a stride-2 conv from the (big) output gradient to the (small) input
        gb = (x.N, Ho, Wo, Co, x.H, x.W, Ci, k, 2, 1)
        if train_w:
            ctx.wgrad_raw(gb, gi, (x.ptr, x.ld), w.gptr, wst_convT_T(Ci, Co, k))
            if y.bias_gptr is None:
                ctx.colsum(gi, y.npix, Co, b.gptr)
            assert y.bias_gptr is None or y.bias_claimed, "bias_via_in: no InstanceNorm consumed %s" % w.name
        gp, gld, gacc = x.grad_out()
        assert x.fused_act is None
        ctx.conv_raw(gb, gi, w, wst_convT_T(Ci, Co, k), None, (gp, gld), acc=gacc)
    ctx.record(bwd)
    return y


def _label(ctx, x, extra=""):
    if ctx.L.profiler is not None:
        ctx.L.profiler.label = "C%d @%dx%d %s" % (x.C, x.H, x.W, extra)


def dwconv(ctx: Ctx, x: Var, w: Param, b: Param, k, out: Var = None, need_dx=True, bias_via_in=False, bias_grad=True):
    """Depthwise k x k, stride 1, pad k//2 (MixConvNeXtML.py:94-97,220)."""
    y = out if out is not None else ctx.new(x.N, x.H, x.W, x.C)
    L, s = ctx.L, ctx.stream
    _label(ctx, x, "k%d" % k)
    L.dwconv_fwd(x.ptr, x.ld, w.ptr, b.ptr, y.ptr, y.ld, ctx.dt, x.N, x.H, x.W, x.C, k, 0, 0, s)
    train_w = ctx.param_grads
    if bias_via_in and train_w and not ctx.no_grad:
        y.bias_gptr = b.gptr

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        if train_w:
            _label(ctx, x, "k%d" % k)
            assert y.bias_gptr is None or y.bias_claimed, "bias_via_in: no InstanceNorm consumed %s" % w.name
            L.dwconv_wgrad(x.ptr, x.ld, gi[0], gi[1], w.gptr, b.gptr if (bias_grad and y.bias_gptr is None) else None,
                           ctx.dt, x.N, x.H, x.W, x.C, k, ctx.stream)
        if not need_dx:
            return
        gp, gld, gacc = x.grad_out()
        assert x.fused_act is None
        _label(ctx, x, "k%d dgrad" % k)
        L.dwconv_fwd(gi[0], gi[1], w.ptr, None, gp, gld, ctx.dt, x.N, x.H, x.W, x.C, k, 1, gacc, ctx.stream)
    ctx.record(bwd)
    return y


def dwconv_multi(ctx: Ctx, x: Var, branches, out: Var = None, bias_grad=True):
    """Depthwise convolutions of different kernel sizes on equal channel slices of x, one launch per pass
    (MidMLKA.X3/X5/X7/X9, MixConvNeXtML.py:94-97,110).  branches: [(weight, bias, k), ...]; slice i = channels [i*q, (i+1)*q)."""
    nb = len(branches)
    q = x.C // nb
    assert q * nb == x.C and nb <= 4
    y = out if out is not None else ctx.new(x.N, x.H, x.W, x.C)
    L = ctx.L

    def table(with_grads, flip_bias):
        arr = (DwBranch * nb)()
        for i, (w, b, k) in enumerate(branches):
            arr[i].w, arr[i].bias = w.ptr, (None if flip_bias else b.ptr)
            arr[i].dw = w.gptr if with_grads else None
            arr[i].db = b.gptr if (with_grads and bias_grad) else None
            arr[i].k, arr[i].c0, arr[i].c = k, i * q, q
        return arr
    _label(ctx, x, "k" + "".join(str(k) for _, _, k in branches))
    L.dwconv_multi_fwd(x.ptr, x.ld, y.ptr, y.ld, ctx.dt, x.N, x.H, x.W, table(False, False), nb, 0, 0, ctx.stream)
    train_w = ctx.param_grads

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        if train_w:
            _label(ctx, x, "multi")
            L.dwconv_multi_wgrad(x.ptr, x.ld, gi[0], gi[1], ctx.dt, x.N, x.H, x.W, table(True, False), nb, ctx.stream)
        gp, gld, gacc = x.grad_out()
        assert x.fused_act is None
        _label(ctx, x, "multi dgrad")
        L.dwconv_multi_fwd(gi[0], gi[1], gp, gld, ctx.dt, x.N, x.H, x.W, table(False, True), nb, 1, gacc, ctx.stream)
    ctx.record(bwd)
    return y


def inorm(ctx: Ctx, x: Var, act=ACT_NONE, res: Var = None, out: Var = None, dsum_nc=None):
    """y = act(InstanceNorm(x) + res); `out` may be a channel slice of a concat buffer.  If x carries `bias_gptr` (set by
    its producer's bias_via_in) the backward pass adds the fp32 sum of dx into that bias gradient.  `dsum_nc`: optional
    callable -> fp32 [N,C] tensor that receives the same sums per plane (MidMLKA bias gradients)."""
    dbias = x.bias_gptr
    if dbias is not None:
        x.bias_claimed = True
    y = out if out is not None else ctx.new(x.N, x.H, x.W, x.C)
    L = ctx.L
    HW = x.H * x.W
    stats = ctx.f32(x.N, x.C, 3)
    _label(ctx, x)
    L.inorm_stats(x.ptr, x.ld, ctx.dt, x.N, HW, x.C, stats.data_ptr(), ctx.stream)
    rp, rld = (res.ptr, res.ld) if res is not None else (None, 0)
    _label(ctx, x)
    L.inorm_apply(x.ptr, x.ld, stats.data_ptr(), rp, rld, y.ptr, y.ld, ctx.dt, x.N, HW, x.C, act, ctx.stream)

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        bst = ctx.f32(x.N, x.C, 2)
        L.inorm_bwd_stats(x.ptr, x.ld, stats.data_ptr(), rp, rld, gi[0], gi[1], ctx.dt, x.N, HW, x.C, act,
                          bst.data_ptr(), ctx.stream)
        gp, gld, gacc = x.grad_out()
        assert x.fused_act is None
        if res is not None:
            qp, qld, qacc = res.grad_out()
            assert res.fused_act is None
        else:
            qp, qld, qacc = None, 0, 0
        assert dbias is None or gacc == 0, "bias_via_in needs the norm to be the only consumer"
        sums = dsum_nc() if dsum_nc is not None else None
        L.inorm_bwd_apply(x.ptr, x.ld, stats.data_ptr(), rp, rld, gi[0], gi[1], bst.data_ptr(), gp, gld, gacc,
                          qp, qld, qacc, ctx.dt, x.N, HW, x.C, act, dbias, sums.data_ptr() if sums is not None else None,
                          ctx.stream)
    ctx.record(bwd)
    return y


def maxpool(ctx: Ctx, x: Var, k):
    y = ctx.new(x.N, x.H // k, x.W // k, x.C)
    ctx.L.maxpool_fwd(x.ptr, x.ld, y.ptr, y.ld, ctx.dt, x.N, x.H, x.W, x.C, k, ctx.stream)

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        gp, gld, gacc = x.grad_out()
        relu = 0
        if x.fused_act is not None:
            assert x.fused_act[0] == ACT_RELU, "maxpool backward only fuses the ReLU mask"
            relu = 1
        ctx.L.maxpool_bwd(x.ptr, x.ld, gi[0], gi[1], gp, gld, ctx.dt, x.N, x.H, x.W, x.C, k, gacc, relu, ctx.stream)
    ctx.record(bwd)
    return y


def multi_maxpool(ctx: Ctx, x: Var, nlev):
    """[MaxPool2d(2)(x), MaxPool2d(4)(x), ... MaxPool2d(2**nlev)(x)] in one pass over x, and one combined backward pass
    (the multi-scale down-skips of an encoder stage, MixConvNeXtML.py:328-426, share their input with the stage's own
    downSample).  Falls back to separate pooling kernels where the fused kernel does not apply (fp32 mode)."""
    if nlev < 2 or not ctx.L.cdll.dsgan_multipool_supported(ctx.dt, x.H, x.W, x.C, nlev, x.ld) or x.ptr % 16:
        return [maxpool(ctx, x, 2 << l) for l in range(nlev)]
    ys = [ctx.new(x.N, x.H >> (l + 1), x.W >> (l + 1), x.C) for l in range(nlev)]
    yp = [y.ptr for y in ys] + [None] * (4 - nlev)
    _label(ctx, x, "k2..%d" % (1 << nlev))
    ctx.L.multipool_fwd(x.ptr, x.ld, yp[0], yp[1], yp[2], yp[3], ctx.dt, x.N, x.H, x.W, x.C, nlev, ctx.stream)

    def bwd():
        gis = [y.grad_in() for y in ys]
        if all(g is None for g in gis):
            return
        assert all(g is None or g[1] == x.C for g in gis) and x.fused_act is None
        gp, gld, gacc = x.grad_out()
        g = [gi[0] if gi is not None else None for gi in gis] + [None] * (4 - nlev)
        _label(ctx, x, "k2..%d" % (1 << nlev))
        ctx.L.multipool_bwd(x.ptr, x.ld, g[0], g[1], g[2], g[3], gp, gld, gacc, ctx.dt, x.N, x.H, x.W, x.C, nlev, ctx.stream)
    ctx.record(bwd)
    return ys


def add_n(ctx: Ctx, xs, share_grad=False):
    """Sum of 2..5 same-shaped tensors (multi-scale skip sums, MixConvNeXtML.py:482-491).  `share_grad`: every addend is
    consumed by this sum ONLY, so its gradient IS the sum's gradient: the addends alias the one gradient tensor instead of
    receiving a copy each (22 copies per step).  Safe because backward kernels only read the gradient of their output; an
    addend with a second consumer would accumulate into the shared buffer -- the caller vouches there is none."""
    x0 = xs[0]
    assert all(v.ld == v.C for v in xs)
    y = ctx.new(x0.N, x0.H, x0.W, x0.C)
    ptrs = [v.ptr for v in xs] + [None] * (5 - len(xs))
    ctx.L.add_n(y.ptr, ctx.dt, x0.npix * x0.C, *ptrs, ctx.stream)

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        for v in xs:
            assert v.fused_act is None
            if share_grad and v.parent is None and v.g is None and y.g is not None:
                v.g = y.g
                continue
            gp, gld, gacc = v.grad_out()
            ctx.copy_channels(gi, (gp, gld), v.npix, v.C, gacc)
    ctx.record(bwd)
    return y


def concat_into(ctx: Ctx, cat: Var, coff, src: Var):
    """cat[..., coff:coff+src.C] = src, with the matching slice-gradient routed back to src."""
    dst = cat.slice(coff, src.C)
    ctx.copy_channels((src.ptr, src.ld), (dst.ptr, dst.ld), src.npix, src.C, 0)

    def bwd():
        gi = dst.grad_in()
        if gi is None:
            return
        gp, gld, gacc = src.grad_out()
        assert src.fused_act is None
        ctx.copy_channels(gi, (gp, gld), src.npix, src.C, gacc)
    ctx.record(bwd)


def ca_scale(ctx: Ctx, x: Var, fc1: Param, slope: Param, fc2: Param, mid_bias=None):
    """y = x * CA(x)  (MixConvNeXtML.py:17-22,112).  `mid_bias` = (sums holder dict, conv weight, conv bias, [4 depthwise
    biases]): the MidMLKA bias gradients are formed here from fp32 plane statistics (dsgan_mid_bias_grads)."""
    N, C, HW = x.N, x.C, x.H * x.W
    assert x.ld == x.C
    L = ctx.L
    avg, mx, s = ctx.f32(N, C), ctx.f32(N, C), ctx.f32(N, C)
    am = torch.empty((N, C), dtype=torch.int32, device=ctx.device)
    ws = torch.empty((N, C), dtype=torch.int64, device=ctx.device)
    L.ca_fwd(x.ptr, ctx.dt, N, HW, C, fc1.ptr, slope.ptr, fc2.ptr, avg.data_ptr(), mx.data_ptr(), am.data_ptr(),
             s.data_ptr(), ws.data_ptr(), ctx.stream)
    y = ctx.new(N, x.H, x.W, C)
    L.scale_nc_fwd(x.ptr, s.data_ptr(), y.ptr, ctx.dt, N, HW, C, ctx.stream)
    train_w = ctx.param_grads

    def bwd():
        gi = y.grad_in()
        if gi is None:
            return
        assert gi[1] == C
        ds, davg, dmax = ctx.f32(N, C), ctx.f32(N, C), ctx.f32(N, C)
        L.scale_nc_bwd_reduce(x.ptr, gi[0], ds.data_ptr(), ctx.dt, N, HW, C, ctx.stream)
        if train_w:
            d1, dsl, d2 = fc1.gptr, slope.gptr, fc2.gptr
        else:
            scratch = ctx.zeros_f32(fc1.data.numel() + 1 + fc2.data.numel())
            d1 = scratch.data_ptr()
            dsl = d1 + 4 * fc1.data.numel()
            d2 = dsl + 4
        L.ca_bwd(ds.data_ptr(), s.data_ptr(), avg.data_ptr(), mx.data_ptr(), N, C, fc1.ptr, slope.ptr, fc2.ptr,
                 d1, dsl, d2, davg.data_ptr(), dmax.data_ptr(), ctx.stream)
        if mid_bias is not None and train_w:
            holder, wc, bc, bq = mid_bias
            L.mid_bias_grads(holder["S"].data_ptr(), s.data_ptr(), davg.data_ptr(), dmax.data_ptr(), N, C, wc.ptr,
                             bc.gptr, bq[0].gptr, bq[1].gptr, bq[2].gptr, bq[3].gptr, ctx.stream)
        gp, gld, gacc = x.grad_out()
        assert gld == C and x.fused_act is None
        L.scale_nc_bwd_apply(gi[0], s.data_ptr(), davg.data_ptr(), dmax.data_ptr(), am.data_ptr(), gp, ctx.dt, N, HW,
                             C, gacc, ctx.stream)
    ctx.record(bwd)
    return y


# ------------------------------------------------------------------------------------------------
# image <-> NHWC boundary
# ------------------------------------------------------------------------------------------------

def image_to_nhwc(ctx: Ctx, img: torch.Tensor, out: Var = None, scale=1.0, shift=0.0):
    """NCHW fp32 image -> NHWC activation (optionally a channel slice, e.g. for cat(real_A, fake_B))."""
    N, C, H, W = img.shape
    assert img.dtype == torch.float32 and img.is_contiguous() and img.is_cuda
    y = out if out is not None else ctx.new(N, H, W, C)
    ctx.L.nchw_to_nhwc(img.data_ptr(), y.ptr, ctx.dt, N, C, H, W, y.ld, scale, shift, ctx.stream)
    return y


def nhwc_grad_to_image(ctx: Ctx, v: Var, dimg: torch.Tensor, alpha=1.0, acc=1):
    """dimg (NCHW fp32) (=|+=) alpha * v.grad."""
    gi = v.grad_in()
    if gi is None:
        return
    ctx.L.nhwc_to_nchw(gi[0], ctx.dt, gi[1], dimg.data_ptr(), v.N, v.C, v.H, v.W, alpha, acc, ctx.stream)
