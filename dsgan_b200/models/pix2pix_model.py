"""Pix2PixModel with the reference's surface (DSGAN/models/pix2pix_model.py:61-310): initialize, set_input, forward,
backward_D, backward_G, optimize_parameters, get_img_*.  Every tensor op below this file is a hand-written
sm_100a kernel reached through the C ABI; orchestration stays in Python like the reference.

Data parallelism (replaces nn.DataParallel, networks.py:77): one process per GPU; each rank runs this step on its
batch shard and gradients are summed with NCCL over the networks' flat gradient buffers, then averaged inside
the fused Adam (grad_scale = 1/world).  The TV term is a batch SUM in the reference (pix2pix_model.py:189-191),
so its gradient is pre-scaled by `world` to keep R-rank training equal to single-GPU big-batch math.
"""
import torch

from .. import losses, nets, parallel
from ..engine import image_to_nhwc, nhwc_grad_to_image
from ..optim import FlatAdam
from ..util.image_pool import ImagePool
from . import networks
from .base_model import BaseModel
from .vgg import Vgg16

SLOT = {"G_GAN": 0, "G_L1": 1, "vgg": 2, "tv": 3, "ssim": 4, "D_fake": 5, "D_real": 6}


class Pix2PixModel(BaseModel):
    def name(self):
        return "Pix2PixModel"

    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        if is_train:
            parser.add_argument("--lambda_L1", type=float, default=100.0, help="weight for L1 loss (unused, Q11)")
        return parser

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        self.precision = getattr(opt, "precision", "bf16")
        self.loss_names = ["G_GAN", "G_L1", "D_real", "D_fake"]
        self.visual_names = ["real_A", "fake_B", "real_B"]
        self.model_names = ["G", "D"] if self.isTrain else ["G"]
        self.use_gan, self.use_condition = opt.use_GAN, opt.use_condition
        self.w_vgg, self.w_tv, self.w_gan, self.w_ss = opt.w_vgg, opt.w_tv, opt.w_gan, opt.w_ss
        # per-network activation precision (extension): --precision sets all three, --precision_{G,D,vgg} override one.
        # The networks only meet at NCHW fp32 images (fake_B, dL/dfake_B), so each can run in its own engine context.
        prec = {k: (getattr(opt, "precision_" + k, "") or self.precision) for k in ("G", "D", "vgg")}
        self.precisions = prec
        networks.KernelNet.precision = self.precision
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.which_model_netG, opt.norm,
                                      not opt.no_dropout, opt.init_type, self.gpu_ids)
        self.netG.precision = prec["G"]
        self.ctx = self.ctxG = networks.get_ctx(self.device, prec["G"])
        self.ctxD = networks.get_ctx(self.device, prec["D"])
        self.ctxV = networks.get_ctx(self.device, prec["vgg"])
        self.world = parallel.world_size()
        self._in_buf, self._gs = {}, None
        self._stage = None     # prefetch_input(): staging buffers + events of the next batch
        self.use_graph = bool(int(getattr(opt, "cuda_graph", 1))) and self.isTrain
        if self.isTrain:
            use_sigmoid = opt.no_lsgan
            d_in = opt.input_nc + opt.output_nc if self.use_condition == 1 else opt.input_nc
            self.netD = networks.define_D(d_in, opt.ndf, opt.which_model_netD, opt.n_layers_D, opt.norm, use_sigmoid,
                                          opt.init_type, self.gpu_ids)
            self.fake_AB_pool = ImagePool(opt.pool_size)
            self.use_lsgan = opt.no_lsgan  # the reference's inverted flag (pix2pix_model.py:98,112-114, Q6)
            self.criterionGAN = networks.GANLoss(use_lsgan=opt.no_lsgan).to(self.device)
            self.netD.precision = prec["D"]
            self.vgg = Vgg16().to(self.device)
            self.vgg.precision = prec["vgg"]
            vw = getattr(opt, "vgg_weights", "")
            if vw:
                self.vgg.load_torchvision(vw)
            elif float(self.w_vgg) != 0:
                import warnings
                warnings.warn("Vgg16 is RANDOM-initialised: pass --vgg_weights <torchvision vgg16 state_dict> to train with "
                              "the reference's ImageNet perceptual loss (models/vgg.py:8 downloads it; no network here)")
            self.optimizer_G = FlatAdam(self.netG, lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizer_D = FlatAdam(self.netD, lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizers = [self.optimizer_G, self.optimizer_D]
            self._loss = torch.zeros(8, dtype=torch.float32, device=self.device)
            init = torch.zeros(8, dtype=torch.float32)
            init[SLOT["ssim"]] = 1.0  # loss_ssim = 1 - ssim
            self._loss_init = init.to(self.device)
            if self.world > 1:  # identical replicas: rank 0's initial weights everywhere
                for net in (self.netG, self.netD, self.vgg):
                    parallel.broadcast_params(net.flat_buffers()[0], 0)

    # ---- reference API ---------------------------------------------------------------------
    def _to_input_buffer(self, name, src):
        """Training inputs land in persistent device buffers (same address every step: a captured graph can read them, and
        the steady state allocates nothing); a shape change re-allocates."""
        buf = self._in_buf.get(name)
        if buf is None or buf.shape != src.shape:
            buf = self._in_buf[name] = torch.empty(src.shape, dtype=torch.float32, device=self.device)
        buf.copy_(src, non_blocking=True)
        return buf

    def prefetch_input(self, input):
        """Optional: start the host-to-device copy of the NEXT batch on a copy stream while the current step is still running
        (what a DataLoader with pinned memory + non_blocking copies gives the reference).  A following set_input() of the same
        tensors only waits for that copy and moves it device-to-device into the persistent input buffers."""
        if not self.isTrain:
            return
        AtoB = self.opt.which_direction == "AtoB"
        a, b = input["A" if AtoB else "B"], input["B" if AtoB else "A"]
        st = self._stage
        if st is None or st["A"].shape != a.shape or st["B"].shape != b.shape:
            st = self._stage = {"A": torch.empty(a.shape, dtype=torch.float32, device=self.device),
                                "B": torch.empty(b.shape, dtype=torch.float32, device=self.device),
                                "stream": torch.cuda.Stream(self.device), "ready": torch.cuda.Event(),
                                "consumed": None, "key": None}
        cs = st["stream"]
        if st["consumed"] is not None:
            cs.wait_event(st["consumed"])          # the previous staged batch has been moved out of the staging buffers
        with torch.cuda.stream(cs):
            st["A"].copy_(a, non_blocking=True)
            st["B"].copy_(b, non_blocking=True)
            st["ready"].record(cs)
        st["key"] = (a.data_ptr(), b.data_ptr(), tuple(a.shape))

    def _take_staged(self, a, b):
        st = self._stage
        if st is None or st["key"] != (a.data_ptr(), b.data_ptr(), tuple(a.shape)):
            return False
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(st["ready"])
        for name in ("A", "B"):
            buf = self._in_buf.get(name)
            if buf is None or buf.shape != st[name].shape:
                buf = self._in_buf[name] = torch.empty(st[name].shape, dtype=torch.float32, device=self.device)
            buf.copy_(st[name], non_blocking=True)
        if st["consumed"] is None:
            st["consumed"] = torch.cuda.Event()
        st["consumed"].record(cur)
        st["key"] = None
        self.real_A, self.real_B = self._in_buf["A"], self._in_buf["B"]
        return True

    def set_input(self, input):
        AtoB = self.opt.which_direction == "AtoB"
        a, b = input["A" if AtoB else "B"], input["B" if AtoB else "A"]
        if self.isTrain:
            if not self._take_staged(a, b):
                self.real_A, self.real_B = self._to_input_buffer("A", a), self._to_input_buffer("B", b)
        else:
            self.real_A = a.to(self.device, non_blocking=True).float().contiguous()
            self.real_B = b.to(self.device, non_blocking=True).float().contiguous()
        self.image_paths = input["A_paths" if AtoB else "B_paths"]

    def forward(self):
        self.ctxG.clear()
        self.fake_B = self.netG(self.real_A)
        self._g_out = self.netG.last_output
        self._g_tape = self.ctxG.take_tape()

    def _slot(self, name):
        return self._loss.data_ptr() + 4 * SLOT[name]

    def _pair(self, ctx, a, b):
        """cat((a, b), 1) as one NHWC tensor (pix2pix_model.py:145,153,168)."""
        N, C, H, W = a.shape
        v = ctx.new(N, H, W, C + b.shape[1])
        image_to_nhwc(ctx, a, out=v.slice(0, C))
        image_to_nhwc(ctx, b, out=v.slice(C, b.shape[1]))
        return v

    def _gan_kw(self):
        return dict(use_lsgan=self.use_lsgan, sigmoid_d=self.use_lsgan)

    def backward_D(self):
        ctx = self.ctxD
        ctx.param_grads = True
        if self.use_condition == 1:
            xf = image_to_nhwc(ctx, self._pooled_fake().contiguous())
            xr = self._pair(ctx, self.real_A, self.real_B)
        else:
            xf, xr = image_to_nhwc(ctx, self.fake_B), image_to_nhwc(ctx, self.real_B)
        P = self.netD.params()  # (refreshes the bf16 operands once, before the two branches fork)
        fk = ctx.fork()         # D(fake) and D(real) are independent: run them on two streams
        with fk:
            pred_fake = nets.discriminator_forward(ctx, P, xf, need_dx=False)
            losses.gan_loss(ctx, pred_fake, False, self._slot("D_fake"), 1.0, 0.5, **self._gan_kw())
        pred_real = nets.discriminator_forward(ctx, P, xr, need_dx=False)
        losses.gan_loss(ctx, pred_real, True, self._slot("D_real"), 1.0, 0.5, **self._gan_kw())
        ctx.join(fk, keep=(xf, xr, pred_fake))
        ctx.backward()  # loss_D = 0.5*(fake+real)

    def backward_G(self):
        cG, cD, cV = self.ctxG, self.ctxD, self.ctxV
        fake, real = self.fake_B, self.real_B
        dfake = torch.empty_like(fake)
        cG.zero_(dfake)
        cD.param_grads = cV.param_grads = False  # D and VGG are frozen here (set_requires_grad(netD, False), vgg.py:27-28)
        # VGG(real_B) needs nothing from this step: run it (no tape) on the side stream, next to the discriminator's small
        # kernels; joined right before the feature losses.
        side = None
        if cV.use_streams and fake.is_cuda:
            main = torch.cuda.current_stream(cV.device)
            if cV._side_stream is None:
                cV._side_stream = torch.cuda.Stream(cV.device)
            side = cV._side_stream
            Pv = self.vgg.params()   # on the MAIN stream: lazy flattening / operand refresh must be ordered before both passes
            side.wait_stream(main)
            with torch.cuda.stream(side):
                cV.no_grad = True
                fr = nets.vgg_forward(cV, Pv, image_to_nhwc(cV, real), need_dx=False)
                cV.no_grad = False
        if self.use_gan == 1:
            x = self._pair(cD, self.real_A, fake) if self.use_condition == 1 else image_to_nhwc(cD, fake)
            pred = self.netD.forward_var(x, need_dx=True)
            losses.gan_loss(cD, pred, True, self._slot("G_GAN"), 1.0, float(self.w_gan), **self._gan_kw())
            cD.backward()
            gv = x.slice(self.real_A.shape[1], fake.shape[1]) if self.use_condition == 1 else x
            nhwc_grad_to_image(cD, gv, dfake)
        losses.l1_images(cG, fake, real, self._slot("G_L1"), 1.0, dfake)
        # VGG perceptual loss on raw [-1,1] images (pix2pix_model.py:180-186, Q14)
        if side is None:
            cV.no_grad = True
            fr = self.vgg.forward_var(image_to_nhwc(cV, real), need_dx=False)
            cV.no_grad = False
        xv = image_to_nhwc(cV, fake)
        ff = self.vgg.forward_var(xv, need_dx=True)
        if side is not None:
            torch.cuda.current_stream(cV.device).wait_stream(side)
        for f, r in zip(ff, fr):
            losses.l1_features(cV, f, r, self._slot("vgg"), float(self.w_vgg))
        cV.backward()
        nhwc_grad_to_image(cV, xv, dfake)
        losses.tv_loss(cG, fake, self._slot("tv"), float(self.w_tv) * parallel.tv_grad_scale(self.world), dfake)
        losses.ssim_training_loss(cG, real, fake, self._slot("ssim"), float(self.w_ss), dfake)
        # dL/dfake_B -> generator backward
        cD.param_grads = cV.param_grads = cG.param_grads = True
        gp, gld, _acc = self._g_out.grad_out()
        cG.L.nchw_to_nhwc(dfake.data_ptr(), gp, cG.dt, fake.shape[0], fake.shape[1], fake.shape[2], fake.shape[3],
                          gld, 1.0, 0.0, cG.stream)
        cG.backward(self._g_tape)
        self._g_tape = self._g_out = None

    def _allreduce(self, net):
        parallel.allreduce_grads(net.flat_buffers()[1])

    def _pooled_fake(self):
        """ImagePool.query(cat(real_A, fake_B)) (pix2pix_model.py:145).  Host logic with python `random`: while a graph is
        being captured / replayed it runs eagerly BETWEEN the graph segments and hands its result over in a static buffer."""
        if self._gs is not None and self._gs.get("pool_in") is not None:
            return self._gs["pool_in"]
        return self.fake_AB_pool.query(torch.cat((self.real_A, self.fake_B), 1))

    # The step as a list of segments; "eager" ones (image pool, NCCL all-reduce) are never captured.
    def _segments(self):
        gscale = parallel.adam_grad_scale(self.world)
        graphed = self._gs is not None
        adam = (lambda o: o.record_step(gscale)) if graphed else (lambda o: o.step(gscale))

        def seg_forward():
            self._loss.copy_(self._loss_init)
            self.forward()

        def seg_pool():   # eager: feed the image pool's answer to the captured D step
            gs = self._gs
            if self.use_gan == 1 and self.use_condition == 1:
                out = self.fake_AB_pool.query(torch.cat((self.real_A, self.fake_B), 1))
                gs["pool_in"].copy_(out)

        def seg_d():
            self.set_requires_grad(self.netD, True)
            self.optimizer_D.zero_grad()
            self.backward_D()

        def seg_d_adam_and_g():
            if self.use_gan == 1:
                adam(self.optimizer_D)
            self.set_requires_grad(self.netD, False)
            self.optimizer_G.zero_grad()
            self.backward_G()

        def seg_g_adam():
            adam(self.optimizer_G)

        segs = [("graph", seg_forward)]
        if graphed:
            segs.append(("eager", seg_pool))
        if self.use_gan == 1:
            segs.append(("graph", seg_d))
            if self.world > 1:
                segs.append(("eager", lambda: self._allreduce(self.netD)))
        segs.append(("graph", seg_d_adam_and_g))
        if self.world > 1:
            segs.append(("eager", lambda: self._allreduce(self.netG)))
        segs.append(("graph", seg_g_adam))
        return segs

    def _eager_step(self):
        gs, self._gs = self._gs, None      # plain path: pool queried inline, Adam with host-side scalars
        try:
            for _kind, fn in self._segments():
                fn()
        finally:
            self._gs = gs

    def _flat_ptrs(self):
        return tuple(t.data_ptr() for net in (self.netG, self.netD, self.vgg) for t in net.flat_buffers()[:2])

    def _capture(self, gs):
        """Capture the graph segments of one step (nothing executes).  All segments share one memory pool: G's forward
        activations (segment 1) are consumed and freed by the backward pass captured in a later segment."""
        torch.cuda.synchronize()
        if self.use_gan == 1 and self.use_condition == 1:
            n, ca, h, w = self.real_A.shape
            gs["pool_in"] = torch.empty((n, ca + self.real_B.shape[1], h, w), dtype=torch.float32, device=self.device)
        else:
            gs["pool_in"] = None
        for o in (self.optimizer_G, self.optimizer_D):
            o._state()
        self._gs = gs
        pool = torch.cuda.graph_pool_handle()
        n0 = self.ctx.L.cdll.dsgan_launch_count()
        plan, run = [], []
        for kind, fn in self._segments():
            if kind == "graph" and run and run[-1][0] == "graph":
                run[-1][1].append(fn)          # merge adjacent graph segments
            else:
                run.append((kind, [fn]))
        for kind, fns in run:
            if kind == "eager":
                plan.append(("eager", fns))
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for fn in fns:
                    fn()
            plan.append(("graph", g))
        gs["plan"], gs["ptrs"] = plan, self._flat_ptrs()
        gs["fake_B"] = self.fake_B      # the tensor the captured forward writes (lives in the graphs' memory pool)
        gs["kernels_per_replay"] = int(self.ctx.L.cdll.dsgan_launch_count() - n0)   # launches recorded into the graphs
        torch.cuda.synchronize()

    def _replay(self, gs):
        # forward()/test() outside the graphs (get_img_gen after every step in the reference's train.py:112) re-bind
        # self.fake_B to a fresh eager tensor; the eager image-pool segment and the visuals must see the captured one.
        self.fake_B = gs["fake_B"]
        if self.use_gan == 1:
            self.optimizer_D.advance()
        self.optimizer_G.advance()
        for kind, item in gs["plan"]:
            if kind == "graph":
                item.replay()
            else:
                for fn in item:
                    fn()

    def optimize_parameters(self):
        if not (self.use_graph and self.ctx.profile is None):
            return self._eager_step()
        key = (tuple(self.real_A.shape), tuple(self.real_B.shape), self.real_A.data_ptr(), self.real_B.data_ptr(),
               self.world, self.ctxG.use_streams, self.ctxD.use_streams, self.ctxV.use_streams)
        gs = self._gs
        if gs is None or gs["key"] != key or (gs.get("plan") is not None and gs["ptrs"] != self._flat_ptrs()):
            self._gs = gs = {"key": key, "warm": 0, "plan": None, "pool_in": None}
        if gs["plan"] is None:
            if gs["warm"] < 2:      # two eager steps first: lazy one-time work (function attributes, frozen-weight packing,
                gs["warm"] += 1     # allocator growth) must not land inside a capture
                return self._eager_step()
            self._capture(gs)
        self._replay(gs)

    # ---- loss attributes (0-d device tensors; float() synchronises, like the reference's) -----
    def __getattr__(self, name):
        if name.startswith("loss_") or name == "tv_loss":
            L = self.__dict__.get("_loss")
            if L is not None:
                key = name[5:] if name.startswith("loss_") else "tv"
                if key in SLOT:
                    return L[SLOT[key]]
                if key == "G":
                    return (L[SLOT["G_GAN"]] * self.w_gan + L[SLOT["G_L1"]] + L[SLOT["vgg"]] * self.w_vgg
                            + L[SLOT["tv"]] * self.w_tv + L[SLOT["ssim"]] * self.w_ss)
                if key == "D":
                    return (L[SLOT["D_fake"]] + L[SLOT["D_real"]]) * 0.5
        raise AttributeError(name)

    # ---- per-iteration monitoring without the extra generator forward -------------------------------
    def batch_metrics(self):
        """SSIM / PSNR of the step that just ran, from tensors it already produced: no second G forward, no image D2H.

        The reference's loop (train.py:110-118) calls get_img_gen -- a full extra generator forward with the UPDATED
        weights -- copies three images to the host and runs skimage on image 0 of the batch.  Here `ssim` is 1 - loss_ssim
        (the Gaussian-window SSIM of (real_B, fake_B) over the whole batch that the loss kernel already reduced) and `psnr`
        is 10 log10(255^2 / MSE) on the [0, 255]-scaled, clipped images of the whole batch.  These are monitoring values:
        same quantities, not the same numbers as skimage's uniform-window SSIM of one post-update image."""
        fake = (self.fake_B.detach().clamp(-1, 1) + 1) * 127.5
        real = (self.real_B.clamp(-1, 1) + 1) * 127.5
        mse = torch.mean((fake - real) ** 2)
        psnr = 10.0 * torch.log10(255.0 ** 2 / torch.clamp(mse, min=1e-12))
        return {"ssim": 1.0 - self._loss[SLOT["ssim"]], "psnr": psnr}

    # ---- host-side helpers of train.py (pix2pix_model.py:292-310) ----------------------------
    def get_img_tir(self, input):
        self.real_A = input["A"].to(self.device).float().contiguous()
        return ((self.real_A + 1) / 2) * 255

    def get_img_gen(self, input):
        AtoB = self.opt.which_direction == "AtoB"
        self.real_B = input["B" if AtoB else "A"].to(self.device).float().contiguous()
        self.test()
        return ((self.fake_B + 1) / 2) * 255

    def get_img_label(self, input):
        AtoB = self.opt.which_direction == "AtoB"
        self.real_B = input["B" if AtoB else "A"].to(self.device).float().contiguous()
        return ((self.real_B + 1) / 2) * 255

    def get_img_nir(self, input):
        AtoB = self.opt.which_direction == "AtoB"
        self.real_A = input["A" if AtoB else "B"].to(self.device).float().contiguous()
        return ((self.real_A + 1) / 2) * 255
