"""Fused Adam over a network's flat parameter buffer (torch.optim.Adam semantics, pix2pix_model.py:122-125)."""
import torch


class FlatAdam:
    def __init__(self, net, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        self.net = net
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps}]
        self.t = 0
        self.m = self.v = None

    def zero_grad(self):
        """Gradients are re-zeroed (kernels accumulate into them), SURVEY Q15."""
        flat, grad, _ = self.net.flat_buffers()
        ctx = self.net.ctx()
        ctx.zero_(grad)

    def step(self, grad_scale=1.0):
        flat, grad, _ = self.net.flat_buffers()
        ctx = self.net.ctx()
        if self.m is None or self.m.device != flat.device:
            self.m = torch.zeros_like(flat)
            self.v = torch.zeros_like(flat)
        self.t += 1
        g = self.param_groups[0]
        ctx.L.adam_step(flat.data_ptr(), grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), flat.numel(),
                        float(g["lr"]), g["betas"][0], g["betas"][1], g["eps"], self.t, float(grad_scale), None,
                        ctx.stream)

    # ---- CUDA-graph form: the kernel reads lr/(1-b1^t) and sqrt(1-b2^t) from device memory ---------------------
    def _state(self):
        flat, grad, _ = self.net.flat_buffers()
        if self.m is None or self.m.device != flat.device:
            self.m = torch.zeros_like(flat)
            self.v = torch.zeros_like(flat)
        if getattr(self, "hyper", None) is None or self.hyper.device != flat.device:
            self.hyper = torch.zeros(2, dtype=torch.float32, device=flat.device)
        return flat, grad

    def record_step(self, grad_scale=1.0):
        """Launch (or capture) the update kernel without advancing t; pair every replay with advance()."""
        flat, grad = self._state()
        ctx = self.net.ctx()
        g = self.param_groups[0]
        ctx.L.adam_step_dev(flat.data_ptr(), grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), flat.numel(),
                            self.hyper.data_ptr(), g["betas"][0], g["betas"][1], g["eps"], float(grad_scale), None, ctx.stream)

    def advance(self):
        """t += 1 and refresh the device-side scalars (stream-ordered before the graph that contains record_step)."""
        self._state()
        ctx = self.net.ctx()
        self.t += 1
        g = self.param_groups[0]
        ctx.L.adam_hyper(self.hyper.data_ptr(), float(g["lr"]), g["betas"][0], g["betas"][1], self.t, ctx.stream)

    def state_dict(self):
        return {"t": self.t, "m": self.m, "v": self.v, "param_groups": self.param_groups}
