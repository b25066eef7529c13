"""Generator-only inference model (`--model test`; reference: DSGAN/models/test_model.py:5-42)."""
from . import networks
from .base_model import BaseModel


class TestModel(BaseModel):
    def name(self):
        return "TestModel"

    def initialize(self, opt):
        assert not opt.isTrain
        BaseModel.initialize(self, opt)
        self.precision = getattr(opt, "precision", "bf16")
        self.loss_names, self.visual_names, self.model_names = [], ["real_A", "fake_B"], ["G"]
        networks.KernelNet.precision = self.precision
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.which_model_netG, opt.norm,
                                      not opt.no_dropout, opt.init_type, self.gpu_ids)

    def set_input(self, input):
        self.real_A = input["A"].to(self.device).float().contiguous()
        self.image_paths = input["A_paths"]

    def forward(self):
        ctx = networks.get_ctx(self.device, self.precision)
        was, ctx.no_grad = ctx.no_grad, True
        try:
            self.fake_B = self.netG(self.real_A)
        finally:
            ctx.no_grad = was
            ctx.clear()
