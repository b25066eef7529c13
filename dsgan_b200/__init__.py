"""dsgan_b200 — B200-native (sm_100a) implementation of DS-GAN's adversarial training step behind the
reference's own Python model/option API (create_model / define_G / define_D / GANLoss / ssim / ms_ssim)."""
__version__ = "0.1.0"
