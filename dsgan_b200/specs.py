"""Parameter inventories of the three networks on the hot path.  Names and shapes are the reference's
state_dict keys (MixConvNeXtML.py:428-459, networks.py:533-579, vgg.py:5-25) so checkpoints interchange."""


def _block(p, dim, plans):
    return [(p + ".shortcut.weight", (plans, dim, 1, 1)), (p + ".dwconv.weight", (dim, 1, 7, 7)),
            (p + ".dwconv.bias", (dim,)), (p + ".pwconv1.weight", (4 * dim, dim)), (p + ".pwconv1.bias", (4 * dim,)),
            (p + ".pwconv2.weight", (plans, 4 * dim)), (p + ".pwconv2.bias", (plans,))]


def _convT(p, cin, cout):
    return [(p + ".weight", (cin, cout, 3, 3)), (p + ".bias", (cout,))]


def _mlka(p, dim):
    s = [(p + ".conv.weight", (dim, dim, 1, 1)), (p + ".conv.bias", (dim,)),
         (p + ".attn.fc1.weight", (dim // 8, dim, 1, 1)), (p + ".attn.relu1.weight", (1,)),
         (p + ".attn.fc2.weight", (dim, dim // 8, 1, 1))]
    for k in (3, 5, 7, 9):
        s += [("%s.X%d.weight" % (p, k), (dim // 4, 1, k, k)), ("%s.X%d.bias" % (p, k), (dim // 4,))]
    return s


ENC = (("c1", 3, 64), ("c2", 64, 128), ("c3", 128, 256), ("c4", 256, 512), ("c5", 512, 1024))
DEC = (("u1", "uc1", 1024, 512), ("u2", "uc2", 512, 256), ("u3", "uc3", 256, 128), ("u4", "uc4", 128, 64))
# (module, input channels, ((branch, pool k, out channels), ...)) — MixConvNeXtML.py:328-426
SKIPS = (("down64", 64, (("to2", 2, 128), ("to4", 4, 256), ("to8", 8, 512), ("to16", 16, 1024))),
         ("down128", 128, (("to4", 2, 256), ("to8", 4, 512), ("to16", 8, 1024))),
         ("down256", 256, (("to8", 2, 512), ("to16", 4, 1024))),
         ("down512", 512, (("to16", 2, 1024),)))


def generator_spec():
    s = []
    for name, cin, cout in ENC:
        s += _block(name, cin, cout)
    for up, blk, cin, cout in DEC:
        s += _convT(up + ".model.0", cin, cout) + _block(blk, cin, cout)
    for mod, cin, branches in SKIPS:
        s += [("%s.%s.1.weight" % (mod, br), (cout, cin, 1, 1)) for br, _k, cout in branches]
    L = "local."
    for a, b in ((3, 32), (32, 64), (64, 128), (128, 256)):
        s += [("%sto%d.weight" % (L, b), (b, a, 1, 1))] + _mlka("%smid%d" % (L, b), b)
    s += _convT(L + "up1.model.0", 256, 128) + [(L + "upc1.0.weight", (128, 256, 1, 1))] + _mlka(L + "upc1.1", 128)
    s += _convT(L + "up2.model.0", 128, 64) + _mlka(L + "upc2", 128)
    s += _convT(L + "up3.model.0", 128, 64) + _mlka(L + "upc3", 128)
    s += _convT(L + "up4.0", 128, 64) + [(L + "shortcut.0.weight", (64, 3, 1, 1))]
    s += [("res.weight", (3, 64, 3, 3)), ("res.bias", (3,))]
    return s


D_LAYERS = ((0, 2), (2, 2), (5, 2), (8, 1), (11, 1))  # (Sequential index, stride), networks.py:543-569


def discriminator_spec(input_nc=6, ndf=32):
    ch = [input_nc, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    s = []
    for i, (m, _st) in enumerate(D_LAYERS):
        s += [("model.%d.weight" % m, (ch[i + 1], ch[i], 4, 4)), ("model.%d.bias" % m, (ch[i + 1],))]
    return s


# torchvision vgg16.features indices regrouped by vgg.py:16-25; "T" = tap, "P" = MaxPool2d(2)
VGG_PLAN = (("to_relu_1_2.0", 3, 64), ("to_relu_1_2.2", 64, 64), "T", "P",
            ("to_relu_2_2.5", 64, 128), ("to_relu_2_2.7", 128, 128), "T", "P",
            ("to_relu_3_3.10", 128, 256), ("to_relu_3_3.12", 256, 256), ("to_relu_3_3.14", 256, 256), "T", "P",
            ("to_relu_4_3.17", 256, 512), ("to_relu_4_3.19", 512, 512), ("to_relu_4_3.21", 512, 512), "T")
VGG_TAIL = (("to_relu_5_3.24", 512, 512), ("to_relu_5_3.26", 512, 512), ("to_relu_5_3.28", 512, 512))


def vgg_spec(with_tail=True):
    """relu5_3's convs are kept as parameters for state_dict compatibility but never executed: the reference
    computes that block and discards it (vgg.py:39-40, pix2pix_model.py:182-186)."""
    s = []
    for e in VGG_PLAN + (VGG_TAIL if with_tail else ()):
        if isinstance(e, tuple):
            s += [(e[0] + ".weight", (e[2], e[1], 3, 3)), (e[0] + ".bias", (e[2],))]
    return s
