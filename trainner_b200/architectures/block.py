"""Parameter-holder building blocks with the reference's module tree and state_dict keys.

Mirrors the subset of victorca25/traiNNer codes/models/modules/architectures/block.py that the
ESRGAN path instantiates (conv_block :214, act :82, norm :113, sequential :198, ShortcutBlock :184,
Upsample :326, upconv_block :390, pixelshuffle_block :374).  These modules only HOLD parameters at
the reference's attribute paths (so init_weights, load/save_network, optimizers and
requires_grad-by-name keep working); the arithmetic runs in the fused sm_100a
engines, never through these modules' own forward.
"""
import torch.nn as nn

_ENGINE_ONLY = ("this block only holds parameters; the forward pass runs in the fused sm_100a "
                "engine of the enclosing trainner_b200 network (no CPU / eager fallback)")


def act(act_type, inplace=True, neg_slope=0.2):
    t = act_type.lower()
    if t == "relu":
        return nn.ReLU(inplace)
    if t in ("leakyrelu", "lrelu"):
        return nn.LeakyReLU(neg_slope, inplace)
    raise NotImplementedError("activation [%s] is not on the ESRGAN hot path" % act_type)


def norm(norm_type, nc):
    t = norm_type.lower()
    if t == "batch":
        return nn.BatchNorm2d(nc, affine=True)
    raise NotImplementedError("normalization [%s] is not on the ESRGAN hot path" % norm_type)


def sequential(*args):
    if len(args) == 1:
        return args[0]
    mods = []
    for m in args:
        if m is None:
            continue
        if isinstance(m, nn.Sequential):
            mods.extend(m.children())
        else:
            mods.append(m)
    return nn.Sequential(*mods)


def conv_block(in_nc, out_nc, kernel_size, stride=1, bias=True, pad_type="zero", norm_type=None,
               act_type="relu", mode="CNA", convtype="Conv2D", spectral_norm=False):
    if mode != "CNA" or pad_type != "zero" or convtype != "Conv2D" or spectral_norm:
        raise NotImplementedError("only CNA / zero-pad / Conv2D conv blocks are on the ESRGAN hot path")
    padding = (kernel_size - 1) // 2
    c = nn.Conv2d(in_nc, out_nc, kernel_size=kernel_size, stride=stride, padding=padding, bias=bool(bias))
    n = norm(norm_type, out_nc) if norm_type else None
    a = act(act_type) if act_type else None
    mods = [m for m in (c, n, a) if m is not None]
    return nn.Sequential(*mods)


class ShortcutBlock(nn.Module):
    def __init__(self, submodule):
        super().__init__()
        self.sub = submodule

    def forward(self, x):
        raise RuntimeError(_ENGINE_ONLY)


class Upsample(nn.Module):
    def __init__(self, scale_factor=2, mode="nearest"):
        super().__init__()
        self.scale_factor = float(scale_factor)
        self.mode = mode

    def forward(self, x):
        raise RuntimeError(_ENGINE_ONLY)

    def extra_repr(self):
        return "scale_factor=%s, mode=%s" % (self.scale_factor, self.mode)


def upconv_block(in_nc, out_nc, upscale_factor=2, kernel_size=3, act_type="relu", convtype="Conv2D"):
    return sequential(Upsample(scale_factor=upscale_factor, mode="nearest"),
                      conv_block(in_nc, out_nc, kernel_size, 1, act_type=act_type, convtype=convtype))


def pixelshuffle_block(in_nc, out_nc, upscale_factor=2, kernel_size=3, act_type="relu", convtype="Conv2D"):
    conv = conv_block(in_nc, out_nc * upscale_factor ** 2, kernel_size, 1, act_type=None, convtype=convtype)
    return sequential(conv, nn.PixelShuffle(upscale_factor), act(act_type) if act_type else None)
