"""Arch registry for the B200 hot path -- mirrors the slice of victorca25/traiNNer
codes/models/networks.py that the ESRGAN step uses (get_network :107-255, init_weights :71-100,
define_G/define_D/define_F :267,283,316) and provides the hook that rebinds the reference's own
registry to these classes: get_network resolves RRDBNet_arch.RRDBNet,
discriminators.Discriminator_VGG and perceptual.FeatureExtractor by attribute lookup at call time.
"""
import functools

import torch.nn as nn
from torch.nn import init

from .architectures import RRDBNet_arch, discriminators, perceptual


def weights_init_kaiming(m, scale=1, bias_fill=0):
    """networks.py:41-54: kaiming_normal_ (fan_in, a=0) then * scale for every module whose class
    name contains 'Conv' or 'Linear'; BatchNorm weight 1 / bias 0."""
    classname = m.__class__.__name__
    if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
        init.kaiming_normal_(m.weight)
        m.weight.data *= scale
        if hasattr(m, "bias") and m.bias is not None:
            m.bias.data.fill_(bias_fill)
    elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
        init.constant_(m.weight, 1)
        if hasattr(m, "bias") and m.bias is not None:
            m.bias.data.fill_(bias_fill)


def init_weights(net, init_type="kaiming", scale=1):
    if init_type != "kaiming":
        raise NotImplementedError("only the reference default init_type='kaiming' is mirrored")
    net.apply(functools.partial(weights_init_kaiming, scale=scale))


_G_KEYS = ("norm_type", "mode", "nf", "nb", "nr", "in_nc", "out_nc", "gc", "convtype", "act_type",
           "gaussian_noise", "plus", "finalact", "upscale", "upsample_mode")
_D_KEYS = ("in_nc", "base_nf", "norm_type", "mode", "act_type", "convtype", "arch", "size")


def define_G(opt_net, scale=4):
    """opt_net: the reference's opt['network_G'] after get_network_defaults (defaults.py:36-63)."""
    o = dict(opt_net)
    kind = str(o.pop("type", "rrdb_net")).lower()
    if kind not in ("rrdb_net", "esrgan"):
        raise NotImplementedError("network_G type [%s] is outside the B200 hot path" % kind)
    init_type, init_scale = o.pop("init_type", "kaiming"), o.pop("init_scale", 0.1)
    o.pop("strict", None)
    if "gaussian" in o:
        o["gaussian_noise"] = o.pop("gaussian")
    if "scale" in o:
        o["upscale"] = o.pop("scale")
    kw = {"in_nc": 3, "out_nc": 3, "nf": 64, "nb": 23, "upscale": scale}
    kw.update({k: v for k, v in o.items() if k in _G_KEYS})
    net = RRDBNet_arch.RRDBNet(**kw)
    init_weights(net, init_type, init_scale)
    return net


def define_D(opt_net, size=None):
    o = dict(opt_net)
    kind = str(o.pop("type", "discriminator_vgg")).lower()
    if kind != "discriminator_vgg":
        raise NotImplementedError("network_D type [%s] is outside the B200 hot path" % kind)
    init_type, init_scale = o.pop("init_type", "kaiming"), o.pop("init_scale", 0.1)
    kw = {"in_nc": 3, "base_nf": 64, "size": size}
    kw.update({k: v for k, v in o.items() if k in _D_KEYS})
    net = discriminators.Discriminator_VGG(**kw)
    init_weights(net, init_type, init_scale)
    return net


def define_F(listen_list=("conv5_4",), net="vgg19", use_input_norm=True, z_norm=False, load_path=None):
    return perceptual.FeatureExtractor(listen_list=list(listen_list), net=net, use_input_norm=use_input_norm,
                                       z_norm=z_norm, load_path=load_path)


def install_into_reference(ref_architectures_pkg):
    """Rebind the reference's architecture classes to the B200 ones, e.g.

        from models.modules import architectures          # the reference package
        import trainner_b200.networks as b200n
        b200n.install_into_reference(architectures)         # before models.create_model(opt)

    after which the reference's YAML-driven loop (train.py) builds and trains the B200 modules
    unchanged.  Returns the original classes so the caller can restore them.
    """
    import importlib

    saved = {}
    for modname, names, src in (("RRDBNet_arch", ("RRDBNet", "RRDB", "ResidualDenseBlock_5C"), RRDBNet_arch),
                                ("discriminators", ("Discriminator_VGG",), discriminators),
                                ("perceptual", ("FeatureExtractor",), perceptual)):
        mod = importlib.import_module(ref_architectures_pkg.__name__ + "." + modname)
        for n in names:
            saved[(modname, n)] = getattr(mod, n)
            setattr(mod, n, getattr(src, n))
    return saved
