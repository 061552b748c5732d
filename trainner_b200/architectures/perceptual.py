"""VGG feature extractor for the perceptual loss -- drop-in for victorca25/traiNNer
codes/models/modules/architectures/perceptual.py:73-214 (FeatureExtractor): same constructor
(networks.py:358-363), same `feature_net.<layer name>.*` parameters and `mean`/`std` buffers,
returns {layer name: feature map} (logical NCHW, channels-last bf16 storage).  Weights come from
torchvision exactly like the reference (pretrained download or load_path); the convolutions run in
the fused sm_100a engine (trainner_b200/engine_f.py).
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from ..engine_f import FeatureEngine

# layer names of torchvision's VGG configs (perceptual.py:10-46)
_CFG = {
    "vgg11": [1, 1, 2, 2, 2], "vgg13": [2, 2, 2, 2, 2], "vgg16": [2, 2, 3, 3, 3], "vgg19": [2, 2, 4, 4, 4],
}


def vgg_layer_names(net):
    names = []
    for b, reps in enumerate(_CFG[net], start=1):
        for r in range(1, reps + 1):
            names += ["conv%d_%d" % (b, r), "relu%d_%d" % (b, r)]
        names.append("pool%d" % b)
    return names


class _FeatureFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, listen):
        need_bwd = any(ctx.needs_input_grad)  # (autograd runs forward() with grad mode off)
        outs, lease = engine.forward(x, need_bwd, listen)
        ctx.engine, ctx.lease, ctx.listen = engine, lease, listen
        # logical NCHW view of the NHWC storage (zero-copy, channels_last strides)
        return tuple(outs[k].permute(0, 3, 1, 2) for k in listen)

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.lease is None:
            raise RuntimeError("FeatureExtractor backward called without a saved forward context")
        grads = {}
        for k, g in zip(ctx.listen, gouts):
            grads[k] = None if g is None else g.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        dx = ctx.engine.backward(ctx.lease, grads)
        return dx, None, None


class FeatureExtractor(nn.Module):
    def __init__(self, listen_list=None, net: str = "vgg19", use_input_norm: bool = True, z_norm: bool = False,
                 requires_grad: bool = False, remove_pooling: bool = False, pooling_stride: int = 2,
                 change_padding: bool = False, load_path=None):
        super().__init__()
        if "vgg" not in net or "bn" in net:
            raise NotImplementedError("B200 FeatureExtractor: plain VGG backbones only")
        if requires_grad or remove_pooling or pooling_stride != 2 or change_padding:
            raise NotImplementedError("B200 FeatureExtractor: frozen net with stock 2x2 pooling only")
        import torchvision.models.vgg as vgg

        self.use_input_norm = use_input_norm
        self.znorm = z_norm
        self.listen_list = set(listen_list)
        self.names = vgg_layer_names(net)
        max_idx = max(self.names.index(v) for v in listen_list)
        if load_path and os.path.exists(load_path):
            feature_net = getattr(vgg, net)(weights=None)
            feature_net.load_state_dict(torch.load(load_path, map_location="cpu"))
        else:
            feature_net = getattr(vgg, net)(pretrained=True)
        features = feature_net.features[:max_idx + 1]
        modified = OrderedDict()
        for k, v in zip(self.names, features):
            modified[k] = nn.MaxPool2d(kernel_size=2, stride=pooling_stride) if "pool" in k else v
        self.feature_net = nn.Sequential(modified)
        if self.use_input_norm:
            self.register_buffer("mean", torch.tensor([[[0.485]], [[0.456]], [[0.406]]]))
            self.register_buffer("std", torch.tensor([[[0.229]], [[0.224]], [[0.225]]]))
        self.feature_net.eval()
        for p in self.parameters():
            p.requires_grad = False
        self._engine = [FeatureEngine(self)]

    def forward(self, x):
        if self.znorm:
            x = (x + 1) / 2
        xin = x.float().contiguous() if (x.dtype != torch.float32 or not x.is_contiguous()) else x
        listen = tuple(k for k in self.names if k in self.listen_list)
        for k in listen:
            if k.startswith("relu") or k.startswith("pool"):
                # taps on relu/pool outputs are served by the stored post-activation tensors
                pass
        outs = _FeatureFunction.apply(xin, self._engine[0], listen)
        return {k: o.clone() for k, o in zip(listen, outs)}
