"""VGG-style discriminator -- drop-in for victorca25/traiNNer
codes/models/modules/architectures/discriminators.py:16-51 (Discriminator_VGG): same constructor
(networks.py:206-208 + defaults.py:342-360), same `features.N.*` / `classifier.{0,2}.*` state_dict
(incl. BatchNorm running_mean / running_var / num_batches_tracked).  The conv/BN/LeakyReLU feature
stack runs in the fused sm_100a engine (trainner_b200/engine_d.py); the 2-layer classifier is a
plain library GEMM (torch / cuBLAS), < 0.01 % of the step.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import block as B
from ..engine_d import DiscriminatorEngine


class _DFeaturesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, engine, training):
        need_bwd = any(ctx.needs_input_grad)  # (autograd runs forward() with grad mode off)
        feat, lease = engine.forward(x, need_bwd, training)
        ctx.engine, ctx.lease, ctx.xgrad = engine, lease, ctx.needs_input_grad[0]
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        if ctx.lease is None:
            raise RuntimeError("Discriminator_VGG backward called without a saved forward context")
        dx = ctx.engine.backward(ctx.lease, dfeat.contiguous().float(), ctx.xgrad)
        return dx, None, None, None


class Discriminator_VGG(nn.Module):
    def __init__(self, size, in_nc, base_nf, norm_type="batch", act_type="leakyrelu", mode="CNA",
                 convtype="Conv2D", arch="ESRGAN"):
        super().__init__()
        if norm_type != "batch" or act_type.lower() not in ("leakyrelu", "lrelu") or mode != "CNA":
            raise NotImplementedError("B200 Discriminator_VGG: batch norm + leakyrelu + CNA only")
        if in_nc > 4 or base_nf % 16:
            raise NotImplementedError("B200 Discriminator_VGG: in_nc <= 4 and base_nf % 16 == 0")
        self.size, self.in_nc, self.base_nf = size, in_nc, base_nf
        blocks = [B.conv_block(in_nc, base_nf, kernel_size=3, stride=1, norm_type=None, act_type=act_type, mode=mode),
                  B.conv_block(base_nf, base_nf, kernel_size=4, stride=2, norm_type=norm_type, act_type=act_type,
                               mode=mode)]
        cur_size = size // 2
        cur_nc = base_nf
        while cur_size > 4:
            out_nc = cur_nc * 2 if cur_nc < 512 else cur_nc
            blocks.append(B.conv_block(cur_nc, out_nc, kernel_size=3, stride=1, norm_type=norm_type,
                                       act_type=act_type, mode=mode))
            blocks.append(B.conv_block(out_nc, out_nc, kernel_size=4, stride=2, norm_type=norm_type,
                                       act_type=act_type, mode=mode))
            cur_nc = out_nc
            cur_size //= 2
        self.features = B.sequential(*blocks)
        hidden = 128 if arch == "PPON" else 100
        self.classifier = nn.Sequential(nn.Linear(cur_nc * cur_size * cur_size, hidden), nn.LeakyReLU(0.2, True),
                                        nn.Linear(hidden, 1))
        self._engine = [DiscriminatorEngine(self)]
        self._anchor = None

    def forward(self, x):
        if self._anchor is None or self._anchor.device != x.device:
            self._anchor = torch.zeros(1, device=x.device)
        self._anchor.requires_grad_(any(p.requires_grad for p in self.features.parameters()))
        xin = x.float().contiguous() if (x.dtype != torch.float32 or not x.is_contiguous()) else x
        feat = _DFeaturesFunction.apply(xin, self._anchor, self._engine[0], self.training)
        with torch.autocast("cuda", enabled=False):
            h = feat.view(feat.size(0), -1)
            h = F.leaky_relu(F.linear(h, self.classifier[0].weight.float(), self.classifier[0].bias.float()), 0.2)
            return F.linear(h, self.classifier[2].weight.float(), self.classifier[2].bias.float())
