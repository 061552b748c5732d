"""RRDBNet generator -- drop-in for victorca25/traiNNer
codes/models/modules/architectures/RRDBNet_arch.py (RRDBNet :14, RRDB :62,
ResidualDenseBlock_5C :98): same constructor kwargs (networks.py:129-131 passes
defaults.py:36-63), same module tree, parameter names and shapes, so init_weights
(networks.py:71-100, matches on class names containing 'Conv'), checkpoints
(base_model.py:353-400) and optimizers work unchanged.  forward() runs the fused sm_100a engine
(trainner_b200/engine_g.py); there is no eager/CPU path.
"""
import math

import torch
import torch.nn as nn

from . import block as B
from ..engine_g import RRDBNetEngine


class ResidualDenseBlock_5C(nn.Module):
    """Parameter holder of one residual dense block (RRDBNet_arch.py:98-163)."""

    def __init__(self, nf=64, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=None,
                 act_type="leakyrelu", mode="CNA", convtype="Conv2D", spectral_norm=False,
                 gaussian_noise=False, plus=False):
        super().__init__()
        if plus or spectral_norm or norm_type or kernel_size != 3 or stride != 1:
            raise NotImplementedError("ESRGAN+ / spectral-norm / normalised RDB variants are outside the "
                                      "B200 hot path (SURVEY.md section 8a)")
        if act_type.lower() not in ("leakyrelu", "lrelu"):
            raise NotImplementedError("RDB activation must be leakyrelu")
        self.noise = None
        self.conv1x1 = None
        self.gaussian_noise = bool(gaussian_noise)
        kw = dict(bias=bias, pad_type=pad_type, norm_type=norm_type, mode=mode, convtype=convtype)
        self.conv1 = B.conv_block(nf, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv2 = B.conv_block(nf + gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv3 = B.conv_block(nf + 2 * gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv4 = B.conv_block(nf + 3 * gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv5 = B.conv_block(nf + 4 * gc, nf, 3, stride, act_type=None, **kw)

    def forward(self, x):
        raise RuntimeError("ResidualDenseBlock_5C is executed by the enclosing RRDBNet's fused engine")


class RRDB(nn.Module):
    """Parameter holder of a residual-in-residual dense block (RRDBNet_arch.py:62-96)."""

    def __init__(self, nf, nr=3, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=None,
                 act_type="leakyrelu", mode="CNA", convtype="Conv2D", spectral_norm=False,
                 gaussian_noise=False, plus=False):
        super().__init__()
        if nr != 3:
            raise NotImplementedError("only nr=3 (RDB1..RDB3) is on the B200 hot path")
        mk = lambda: ResidualDenseBlock_5C(nf, kernel_size, gc, stride, bias, pad_type, norm_type, act_type,
                                           mode, convtype, spectral_norm=spectral_norm,
                                           gaussian_noise=gaussian_noise, plus=plus)
        self.RDB1 = mk()
        self.RDB2 = mk()
        self.RDB3 = mk()

    def forward(self, x):
        raise RuntimeError("RRDB is executed by the enclosing RRDBNet's fused engine")


class _RRDBNetFunction(torch.autograd.Function):
    """Autograd boundary: one node for the whole generator.  Parameter gradients are accumulated
    by the wgrad kernels directly into param.grad (views of one flat buffer); `anchor` only tells
    autograd that the output depends on trainable state."""

    @staticmethod
    def forward(ctx, x, anchor, engine):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("RRDBNet: the gradient wrt the LR input image is not computed by the B200 "
                                      "engine (the ESRGAN step never needs it); detach the input")
        need_bwd = any(ctx.needs_input_grad)  # (autograd runs forward() with grad mode off)
        out, lease = engine.forward(x, need_bwd)
        ctx.engine, ctx.lease = engine, lease
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.lease is None:
            raise RuntimeError("RRDBNet backward called without a saved forward context")
        ctx.engine.backward(ctx.lease, dout.contiguous().float())
        return None, None, None


class RRDBNet(nn.Module):
    def __init__(self, in_nc, out_nc, nf, nb, nr=3, gc=32, upscale=4, norm_type=None, act_type="leakyrelu",
                 mode="CNA", upsample_mode="upconv", convtype="Conv2D", finalact=None, gaussian_noise=False,
                 plus=False):
        super().__init__()
        n_upscale = int(math.log(upscale, 2))
        if upscale == 3 or 2 ** n_upscale != upscale:
            raise NotImplementedError("B200 RRDBNet supports power-of-two upscale factors")
        if in_nc > 4 or out_nc > 4:
            raise NotImplementedError("B200 RRDBNet expects image-like (<= 4 channel) input/output")
        if nf % 16 or norm_type or finalact:
            raise NotImplementedError("B200 RRDBNet: nf must be a multiple of 16, no norm, no finalact")
        self.in_nc, self.out_nc, self.nf, self.nb, self.gc = in_nc, out_nc, nf, nb, 32
        self.upscale, self.upsample_mode = upscale, upsample_mode
        self.gaussian_noise = bool(gaussian_noise)

        fea_conv = B.conv_block(in_nc, nf, kernel_size=3, norm_type=None, act_type=None, convtype=convtype)
        # NB the reference passes gc=32 literally, ignoring the ctor's gc (RRDBNet_arch.py:24)
        rb_blocks = [RRDB(nf, nr, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=norm_type,
                          act_type=act_type, mode="CNA", convtype=convtype, gaussian_noise=gaussian_noise,
                          plus=plus) for _ in range(nb)]
        LR_conv = B.conv_block(nf, nf, kernel_size=3, norm_type=norm_type, act_type=None, mode=mode,
                               convtype=convtype)
        if upsample_mode == "upconv":
            upsample_block = B.upconv_block
        elif upsample_mode == "pixelshuffle":
            upsample_block = B.pixelshuffle_block
        else:
            raise NotImplementedError("upsample mode [%s] is not found" % upsample_mode)
        upsampler = [upsample_block(nf, nf, act_type=act_type, convtype=convtype) for _ in range(n_upscale)]
        HR_conv0 = B.conv_block(nf, nf, kernel_size=3, norm_type=None, act_type=act_type, convtype=convtype)
        HR_conv1 = B.conv_block(nf, out_nc, kernel_size=3, norm_type=None, act_type=None, convtype=convtype)
        self.model = B.sequential(fea_conv, B.ShortcutBlock(B.sequential(*rb_blocks, LR_conv)), *upsampler,
                                  HR_conv0, HR_conv1)
        self._engine = [RRDBNetEngine(self)]  # list: keep the engine out of nn.Module registration
        self._anchor = None

    # ---- structure helpers
    def _conv_index(self):
        m = self.model
        sub = m[1].sub
        convs = [c for c in m if isinstance(c, nn.Conv2d)]
        rdbs = []
        for i in range(self.nb):
            for name in ("RDB1", "RDB2", "RDB3"):
                rdb = getattr(sub[i], name)
                rdbs.append([getattr(rdb, "conv%d" % j)[0] for j in range(1, 6)])
        # top-level convs: fea, [ups...], hr0, hr1
        return {"fea": convs[0], "rdbs": rdbs, "lr": sub[self.nb], "ups": convs[1:-2], "hr0": convs[-2],
                "hr1": convs[-1]}

    def forward(self, x, outm=None):
        if self.gaussian_noise and self.training:
            raise NotImplementedError("network_G.gaussian (ESRGAN+ noise, block.py:587) is not on the B200 "
                                      "path; set gaussian: false (SURVEY.md 8c)")
        if self._anchor is None or self._anchor.device != x.device:
            self._anchor = torch.zeros(1, device=x.device)
        self._anchor.requires_grad_(any(p.requires_grad for p in self.parameters()))
        xin = x.float().contiguous() if (x.dtype != torch.float32 or not x.is_contiguous()) else x
        y = _RRDBNetFunction.apply(xin, self._anchor, self._engine[0])
        if outm == "scaltanh":
            return (torch.tanh(y) + 1.0) / 2.0
        if outm == "tanh":
            return torch.tanh(y)
        if outm == "sigmoid":
            return torch.sigmoid(y)
        if outm == "clamp":
            return torch.clamp(y, min=0.0, max=1.0)
        return y
