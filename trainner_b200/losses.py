"""Loss terms of the ESRGAN step -- mirrors victorca25/traiNNer codes/models/losses.py
(get_loss_fn :23-39 'l1', PerceptualLoss :220-340, Adversarial :343-604 relativistic vanilla form,
GeneratorLoss.calc_losses_regular :838-862) and codes/models/modules/loss.py GANLoss :61-137.
L1 terms run one fused CUDA pass (value + gradient, 128-bit loads, warp-shuffle reduction);
the 16-logit RaGAN term stays in PyTorch.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _L1Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, weight):
        perm = False
        if a.dim() == 4 and not a.is_contiguous() and a.permute(0, 2, 3, 1).is_contiguous():
            perm = True  # logical NCHW over NHWC storage (engine feature maps)
            am = a.permute(0, 2, 3, 1)
            bm = b.permute(0, 2, 3, 1).contiguous().to(am.dtype)
        else:
            am = a.contiguous()
            bm = b.to(am.dtype).contiguous()
        loss, grad = ops.l1_loss_with_grad(am, bm, weight)
        ctx.perm = perm
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        out = grad * g.to(grad.dtype)
        if ctx.perm:
            out = out.permute(0, 3, 1, 2)
        return out, None, None


class L1Loss(nn.Module):
    """mean |x - y| (nn.L1Loss, losses.py:37-39) with a fused value+gradient kernel."""

    def forward(self, x, y, weight=1.0):
        if not x.is_cuda:
            raise RuntimeError("trainner_b200.losses.L1Loss runs only on a CUDA (sm_100a) device")
        return _L1Function.apply(x, y.detach(), float(weight))


class GANLoss(nn.Module):
    """modules/loss.py:61-137, gan_type 'vanilla' (BCEWithLogits against 1.0 / 0.0 labels)."""

    def __init__(self, gan_type="vanilla", real_label_val=1.0, fake_label_val=0.0):
        super().__init__()
        if gan_type.lower() != "vanilla":
            raise NotImplementedError("only gan_type 'vanilla' is on the B200 hot path")
        self.real_label_val, self.fake_label_val = real_label_val, fake_label_val

    def forward(self, x, target_is_real):
        t = torch.full_like(x, self.real_label_val if target_is_real else self.fake_label_val)
        return F.binary_cross_entropy_with_logits(x, t)


class PerceptualLoss(nn.Module):
    """losses.py:295-340 for layer weights {'conv5_4': 1}, no style term."""

    def __init__(self, network, perceptual_weight=1.0, layer_weights=None):
        super().__init__()
        self.network = network
        self.perceptual_weight = perceptual_weight
        self.w_l_p = layer_weights or {"conv5_4": 1}
        self.criterion = L1Loss()

    def forward(self, x, y):
        fea_x = self.network(x)
        with torch.no_grad():
            fea_y = self.network(y.detach())
        percep = 0
        for k, w in self.w_l_p.items():
            percep = percep + self.criterion(fea_x[k], fea_y[k]) * w
        return percep * self.perceptual_weight, None


class Adversarial(nn.Module):
    """losses.py:343-604, relativistic average form ('form' != 'standard'), single-scale D."""

    def __init__(self, gan_type="vanilla", gan_weight=5e-3):
        super().__init__()
        self.cri_gan = GANLoss(gan_type, 1.0, 0.0)
        self.l_gan_w = gan_weight

    def forward(self, fake, real, netD, stage):
        if stage == "generator":  # losses.py:457-468, :428-433
            pred_g_fake = netD(fake)
            pred_g_real = netD(real).detach()
            return self.l_gan_w * (self.cri_gan(pred_g_real - torch.mean(pred_g_fake), False) +
                                   self.cri_gan(pred_g_fake - torch.mean(pred_g_real), True)) / 2
        pred_d_fake = netD(fake.detach())   # losses.py:471-478
        pred_d_real = netD(real)
        l_d_real = self.cri_gan(pred_d_real - torch.mean(pred_d_fake), True)   # :506-509
        l_d_fake = self.cri_gan(pred_d_fake - torch.mean(pred_d_real), False)
        l_d_total = (l_d_fake + l_d_real) * 0.5
        logs = OrderedDict(l_d_real=l_d_real.detach(), l_d_fake=l_d_fake.detach(),
                           D_real=torch.mean(pred_d_real.detach()), D_fake=torch.mean(pred_d_fake.detach()))
        return l_d_total, logs


class GeneratorLoss(nn.Module):
    """losses.py:607-962 restricted to pix-l1 + fea-vgg19-l1 (loss_list order: pixel, feature)."""

    def __init__(self, pixel_weight=1e-2, feature_weight=1.0, netF=None):
        super().__init__()
        self.pixel_weight, self.feature_weight = pixel_weight, feature_weight
        self.pix = L1Loss() if pixel_weight else None
        self.fea = PerceptualLoss(netF, perceptual_weight=feature_weight) if (feature_weight and netF is not None) else None

    def forward(self, sr, hr, log_dict):
        results = []
        if self.pix is not None:
            l = self.pixel_weight * self.pix(sr, hr)
            results.append(l)
            log_dict["pix-l1"] = l.detach()
        if self.fea is not None:
            percep, _ = self.fea(sr, hr)
            l = 1 * percep
            results.append(l)
            log_dict["fea-vgg19-l1"] = l.detach()
        return results, log_dict
