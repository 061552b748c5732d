"""CPU oracle for the ESRGAN training hot path of victorca25/traiNNer.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(trainner_b200/) never imports oracle/ and fails loudly without its CUDA library.

It restates, in plain fp32 PyTorch functional ops on state_dict-keyed tensors, the algorithm of
the reference files (paths relative to /root/reference/codes):

  rdb_forward            models/modules/architectures/RRDBNet_arch.py:150-163
  rrdb_forward           RRDBNet_arch.py:89-96
  rrdbnet_forward        RRDBNet_arch.py:15-60, block.py:184-192 (ShortcutBlock),
                         block.py:390-404 (upconv_block), block.py:374-387 (pixelshuffle_block)
  discriminator_vgg_forward   architectures/discriminators.py:16-51, block.py:113-133 (BatchNorm2d)
  vgg19_features         architectures/perceptual.py:103-214 (torchvision vgg19.features[:35])
  generator_losses       models/losses.py:838-862 (calc_losses_regular), :295-340 (PerceptualLoss)
  ragan_g_loss/ragan_d_loss   models/losses.py:406-433, :480-520; modules/loss.py:61-137 (GANLoss vanilla)
  ESRGANStepOracle.optimize_parameters   models/sr_model.py:195-267, base_model.py:805-883

The arithmetic underneath (conv2d, batch_norm, l1_loss, BCEWithLogits, Adam) lives in PyTorch,
an un-vendored dependency of the reference (requirements.txt pins torch==1.9.1; this image has
2.11).  Parity pinning: tests/golden/make_golden.py runs the real reference in the build
container with seeded weights/inputs and commits its outputs; tests/test_oracle_golden.py checks
this oracle against those fixtures (fp32, tolerance 2e-5 relative).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2  # block.py:91 act('leakyrelu', neg_slope=0.2)


def _conv(sd, key, x, stride=1, padding=1):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def rdb_forward(sd, prefix, x):
    """ResidualDenseBlock_5C.forward, RRDBNet_arch.py:150-163 (plus/gaussian branches off)."""
    x1 = F.leaky_relu(_conv(sd, prefix + "conv1.0", x), LRELU_SLOPE)
    x2 = F.leaky_relu(_conv(sd, prefix + "conv2.0", torch.cat((x, x1), 1)), LRELU_SLOPE)
    x3 = F.leaky_relu(_conv(sd, prefix + "conv3.0", torch.cat((x, x1, x2), 1)), LRELU_SLOPE)
    x4 = F.leaky_relu(_conv(sd, prefix + "conv4.0", torch.cat((x, x1, x2, x3), 1)), LRELU_SLOPE)
    x5 = _conv(sd, prefix + "conv5.0", torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def rrdb_forward(sd, prefix, x):
    """RRDB.forward, RRDBNet_arch.py:89-96."""
    out = rdb_forward(sd, prefix + "RDB1.", x)
    out = rdb_forward(sd, prefix + "RDB2.", out)
    out = rdb_forward(sd, prefix + "RDB3.", out)
    return out * 0.2 + x


def rrdbnet_keys(nb, upsample_mode="upconv", upscale=4):
    """state_dict conv keys of RRDBNet in nn.Sequential order (RRDBNet_arch.py:45-46)."""
    n_up = 1 if upscale == 3 else {1: 0, 2: 1, 4: 2, 8: 3}[upscale]
    step = 3  # upconv: (Upsample, conv, act); pixelshuffle: (conv, shuffle, act)
    idx = 2
    ups = []
    for _ in range(n_up):
        ups.append(idx + 1 if upsample_mode == "upconv" else idx)
        idx += step
    return {"fea": "model.0", "trunk": ["model.1.sub.%d." % i for i in range(nb)],
            "lr": "model.1.sub.%d" % nb, "ups": ["model.%d" % u for u in ups],
            "hr0": "model.%d" % idx, "hr1": "model.%d" % (idx + 2)}


def rrdbnet_forward(sd, x, nb, upsample_mode="upconv", upscale=4):
    """RRDBNet.forward, RRDBNet_arch.py:48-60 (outm=None)."""
    k = rrdbnet_keys(nb, upsample_mode, upscale)
    fea = _conv(sd, k["fea"], x)
    t = fea
    for p in k["trunk"]:
        t = rrdb_forward(sd, p, t)
    t = _conv(sd, k["lr"], t)
    t = fea + t  # ShortcutBlock, block.py:190-192
    for u in k["ups"]:
        if upsample_mode == "upconv":  # block.py:390-404
            t = F.interpolate(t, scale_factor=2.0 if upscale != 3 else 3.0, mode="nearest")
            t = F.leaky_relu(_conv(sd, u, t), LRELU_SLOPE)
        else:  # block.py:374-387
            t = F.pixel_shuffle(_conv(sd, u, t), 2 if upscale != 3 else 3)
            t = F.leaky_relu(t, LRELU_SLOPE)
    t = F.leaky_relu(_conv(sd, k["hr0"], t), LRELU_SLOPE)
    return _conv(sd, k["hr1"], t)


def discriminator_vgg_layers(size, in_nc=3, base_nf=64):
    """(key index, cin, cout, kernel, stride, has_bn) per conv of Discriminator_VGG.features
    (discriminators.py:20-35) and the index bookkeeping of B.sequential."""
    layers = []
    idx = 0
    layers.append((idx, in_nc, base_nf, 3, 1, False)); idx += 2          # conv, act
    layers.append((idx, base_nf, base_nf, 4, 2, True)); idx += 3         # conv, bn, act
    cur, nc = size // 2, base_nf
    while cur > 4:
        out = nc * 2 if nc < 512 else nc
        layers.append((idx, nc, out, 3, 1, True)); idx += 3
        layers.append((idx, out, out, 4, 2, True)); idx += 3
        nc, cur = out, cur // 2
    return layers, nc, cur


def discriminator_vgg_forward(sd, x, size, training=True, momentum=0.1, eps=1e-5):
    """Discriminator_VGG.forward, discriminators.py:47-51.  BatchNorm2d in train mode uses batch
    statistics and updates running_mean/var/num_batches_tracked in `sd` in place."""
    layers, nc, cur = discriminator_vgg_layers(size, x.shape[1], sd["features.0.weight"].shape[0])
    for (i, cin, cout, k, s, bn) in layers:
        x = _conv(sd, "features.%d" % i, x, stride=s, padding=1)
        if bn:
            p = "features.%d." % (i + 1)
            if training:
                sd[p + "num_batches_tracked"] += 1
            x = F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                             sd[p + "bias"], training, momentum, eps)
        x = F.leaky_relu(x, LRELU_SLOPE)
    x = x.reshape(x.size(0), -1)
    x = F.leaky_relu(F.linear(x, sd["classifier.0.weight"], sd["classifier.0.bias"]), LRELU_SLOPE)
    return F.linear(x, sd["classifier.2.weight"], sd["classifier.2.bias"])


# torchvision vgg19 'E' config up to conv5_4 (features[:35]); names per perceptual.py:36-46
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
             512, 512, 512, 512]
VGG19_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
               "conv3_4", "conv4_1", "conv4_2", "conv4_3", "conv4_4", "conv5_1", "conv5_2",
               "conv5_3", "conv5_4"]
VGG_MEAN = (0.485, 0.456, 0.406)  # perceptual.py:172-176
VGG_STD = (0.229, 0.224, 0.225)


def vgg19_features(sd, x, use_input_norm=True, prefix="feature_net."):
    """FeatureExtractor.forward for listen_list=['conv5_4'], perceptual.py:201-214."""
    if use_input_norm:
        mean = torch.tensor(VGG_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(VGG_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        x = (x - mean) / std
    ci = 0
    for v in VGG19_CFG:
        if v == "M":
            x = F.max_pool2d(x, kernel_size=2, stride=2)
        else:
            name = VGG19_NAMES[ci]
            x = _conv(sd, prefix + name, x)
            ci += 1
            if name != "conv5_4":
                x = F.relu(x)
    return {"conv5_4": x}


def torchvision_vgg_to_feature_net(tv_sd):
    """Map torchvision vgg19 state_dict keys (features.N.*) to FeatureExtractor's
    feature_net.<name>.* (perceptual.py:152-165)."""
    out = OrderedDict()
    idx = 0
    ci = 0
    for v in VGG19_CFG:
        if v == "M":
            idx += 1
        else:
            name = VGG19_NAMES[ci]
            out["feature_net.%s.weight" % name] = tv_sd["features.%d.weight" % idx]
            out["feature_net.%s.bias" % name] = tv_sd["features.%d.bias" % idx]
            idx += 2
            ci += 1
    return out


def generator_losses(sr, hr, vgg_sd, pixel_weight, feature_weight):
    """GeneratorLoss.calc_losses_regular, losses.py:838-862.  Returns ordered (name, loss)."""
    out = []
    if pixel_weight:
        out.append(("pix-l1", pixel_weight * F.l1_loss(sr, hr)))
    if feature_weight:
        fx = vgg19_features(vgg_sd, sr)
        with torch.no_grad():
            fy = vgg19_features(vgg_sd, hr.detach())
        percep = 0
        percep = percep + F.l1_loss(fx["conv5_4"], fy["conv5_4"]) * 1  # w_l_p = {'conv5_4': 1}
        percep = percep * feature_weight  # PerceptualLoss.perceptual_weight, losses.py:273,324
        out.append(("fea-vgg19-l1", 1 * percep))  # l['weight'] = 1, losses.py:849-856
    return out


def _bce(x, target_is_real):
    t = torch.ones_like(x) if target_is_real else torch.zeros_like(x)
    return F.binary_cross_entropy_with_logits(x, t)


def ragan_g_loss(pred_g_fake, pred_g_real, gan_weight):
    """Adversarial.calculate_gen_loss relativistic branch, losses.py:428-433."""
    pred_g_real = pred_g_real.detach()
    return gan_weight * (_bce(pred_g_real - torch.mean(pred_g_fake), False) +
                         _bce(pred_g_fake - torch.mean(pred_g_real), True)) / 2


def ragan_d_loss(pred_d_fake, pred_d_real):
    """Adversarial.calculate_dis_loss relativistic branch, losses.py:504-520."""
    l_d_real = _bce(pred_d_real - torch.mean(pred_d_fake), True)
    l_d_fake = _bce(pred_d_fake - torch.mean(pred_d_real), False)
    total = (l_d_fake + l_d_real) * 0.5
    logs = OrderedDict(l_d_real=l_d_real.item(), l_d_fake=l_d_fake.item(),
                       D_real=torch.mean(pred_d_real.detach()).item(),
                       D_fake=torch.mean(pred_d_fake.detach()).item())
    return total, logs


class ESRGANStepOracle:
    """SRModel.optimize_parameters restated (sr_model.py:195-267) for the ESRGAN config:
    pix-l1 + fea-vgg19-l1 + vanilla RaGAN, Adam(lr, betas=(0.9, 0.999)), no AMP, no accumulation.

    g_sd / d_sd / vgg_sd are state_dicts keyed as the reference modules' (fp32).  Tensors of
    g_sd and the float params of d_sd become leaf parameters updated in place.
    """

    def __init__(self, g_sd, nb, d_sd=None, d_size=None, vgg_sd=None, pixel_weight=1e-2,
                 feature_weight=1.0, gan_weight=5e-3, lr=1e-4, upsample_mode="upconv",
                 device="cpu"):
        self.nb, self.d_size, self.upsample_mode = nb, d_size, upsample_mode
        self.pixel_weight, self.feature_weight = pixel_weight, feature_weight
        self.gan_weight = gan_weight if d_sd is not None else 0
        self.g = OrderedDict((k, v.detach().clone().to(device).requires_grad_(True))
                             for k, v in g_sd.items())
        self.opt_g = torch.optim.Adam(list(self.g.values()), lr=lr, betas=(0.9, 0.999))
        self.vgg = None
        if vgg_sd is not None and feature_weight:
            self.vgg = OrderedDict((k, v.detach().clone().to(device)) for k, v in vgg_sd.items())
        self.d = None
        if d_sd is not None:
            self.d = OrderedDict()
            for k, v in d_sd.items():
                t = v.detach().clone().to(device)
                if t.is_floating_point() and "running_" not in k:
                    t.requires_grad_(True)
                self.d[k] = t
            self.d_params = [v for v in self.d.values() if v.requires_grad]
            self.opt_d = torch.optim.Adam(self.d_params, lr=lr, betas=(0.9, 0.999))
        self.log_dict = OrderedDict()

    def netG(self, x):
        return rrdbnet_forward(self.g, x, self.nb, self.upsample_mode)

    def netD(self, x):
        return discriminator_vgg_forward(self.d, x, self.d_size, training=True)

    def optimize_parameters(self, lr_img, hr_img):
        # --- G step (sr_model.py:201-252) ---
        if self.d is not None:
            for p in self.d_params:
                p.requires_grad_(False)
        fake = self.netG(lr_img)
        self.fake_H = fake
        l_g_total = 0
        for name, l in generator_losses(fake, hr_img, self.vgg, self.pixel_weight,
                                        self.feature_weight if self.vgg is not None else 0):
            self.log_dict[name] = l.item()
            l_g_total = l_g_total + l
        if self.d is not None:
            pred_g_fake = self.netD(fake)
            pred_g_real = self.netD(hr_img)
            l_g_gan = ragan_g_loss(pred_g_fake, pred_g_real, self.gan_weight)
            self.log_dict["l_g_gan"] = l_g_gan.item()
            l_g_total = l_g_total + l_g_gan
        l_g_total.backward()
        self.opt_g.step()
        self.opt_g.zero_grad()
        # --- D step (sr_model.py:254-267, base_model.py:852-883) ---
        if self.d is not None:
            for p in self.d_params:
                p.requires_grad_(True)
            pred_d_fake = self.netD(fake.detach())
            pred_d_real = self.netD(hr_img)
            l_d_total, logs = ragan_d_loss(pred_d_fake, pred_d_real)
            self.log_dict.update(logs)
            l_d_total.backward()
            self.opt_d.step()
            self.opt_d.zero_grad()
        return self.log_dict
