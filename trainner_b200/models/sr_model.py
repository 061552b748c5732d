"""ESRGAN G/D training step -- mirrors victorca25/traiNNer codes/models/sr_model.py (SRModel :17,
feed_data :115, forward :134, backward_G :162, backward_D :190, optimize_parameters :195, test :269)
and the step-time parts of codes/models/base_model.py (calc_gradients :805, optimizer_step :815,
backward_D_Basic :852, requires_grad :325) for the ESRGAN recipe: pix-l1 + fea-vgg19-l1 + vanilla
relativistic GAN, Adam, bf16 compute with fp32 master weights (no GradScaler needed for bf16).

Differences that do not change values: the 7 per-iteration `.item()` host syncs of the reference
(losses.py:862,516-519, sr_model.py:177) become lazy -- log_dict holds device scalars that are read
back in ONE transfer by get_current_log(); gradients of each network live in one flat buffer, so
the data-parallel exchange is one all-reduce per optimizer step (parallel.py).
"""
from collections import OrderedDict

import torch

from .. import losses, networks
from ..parallel import GradExchange


class LazyLog(OrderedDict):
    """log_dict whose values may be 0-dim device tensors; floats() resolves them with one sync."""

    def floats(self):
        keys = list(self.keys())
        tens = [(k, v) for k, v in self.items() if isinstance(v, torch.Tensor)]
        out = OrderedDict((k, self[k]) for k in keys)
        if tens:
            vals = torch.stack([v.detach().float().reshape(()) for _, v in tens]).tolist()
            for (k, _), f in zip(tens, vals):
                out[k] = f
        return out


class SRModel:
    def __init__(self, opt, step=0, device=None):
        self.opt = opt
        train_opt = opt["train"]
        self.device = torch.device(device if device is not None else "cuda")
        self.is_train = opt.get("is_train", True)
        scale = opt.get("scale", 4)
        self.netG = networks.define_G(opt["network_G"], scale=scale).to(self.device)
        self.netG.train()
        self.cri_gan = bool(train_opt.get("gan_weight"))
        self.netD = None
        if self.cri_gan:
            size = opt["network_D"].get("size") or int(opt["datasets"]["train"]["crop_size"])
            self.netD = networks.define_D(opt["network_D"], size=size).to(self.device)
            self.netD.train()
        netF = None
        if train_opt.get("feature_weight"):
            popt = train_opt.get("perceptual_opt") or {}
            z_norm = bool(((opt.get("datasets") or {}).get("train") or {}).get("znorm", False))   # networks.py:322
            netF = networks.define_F(load_path=popt.get("pretrained_path"), z_norm=z_norm).to(self.device)
        self.generatorlosses = losses.GeneratorLoss(train_opt.get("pixel_weight", 0),
                                                   train_opt.get("feature_weight", 0), netF)
        if self.cri_gan:
            self.adversarial = losses.Adversarial(train_opt.get("gan_type", "vanilla"), train_opt["gan_weight"])
        self.D_update_ratio = train_opt.get("D_update_ratio", 1)
        self.D_init_iters = train_opt.get("D_init_iters", 0)
        adam = dict(betas=(train_opt.get("beta1_G", 0.9), train_opt.get("beta2_G", 0.999)),
                    weight_decay=train_opt.get("weight_decay_G", 0) or 0)
        self.optimizer_G = torch.optim.Adam(self.netG.parameters(), lr=train_opt.get("lr_G", 1e-4), fused=True,
                                            **adam)
        self.optimizers = [self.optimizer_G]
        if self.cri_gan:
            self.optimizer_D = torch.optim.Adam(
                self.netD.parameters(), lr=train_opt.get("lr_D", 1e-4), fused=True,
                betas=(train_opt.get("beta1_D", 0.9), train_opt.get("beta2_D", 0.999)),
                weight_decay=train_opt.get("weight_decay_D", 0) or 0)
            self.optimizers.append(self.optimizer_D)
        # virtual batch (base_model.py:722-734): gradients accumulate over `accumulations` iterations
        ds = (opt.get("datasets") or {}).get("train") or {}
        bs, vb = ds.get("batch_size"), ds.get("virtual_batch_size")
        self.accumulations = (vb // bs) if (bs and vb and vb > bs) else 1
        # gradient clipping of G before its optimizer step (base_model.py:774-787, 911-922)
        self.grad_clip = None
        gc = train_opt.get("grad_clip")
        if gc:
            self.grad_clip = torch.nn.utils.clip_grad_value_ if str(gc).lower() == "value" \
                else torch.nn.utils.clip_grad_norm_
            self.grad_clip_value = train_opt.get("grad_clip_value", 0.1)
            self.grad_history = []
        self.outm = train_opt.get("finalcap", None)
        self.log_dict = LazyLog()
        self.exchange = GradExchange()
        self.exchange.broadcast_params([self.netG] + ([self.netD] if self.netD is not None else []))
        self.optGstep = self.optDstep = False

    @staticmethod
    def _engines(net):
        return [m._engine[0] for m in net.modules() if getattr(m, "_engine", None)]

    def _mark_dirty(self, net):
        """This model owns the optimizers: weights are repacked once per optimizer step instead of
        at every forward (D runs 4 forwards per iteration)."""
        for e in self._engines(net):
            if getattr(e, "packer", None) is not None:
                e.packer.explicit = True
                e.packer.mark_dirty()
            if hasattr(e, "reuse"):
                e.reuse = True   # D(fake) / D(real) of the G step are reused by the D step

    # ------------------------------------------------------------------ reference-facing API
    def feed_data(self, data, need_HR=True):
        """sr_model.py:115-128: H2D of LR (+HR, ref)."""
        self.var_L = data["LR"].to(self.device, non_blocking=True)
        if need_HR:
            self.real_H = data["HR"].to(self.device, non_blocking=True)
            ref = data.get("ref", data["HR"])
            self.var_ref = self.real_H if ref is data["HR"] else ref.to(self.device, non_blocking=True)

    def forward(self):
        self.fake_H = self.netG(self.var_L, outm=self.outm) if self.outm else self.netG(self.var_L)

    @staticmethod
    def requires_grad(model, flag=True):
        for p in model.parameters():
            p.requires_grad = flag

    def backward_G(self):
        l_g_total = 0
        loss_results, self.log_dict = self.generatorlosses(self.fake_H, self.real_H, self.log_dict)
        l_g_total = l_g_total + sum(loss_results) / self.accumulations
        if self.cri_gan:
            l_g_gan = self.adversarial(self.fake_H, self.var_ref, self.netD, "generator")
            self.log_dict["l_g_gan"] = l_g_gan.detach()
            l_g_total = l_g_total + l_g_gan / self.accumulations
        l_g_total.backward()

    def backward_D(self):
        l_d_total, gan_logs = self.adversarial(self.fake_H, self.var_ref, self.netD, "discriminator")
        for k, v in gan_logs.items():
            self.log_dict[k] = v
        (l_d_total / self.accumulations).backward()

    def optimizer_step(self, step, optimizer, opt_flag):
        if step % self.accumulations == 0:
            net = self.netG if opt_flag == "G" else self.netD

            def do_step():
                if opt_flag == "G":
                    self.apply_gradclip()
                optimizer.step()
                optimizer.zero_grad()

            # data-parallel: gradient all-reduce + Adam run on a side stream (parallel.GradExchange.step_async);
            # the main stream joins where the new weights are first used (optimize_parameters / backward_G)
            self.exchange.step_async(opt_flag, net, do_step)
            self._mark_dirty(net)
            if opt_flag == "G":
                self.optGstep = True
            else:
                self.optDstep = True

    def apply_gradclip(self):
        """base_model.py:911-922 (+ get_auto_norm :897-909): clip G's gradients by value or by norm; 'auto' clips
        to the 10th percentile of the gradient-norm history."""
        if self.grad_clip is None:
            return
        value = self.grad_clip_value
        if value == "auto":
            sq = torch.stack([p.grad.detach().float().norm(2) ** 2 for p in self.netG.parameters() if p.grad is not None])
            self.grad_history.append(float(sq.sum().sqrt()))
            value = float(torch.quantile(torch.tensor(self.grad_history, dtype=torch.float32), 0.10))
        self.grad_clip(self.netG.parameters(), value)

    def optimize_parameters(self, step):
        eff_step = step / self.accumulations
        if self.cri_gan:
            self.requires_grad(self.netD, False)
        self.exchange.wait("G")       # G's weights of the previous iteration's update
        self.forward()
        self.exchange.wait("D")       # D's update overlapped the G forward; its weights are needed from here on
        if (self.cri_gan is not True) or (eff_step % self.D_update_ratio == 0 and eff_step > self.D_init_iters):
            self.backward_G()
            self.optimizer_step(step, self.optimizer_G, "G")
        if self.cri_gan:
            self.requires_grad(self.netD, True)
            self.backward_D()
            self.optimizer_step(step, self.optimizer_D, "D")

    def test(self):
        self.exchange.wait("G")
        self.netG.eval()
        with torch.no_grad():
            self.fake_H = self.netG(self.var_L)
        self.netG.train()

    def get_current_log(self):
        return self.log_dict.floats()

    def synchronize(self):
        """join the side stream (pending gradient exchange + optimizer steps); call before reading parameters"""
        self.exchange.wait("G")
        self.exchange.wait("D")

    def get_current_visuals(self, need_HR=True):
        out = OrderedDict()
        out["LR"] = self.var_L.detach()[0].float().cpu()
        out["SR"] = self.fake_H.detach()[0].float().cpu()
        if need_HR:
            out["HR"] = self.real_H.detach()[0].float().cpu()
        return out


def create_model(opt, step=0, device=None):
    """models/__init__.py:46-62: 'sr' and its aliases srgan / srragan / esrgan map to SRModel."""
    kind = str(opt.get("model", "sr")).lower()
    if kind not in ("sr", "srgan", "srragan", "esrgan", "blind"):
        raise NotImplementedError("model [%s] is outside the B200 hot path" % kind)
    return SRModel(opt, step, device)
