"""Full-size parity on the B200 against the UNMODIFIED reference (baseline/_ref, staged by
tools/stage_reference.py): BASELINE config 2 at its real size -- RRDBNet nb=23, 16 x 64^2 -> 256^2,
Discriminator_VGG(256), VGG19 conv5_4, L1 + perceptual + vanilla RaGAN, Adam.

The reference's own SRModel (codes/models/sr_model.py:17; feed_data :115, optimize_parameters :195)
runs ON THE SAME GPU through PyTorch/cuDNN twice -- fp32 (TF32 off) = the ground truth, and under
bf16 autocast = the reference's own reduced-precision path, the like-for-like yardstick -- and
trainner_b200 runs the identical step from identical weights on the identical batch.

Tolerances (no absolute floors on the scalars):
  * the log_dict scalars pix-l1, fea-vgg19-l1, l_g_gan, l_d_real, l_d_fake, D_real:
    |v - v_ref32| <= 2e-2 |v_ref32|;
  * D_fake / D_real are MEANS of 16 raw logits that nearly cancel (measured: mean -7e-3, std 7e-3), and D(fake)
    amplifies the ~1.3e-2 relative bf16 error of the SR input through ten BatchNorm layers: the reference's OWN
    bf16 path misses 2e-2 on D_fake (measured 2.5e-2).  They are therefore checked where the noise can be
    measured, per logit: the step's D(fake) / D(real) forwards are repeated from the initial weights and
    rms(logit error) of the CUDA path must be <= 1.25 (1 + 2/sqrt(16)) x that of the reference's bf16 path (the factor
    in brackets is the scatter of an rms estimated from 16 samples); the logged means must be
    within max(2e-2 |v|, 3 sigma) with sigma = rms(reference-bf16 logit error) / sqrt(16);
  * SR (fake_H): rel-L2 vs reference-fp32 <= max(1e-2, 1.25 x the reference-bf16 rel-L2);
  * every gradient tensor of G and D: rel-L2 error vs reference-fp32 <= 1.25 x the error of the reference's
    bf16 path on that tensor, x (1 + 2/sqrt(numel)) for the scatter of an rms estimated from numel samples
    (matters for the 32-/64-element biases only; fp32-rounding-level errors <= 1e-5 pass);
  * every updated parameter tensor: Adam's first update is -lr * sign(g) wherever |g| >> eps, so the error of an
    update is a COUNT of sign flips (elements whose gradient is within the rounding noise of zero); per tensor
    flips <= 1.25 x flips of the reference-bf16 path + 3 + 3 sqrt(flips_ref) (Poisson slack for small tensors),
    and over all tensors of a network flips <= 1.1 x the reference-bf16 total; BatchNorm running statistics and
    num_batches_tracked directly.
G uses network_G.init_scale 0.3 (the reference's option, networks.py:116-120): with the default
0.1 an untrained 23-block G outputs ~1e-4 and every comparison would be vacuous.
"""
import os
import sys
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import reference_arm as RA  # noqa: E402

needs_ref = pytest.mark.skipif(not RA.reference_available(),
                               reason="reference tree not staged: run tools/stage_reference.py in the build container")

TORCH_HOME = "/tmp/_parity_torch_home"
NB, HR, BS = 23, 256, 16


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _batch(bs=BS, hr=HR, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return {"LR": torch.rand(bs, 3, hr // 4, hr // 4, generator=g).cuda(),
            "HR": torch.rand(bs, 3, hr, hr, generator=g).cuda()}


def _snapshot_hooks(model, store):
    """record the gradients each optimizer is about to apply (optimizer_step zeroes them afterwards)"""
    def hook_for(name, net):
        def hook(opt, args, kwargs):
            store[name] = OrderedDict((k, p.grad.detach().float().clone()) for k, p in RA.unwrap(net).named_parameters()
                                      if p.grad is not None)
        return hook
    model.optimizer_G.register_step_pre_hook(hook_for("G", model.netG))
    model.optimizer_D.register_step_pre_hook(hook_for("D", model.netD))


def _state(net):
    return OrderedDict((k, v.detach().clone()) for k, v in RA.unwrap(net).state_dict().items())


def _reference_run(precision, g_sd, d_sd, batch, nb=NB, hr=HR, steps=1):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    model, _ = RA.create_reference_model(torch_home=TORCH_HOME, precision=precision, nb=nb, hr_size=hr, use_gan=True,
                                         use_fea=True, pixel_weight=1e-2, feature_weight=1.0, gan_weight=5e-3, gpu=True,
                                         batch_size=batch["LR"].shape[0], init_scale=0.3)
    if g_sd is not None:
        RA.unwrap(model.netG).load_state_dict(g_sd)
        RA.unwrap(model.netD).load_state_dict(d_sd)
    out = {"g0": _state(model.netG), "d0": _state(model.netD), "grads": {}, "logs": []}
    _snapshot_hooks(model, out["grads"])
    for s in range(1, steps + 1):
        model.feed_data(batch)
        model.optimize_parameters(s)
        out["logs"].append(OrderedDict((k, float(v)) for k, v in model.log_dict.items()))
        if s == 1:
            out["sr"] = model.fake_H.detach().float().clone()
            out["g1"], out["d1"] = _state(model.netG), _state(model.netD)
    del model
    torch.cuda.empty_cache()
    return out


def _b200_opt(nb, hr):
    return {"model": "sr", "scale": 4, "is_train": True, "datasets": {"train": {"crop_size": hr}},
            "network_G": {"type": "esrgan", "nb": nb, "nf": 64, "gc": 32, "gaussian": False, "upsample_mode": "upconv"},
            "network_D": {"type": "discriminator_vgg"},
            "train": {"pixel_criterion": "l1", "pixel_weight": 1e-2, "feature_criterion": "l1", "feature_weight": 1,
                      "gan_type": "vanilla", "gan_weight": 5e-3, "lr_G": 1e-4, "lr_D": 1e-4,
                      "perceptual_opt": {"pretrained_path": RA.make_vgg19_checkpoint(TORCH_HOME)}}}


def _b200_run(g_sd, d_sd, batch, nb=NB, hr=HR, steps=1):
    from trainner_b200.models.sr_model import create_model
    model = create_model(_b200_opt(nb, hr))
    model.netG.load_state_dict(g_sd)
    model.netD.load_state_dict(d_sd)
    out = {"grads": {}, "logs": []}
    _snapshot_hooks(model, out["grads"])
    for s in range(1, steps + 1):
        model.feed_data(batch)
        model.optimize_parameters(s)
        out["logs"].append(model.get_current_log())
        if s == 1:
            out["sr"] = model.fake_H.detach().float().clone()
            out["g1"], out["d1"] = _state(model.netG), _state(model.netD)
    return out


def _compare_tensors(what, ours, ref16, ref32, zero_abs):
    """per tensor: err(ours, ref32) <= 1.25 err(ref16, ref32); numerically-zero references by absolute size"""
    bad, e_all, e_ref_all = [], [], []
    for k, t32 in ref32.items():
        if not t32.is_floating_point():
            continue
        n32 = float(t32.double().norm())
        if n32 <= zero_abs * t32.numel() ** 0.5:
            # gradient that is exactly 0 in exact arithmetic (bias in front of BatchNorm): only rounding noise
            if float(ours[k].double().norm()) > max(4.0 * float(ref16[k].double().norm()), zero_abs * t32.numel() ** 0.5):
                bad.append((k, "zero-ref", float(ours[k].double().norm()), float(ref16[k].double().norm())))
            continue
        e, e_ref = rel(ours[k], t32), rel(ref16[k], t32)
        e_all.append(e)
        e_ref_all.append(e_ref)
        # both errors are rms estimates over numel samples: allow the estimate's own scatter on small tensors
        if e > max(1.25 * (1.0 + 2.0 / t32.numel() ** 0.5) * e_ref, 1e-5):
            bad.append((k, e, e_ref))
    med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
    print("%s: %d tensors, median rel-err trainner_b200 %.4f | reference bf16 %.4f; worst ratio %.2f" %
          (what, len(e_all), med(e_all), med(e_ref_all), max(a / max(b, 1e-30) for a, b in zip(e_all, e_ref_all))))
    return bad


def _discriminator_logits(r32, r16, ours, batch):
    """The D-step's forwards (losses.py:471-478: netD(fake.detach()), netD(real)) repeated from the initial D
    weights on each path's own fake_H: reference fp32, reference bf16 autocast, trainner_b200."""
    from models.modules.architectures import discriminators as ref_disc
    from trainner_b200.architectures import discriminators as b200_disc
    dref = ref_disc.Discriminator_VGG(HR, 3, 64).cuda()
    db = b200_disc.Discriminator_VGG(HR, 3, 64).cuda()
    out = {}
    with torch.no_grad():
        for key, x32, x16, xb in (("fake", r32["sr"], r16["sr"], ours["sr"]), ("real", batch["HR"], batch["HR"], batch["HR"])):
            dref.load_state_dict(r32["d0"]); dref.train()
            l32 = dref(x32).float().flatten().double()
            dref.load_state_dict(r32["d0"])
            with torch.autocast("cuda", dtype=torch.bfloat16):
                l16 = dref(x16).float().flatten().double()
            db.load_state_dict(r32["d0"]); db.train()
            lb = db(xb).float().flatten().double()
            out[key] = (l32, l16, lb)
    return out


@needs_ref
def test_full_size_step_vs_unmodified_reference():
    batch = _batch()
    r32 = _reference_run("fp32", None, None, batch)
    r16 = _reference_run("bf16", r32["g0"], r32["d0"], batch)
    ours = _b200_run(r32["g0"], r32["d0"], batch)
    # ---- the 7 log_dict scalars, relative, no floor
    assert list(ours["logs"][0].keys()) == list(r32["logs"][0].keys())
    rows = []
    for k, v in r32["logs"][0].items():
        rows.append((k, v, r16["logs"][0][k], ours["logs"][0][k]))
    print("log_dict (reference fp32 | reference bf16 | trainner_b200):")
    for k, a, b, c in rows:
        print("  %-14s % .6e  % .6e (rel %.2e)  % .6e (rel %.2e)" % (k, a, b, abs(b - a) / abs(a), c, abs(c - a) / abs(a)))
    problems = []
    for k, a, b, c in rows:
        if k in ("D_real", "D_fake"):
            continue   # checked per logit below
        if not abs(c - a) <= 2e-2 * abs(a):
            problems.append(("log", k, a, c))
    # ---- D_real / D_fake: repeat the D-step's forwards (initial D weights, the step's own fake_H) per logit
    logit_rows = _discriminator_logits(r32, r16, ours, batch)
    for name, key in (("D_fake", "fake"), ("D_real", "real")):
        l32, l16, lb = logit_rows[key]
        e16, eb = float((l16 - l32).pow(2).mean().sqrt()), float((lb - l32).pow(2).mean().sqrt())
        sigma = e16 / (l32.numel() ** 0.5)
        a, c = r32["logs"][0][name], ours["logs"][0][name]
        print("%s logits: mean %.4e std %.3e | rms error reference-bf16 %.3e, trainner_b200 %.3e | logged mean off by "
              "%.3e (%.1f sigma)" % (name, float(l32.mean()), float(l32.std()), e16, eb, abs(c - a), abs(c - a) / sigma))
        assert abs(float(l32.mean()) - a) <= 1e-4 * abs(a) + 1e-7, "the repeated forward must reproduce the logged mean"
        if not eb <= 1.25 * (1.0 + 2.0 / l32.numel() ** 0.5) * e16:   # rms over 16 logits: +-18 % scatter of the estimate
            problems.append(("logits", name, eb, e16))
        if not abs(c - a) <= max(2e-2 * abs(a), 3.0 * sigma):
            problems.append(("log", name, a, c, sigma))
    # ---- SR
    e_sr, e_sr16 = rel(ours["sr"], r32["sr"]), rel(r16["sr"], r32["sr"])
    print("SR rel-L2: trainner_b200 %.4e | reference bf16 %.4e; SR std %.3e" % (e_sr, e_sr16, float(r32["sr"].std())))
    assert float(r32["sr"].std()) > 1e-2, "degenerate generator output (init_scale not applied?)"
    if not e_sr <= max(1e-2, 1.25 * e_sr16):
        problems.append(("sr", e_sr, e_sr16))
    # ---- gradients, every tensor
    for net in ("G", "D"):
        bad = _compare_tensors("grad " + net, ours["grads"][net], r16["grads"][net], r32["grads"][net], 1e-9)
        if bad:
            problems.append(("grad " + net, len(bad), bad[:8]))
    # ---- every updated parameter tensor (sign flips of Adam's first update) and the BatchNorm running statistics
    lr = 1e-4
    for net, k0, k1 in (("G", "g0", "g1"), ("D", "d0", "d1")):
        tot_o = tot_r = 0
        for k, p0 in r32[k0].items():
            if not p0.is_floating_point():
                assert int(ours[k1][k]) == int(r32[k1][k]), k   # num_batches_tracked
                continue
            u32 = r32[k1][k].double() - p0.double()
            if "running_" in k:
                eo, er = rel(ours[k1][k].double() - p0.double(), u32), rel(r16[k1][k].double() - p0.double(), u32)
                if not eo <= max(1.25 * er, 1e-5):
                    problems.append(("bn stat", k, eo, er))
                continue
            fo = int(((ours[k1][k].double() - p0.double() - u32).abs() > lr).sum())
            fr = int(((r16[k1][k].double() - p0.double() - u32).abs() > lr).sum())
            tot_o += fo
            tot_r += fr
            if fo > 1.25 * fr + 3 + 3 * fr ** 0.5:
                problems.append(("update " + net, k, fo, fr, p0.numel()))
        n_el = sum(v.numel() for v in r32[k0].values() if v.is_floating_point())
        print("update %s: sign flips of Adam's first update vs reference fp32: trainner_b200 %d | reference bf16 %d of %d "
              "elements" % (net, tot_o, tot_r, n_el))
        if tot_o > 1.1 * tot_r + 10:
            problems.append(("update total " + net, tot_o, tot_r))
    assert not problems, problems


@needs_ref
def test_reference_srmodel_drives_b200_modules_full_size():
    """The reference's OWN SRModel / create_model / losses / optimizers (codes/train.py:224-238 loop body), with
    networks.install_into_reference rebinding the architecture classes (networks.py:129-131, 206-208, :358):
    3 full-size steps, scalars against the stock reference run from the same weights and batches."""
    import trainner_b200.networks as b200n
    batch = _batch(seed=4321)
    stock = _reference_run("bf16", None, None, batch, steps=3)
    stock32 = _reference_run("fp32", stock["g0"], stock["d0"], batch, steps=3)
    RA.install_shims(TORCH_HOME)
    from models.modules import architectures
    saved = b200n.install_into_reference(architectures)
    try:
        model, _ = RA.create_reference_model(torch_home=TORCH_HOME, precision="fp32", nb=NB, hr_size=HR, use_gan=True,
                                             use_fea=True, pixel_weight=1e-2, feature_weight=1.0, gan_weight=5e-3,
                                             gpu=True, batch_size=BS, init_scale=0.3)
        from trainner_b200.architectures import RRDBNet_arch, discriminators, perceptual
        assert isinstance(RA.unwrap(model.netG), RRDBNet_arch.RRDBNet)
        assert isinstance(RA.unwrap(model.netD), discriminators.Discriminator_VGG)
        assert isinstance(RA.unwrap(RA.perceptual_network(model)), perceptual.FeatureExtractor)
        RA.unwrap(model.netG).load_state_dict(stock["g0"])
        RA.unwrap(model.netD).load_state_dict(stock["d0"])
        logs = []
        for s in range(1, 4):
            model.feed_data(batch)
            model.optimize_parameters(s)
            logs.append(OrderedDict((k, float(v)) for k, v in model.log_dict.items()))
    finally:
        for (modname, n), cls in saved.items():
            import importlib
            setattr(importlib.import_module("models.modules.architectures." + modname), n, cls)
    for s in range(3):
        for k, v in stock32["logs"][s].items():
            e, e16 = abs(logs[s][k] - v) / abs(v), abs(stock["logs"][s][k] - v) / abs(v)
            print("step %d %-14s ref32 % .5e | ref-bf16 rel %.2e | reference SRModel + b200 modules rel %.2e" %
                  (s + 1, k, v, e16, e))
            # step 1 is a pure function of the inputs; later steps follow two Adam updates of G and D
            assert e <= (2e-2 if s == 0 else max(5e-2, 2.0 * e16)), (s + 1, k, logs[s][k], v)
