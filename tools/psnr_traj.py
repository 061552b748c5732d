"""PSNR trajectories of the three paths (fp32 oracle, reference-bf16 = oracle under autocast, trainner_b200) on the
learnable synthetic task of tests/test_modules_gpu.py::test_psnr_after_training_matches_reference_paths."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from oracle import esrgan_oracle as O
from trainner_b200.models.sr_model import create_model
nb, hr, bs, lr = 2, 64, 8, float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
def psnr(sr, h, crop=4):
    a = (sr.float().clamp(0, 1) * 255.0).round()[..., crop:-crop, crop:-crop].double()
    b = (h.float().clamp(0, 1) * 255.0).round()[..., crop:-crop, crop:-crop].double()
    return float(20.0 * torch.log10(255.0 / torch.sqrt(((a - b) ** 2).mean())))
torch.manual_seed(0)
opt = {"model": "sr", "scale": 4, "is_train": True, "datasets": {"train": {"crop_size": hr}},
       "network_G": {"type": "esrgan", "nb": nb, "nf": 64, "gaussian": False, "init_scale": 0.3},
       "train": {"pixel_weight": 1.0, "feature_weight": 0, "gan_weight": 0, "lr_G": lr}}
model = create_model(opt)
g_sd = OrderedDict((k, v.detach().clone()) for k, v in model.netG.state_dict().items())
def batch(seed, n=bs):
    g = torch.Generator().manual_seed(seed)
    h = F.interpolate(torch.rand(n, 3, 8, 8, generator=g), size=hr, mode="bicubic", align_corners=False).clamp(0, 1)
    return F.interpolate(h, scale_factor=0.25, mode="area").cuda(), h.cuda()
o32 = O.ESRGANStepOracle(g_sd, nb, pixel_weight=1.0, feature_weight=0, lr=lr, device="cuda")
o16 = O.ESRGANStepOracle(g_sd, nb, pixel_weight=1.0, feature_weight=0, lr=lr, device="cuda")
lv, hv = batch(7, 32)
acc = [0.0, 0.0, 0.0]; cnt = 0
for s in range(1, steps + 1):
    l, h = batch(1000 + s)
    o32.optimize_parameters(l, h)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o16.optimize_parameters(l, h)
    model.feed_data({"LR": l, "HR": h}); model.optimize_parameters(s)
    if s % 50 == 0:
        with torch.no_grad():
            p32 = psnr(o32.netG(lv), hv)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                p16 = psnr(o16.netG(lv).float(), hv)
        model.feed_data({"LR": lv, "HR": hv}); model.test(); pb = psnr(model.fake_H, hv)
        ll = model.get_current_log()
        print("step %4d  fp32 %.2f  ref-bf16 %.2f  b200 %.2f   (train L1: fp32 %.4f bf16 %.4f b200 %.4f)" % (s, p32, p16, pb, o32.log_dict["pix-l1"], o16.log_dict["pix-l1"], ll["pix-l1"]), flush=True)
        if s > steps // 2:
            acc[0] += p32; acc[1] += p16; acc[2] += pb; cnt += 1
print("mean over 2nd half: fp32 %.3f  ref-bf16 %.3f  b200 %.3f" % tuple(a / cnt for a in acc))
