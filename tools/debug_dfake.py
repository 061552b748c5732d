"""Where does the D_fake gap of the full-size parity test come from?  (GPU, needs baseline/_ref)
D_ref32(SR_ref32) vs D_ref32(SR_b200) [input sensitivity], D_b200(SR_ref32) [our D at identical input],
D_refbf16(SR_ref32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from baseline import reference_arm as RA
import test_reference_parity_gpu as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
batch = T._batch()
model, _ = RA.create_reference_model(torch_home=T.TORCH_HOME, precision="fp32", nb=T.NB, hr_size=T.HR, use_gan=True,
                                     use_fea=True, pixel_weight=1e-2, gpu=True, batch_size=T.BS, init_scale=0.3)
G, D = RA.unwrap(model.netG), RA.unwrap(model.netD)
g_sd = {k: v.clone() for k, v in G.state_dict().items()}
d_sd = {k: v.clone() for k, v in D.state_dict().items()}
from trainner_b200.architectures import RRDBNet_arch, discriminators
Gb = RRDBNet_arch.RRDBNet(3, 3, 64, T.NB).cuda(); Gb.load_state_dict(g_sd)
Db = discriminators.Discriminator_VGG(T.HR, 3, 64).cuda(); Db.load_state_dict(d_sd); Db.train()
D.train()
with torch.no_grad():
    sr32 = G(batch["LR"])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        sr16 = G(batch["LR"]).float()
    srb = Gb(batch["LR"])
    print("SR std %.4e  rel: bf16 %.3e  b200 %.3e" % (float(sr32.std()), T.rel(sr16, sr32), T.rel(srb, sr32)))
    def dref(x, ac=False):
        D.load_state_dict(d_sd)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            return D(x).float().flatten()
    def db(x):
        Db.load_state_dict(d_sd)
        return Db(x).float().flatten()
    a = dref(sr32)
    rows = {"Dref32(SRb200)": dref(srb), "Dref32(SRbf16)": dref(sr16), "Dbf16(SR32)": dref(sr32, True), "Db200(SR32)": db(sr32),
            "Db200(SRb200)": db(srb), "Dbf16(SRbf16)": dref(sr16, True)}
    print("logits ref32:", [round(float(v), 6) for v in a])
    print("mean %.6e std %.3e" % (float(a.mean()), float(a.std())))
    for k, v in rows.items():
        print("%-16s mean %.6e  dmean %.3e  rms diff %.3e" % (k, float(v.mean()), float(v.mean() - a.mean()), float((v - a).pow(2).mean().sqrt())))
    hr = batch["HR"]
    a = dref(hr)
    print("REAL: mean %.6e std %.3e | Dbf16 dmean %.3e | Db200 dmean %.3e" % (float(a.mean()), float(a.std()), float(dref(hr, True).mean() - a.mean()), float(db(hr).mean() - a.mean())))
    # per-layer input statistics of the fake batch
    print("fake batch: per-image mean/std", [(round(float(s.mean()), 4), round(float(s.std()), 4)) for s in sr32[:4]])
