"""Generate the committed golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    cd /root/repo && python tests/golden/make_golden.py

Weights are produced by ref_harness.seeded_state (a pure function of key order, shapes and a
seed) and loaded into the reference modules with load_state_dict, so the tests can regenerate the
identical weights without the reference and compare against the stored reference OUTPUTS.

Fixtures (tests/golden/*.pt, fp32 CPU, torch 2.11):
  modules.pt   RRDBNet fwd (upconv + pixelshuffle), Discriminator_VGG fwd (train mode, BN stats
               after one call), FeatureExtractor conv5_4 for fixed seeded inputs
  config1.pt   BASELINE config 1: nb=1, 32x32 -> 128x128, L1 only, 3 optimize_parameters steps
  mini2.pt     shrunk config 2: nb=2, 16x16 -> 64x64, pix-l1 + fea-vgg19-l1 + vanilla RaGAN,
               Discriminator_VGG(size=64), 2 steps: log_dict, SR, param/BN checksums
"""
import os
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as R  # noqa: E402

R._install_shims("/tmp/_golden_torch_home")
import torch  # noqa: E402

from oracle import esrgan_oracle as O  # noqa: E402

torch.set_num_threads(8)


def shapes_of(module):
    return OrderedDict((k, tuple(v.shape)) for k, v in module.state_dict().items())


def checksums(sd):
    out = OrderedDict()
    for k, v in sd.items():
        v = v.detach().double()
        out[k] = (float(v.sum()), float(v.abs().sum()))
    return out


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def gen_modules():
    from models.modules.architectures import RRDBNet_arch, discriminators, perceptual

    fx = OrderedDict()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 12, 20, generator=g)
    for mode in ("upconv", "pixelshuffle"):
        net = RRDBNet_arch.RRDBNet(3, 3, 64, 2, upsample_mode=mode, gaussian_noise=False)
        shp = shapes_of(net)
        sd = R.seeded_state(shp, seed=21, scale=None)
        # damp so that a 2-block net has O(1) outputs
        net.load_state_dict(sd)
        with torch.no_grad():
            y = net(x)
            yo = O.rrdbnet_forward(sd, x, 2, mode)
        print("RRDBNet", mode, "ref-vs-oracle rel", rel(yo, y), "out std", float(y.std()))
        assert rel(yo, y) < 1e-6
        fx["rrdb_%s" % mode] = {"shapes": shp, "seed": 21, "x": x, "y": y}

    for size in (32, 64):
        net = discriminators.Discriminator_VGG(size, 3, 64)
        shp = shapes_of(net)
        sd = R.seeded_state(shp, seed=31)
        net.load_state_dict(sd)
        net.train()
        xd = torch.rand(4, 3, size, size, generator=g)
        with torch.no_grad():
            y = net(xd)
        sd_after = OrderedDict((k, v.clone()) for k, v in net.state_dict().items())
        sdo = OrderedDict((k, v.clone()) for k, v in sd.items())
        with torch.no_grad():
            yo = O.discriminator_vgg_forward(sdo, xd, size, training=True)
        print("D size", size, "ref-vs-oracle rel", rel(yo, y), float(y.abs().mean()))
        assert rel(yo, y) < 1e-5
        for k in sd_after:
            if "running" in k:
                assert rel(sdo[k], sd_after[k]) < 1e-5, k
        net.eval()
        with torch.no_grad():
            ye = net(xd)
        bn = OrderedDict((k, v) for k, v in sd_after.items() if "running" in k or "tracked" in k)
        fx["disc_%d" % size] = {"shapes": shp, "seed": 31, "x": xd, "y_train": y, "y_eval": ye,
                                "bn_after": bn}

    netF = perceptual.FeatureExtractor(listen_list=["conv5_4"], net="vgg19", use_input_norm=True)
    shp = shapes_of(netF)
    import torchvision
    tv_shapes = OrderedDict((k, tuple(v.shape)) for k, v in
                            torchvision.models.vgg19(weights=None).state_dict().items())
    tv_sd = R.seeded_state(tv_shapes, 7)
    fsd = O.torchvision_vgg_to_feature_net(tv_sd)
    for k in fsd:
        assert torch.equal(fsd[k], netF.state_dict()[k]), k
    xf = torch.rand(2, 3, 48, 64, generator=g)
    with torch.no_grad():
        y = netF(xf)["conv5_4"]
        yo = O.vgg19_features(fsd, xf)["conv5_4"]
    print("VGG ref-vs-oracle rel", rel(yo, y), float(y.std()))
    assert rel(yo, y) < 1e-6
    fx["vgg19"] = {"tv_seed": 7, "x": xf, "conv5_4": y}
    torch.save(fx, os.path.join(HERE, "modules.pt"))


def run_steps(nb, hr, bs, steps, use_gan, use_fea, pixel_weight, seed_w, seed_x):
    model, opt = R.create_reference_model(nb=nb, hr_size=hr, use_gan=use_gan, use_fea=use_fea,
                                          pixel_weight=pixel_weight)
    netG = model.netG
    gshp = shapes_of(netG)
    # reference init scale (kaiming x 0.1) gives ~0 outputs (SURVEY 8d trap); use fan-in scaled x0.5
    g_sd = R.seeded_state(gshp, seed_w)
    g_sd = OrderedDict((k, v * (0.5 if v.dim() > 1 else 1.0)) for k, v in g_sd.items())
    netG.load_state_dict(g_sd)
    d_sd = None
    dshp = None
    if use_gan:
        dshp = shapes_of(model.netD)
        d_sd = R.seeded_state(dshp, seed_w + 1)
        model.netD.load_state_dict(d_sd)
    vgg_sd = None
    if use_fea:
        netF = [l for l in model.generatorlosses.loss_list if "fea" in l["name"]][0]["function"].network
        vgg_sd = OrderedDict((k, v.clone()) for k, v in netF.state_dict().items()
                             if k.startswith("feature_net"))
    orc = O.ESRGANStepOracle(g_sd, nb, d_sd, hr if use_gan else None, vgg_sd,
                             pixel_weight=pixel_weight, feature_weight=1.0 if use_fea else 0,
                             gan_weight=5e-3)
    g = torch.Generator().manual_seed(seed_x)
    logs = []
    batches = []
    for s in range(1, steps + 1):
        lr_img = torch.rand(bs, 3, hr // 4, hr // 4, generator=g)
        hr_img = torch.rand(bs, 3, hr, hr, generator=g)
        batches.append((lr_img, hr_img))
        model.feed_data({"LR": lr_img, "HR": hr_img})
        model.optimize_parameters(s)
        ref_log = OrderedDict(model.log_dict)
        olog = OrderedDict(orc.optimize_parameters(lr_img, hr_img))
        logs.append(ref_log)
        for k in ref_log:
            e = abs(ref_log[k] - olog[k]) / (abs(ref_log[k]) + 1e-12)
            print("  step", s, k, ref_log[k], olog[k], "rel", e)
            assert e < 2e-4, (k, ref_log[k], olog[k])
    lr_t = torch.rand(bs, 3, hr // 4, hr // 4, generator=g)
    model.feed_data({"LR": lr_t, "HR": torch.zeros(bs, 3, hr, hr)})
    model.test()
    sr = model.fake_H.detach().clone()
    with torch.no_grad():
        sro = orc.netG(lr_t)
    print("  SR rel oracle-vs-ref", rel(sro, sr), "std", float(sr.std()))
    assert rel(sro, sr) < 1e-4
    out = {"nb": nb, "hr": hr, "bs": bs, "g_shapes": gshp, "g_seed": seed_w, "g_gain": 0.5,
           "d_shapes": dshp, "d_seed": seed_w + 1, "vgg_tv_seed": 7 if use_fea else None,
           "pixel_weight": pixel_weight, "batches": batches, "logs": logs, "lr_test": lr_t,
           "sr_test": sr, "g_after": checksums(netG.state_dict())}
    for k, v in netG.state_dict().items():
        assert rel(orc.g[k].detach(), v) < 1e-4, k
    if use_gan:
        out["d_after"] = checksums(model.netD.state_dict())
        out["d_bn_after"] = OrderedDict((k, v.clone()) for k, v in model.netD.state_dict().items()
                                        if "running" in k or "tracked" in k)
        for k, v in model.netD.state_dict().items():
            if v.is_floating_point():
                assert rel(orc.d[k].detach(), v) < 2e-4, (k, rel(orc.d[k].detach(), v))
            else:
                assert int(orc.d[k]) == int(v), k
    return out


if __name__ == "__main__":
    gen_modules()
    print("config1")
    c1 = run_steps(nb=1, hr=128, bs=1, steps=3, use_gan=False, use_fea=False, pixel_weight=1.0,
                   seed_w=41, seed_x=51)
    torch.save(c1, os.path.join(HERE, "config1.pt"))
    print("mini2")
    m2 = run_steps(nb=2, hr=64, bs=2, steps=2, use_gan=True, use_fea=True, pixel_weight=1e-2,
                   seed_w=61, seed_x=71)
    torch.save(m2, os.path.join(HERE, "mini2.pt"))
    for f in ("modules.pt", "config1.pt", "mini2.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
