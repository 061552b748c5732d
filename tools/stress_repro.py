"""Stress the bit-reproducibility claim: N runs of 8 identical GAN steps (the configuration of
tests/test_modules_gpu.py::test_training_steps_are_bit_reproducible), every run compared with the first.
    python tools/stress_repro.py [runs]
Prints, for a run that differs, the first step whose losses differ and the parameter tensors that differ."""
import os, sys, tempfile
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from trainner_b200.models.sr_model import create_model

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tmp = tempfile.mkdtemp()
vgg_path = os.path.join(tmp, "vgg19.pth")
torch.manual_seed(5)
torch.save(torchvision.models.vgg19(weights=None).state_dict(), vgg_path)
opt = {"model": "sr", "scale": 4, "is_train": True, "datasets": {"train": {"crop_size": 64}},
       "network_G": {"type": "esrgan", "nb": 3, "nf": 64, "gaussian": False, "init_scale": 0.3},
       "network_D": {"type": "discriminator_vgg"},
       "train": {"pixel_weight": 1e-2, "feature_weight": 1.0, "gan_weight": 5e-3, "gan_type": "vanilla",
                 "lr_G": 1e-4, "lr_D": 1e-4, "perceptual_opt": {"pretrained_path": vgg_path}}}
first, init, bad = None, None, 0
POISON = os.environ.get("STRESS_POISON") == "1"
for r in range(runs):
    if POISON:   # stale contents of recycled allocator blocks become NaN: a read of uninitialised memory shows up as NaN
        junk = [torch.full((64 << 20,), float("nan"), device="cuda") for _ in range(8)]
        del junk
    torch.manual_seed(0)
    model = create_model(opt)
    if init is None:
        init = (OrderedDict((k, v.clone()) for k, v in model.netG.state_dict().items()),
                OrderedDict((k, v.clone()) for k, v in model.netD.state_dict().items()))
    model.netG.load_state_dict(init[0])
    model.netD.load_state_dict(init[1])
    logs = []
    trace = []   # (step, what, name, checksum) in execution order
    if os.environ.get("STRESS_TRACE") == "1":
        orig_step = model.optimizer_step

        def traced_step(step, optimizer, flag, _orig=orig_step, _m=model):
            net = _m.netG if flag == "G" else _m.netD
            for n_, p_ in net.named_parameters():
                if p_.grad is not None:
                    g64 = p_.grad.double()
                    trace.append((step, "grad" + flag, n_, float(g64.sum()), float(g64.abs().sum())))
            _orig(step, optimizer, flag)
            _m.synchronize()
            for n_, p_ in net.named_parameters():
                trace.append((step, "param" + flag, n_, float(p_.double().sum()), float(p_.double().abs().sum())))
            if flag == "D":
                for n_, b_ in net.named_buffers():
                    trace.append((step, "buf" + flag, n_, float(b_.double().sum()), 0.0))

        model.optimizer_step = traced_step
    for s in range(1, 9):
        g = torch.Generator().manual_seed(100 + s)
        model.feed_data({"LR": torch.rand(6, 3, 16, 16, generator=g), "HR": torch.rand(6, 3, 64, 64, generator=g)})
        model.optimize_parameters(s)
        logs.append(model.get_current_log())
    model.synchronize()
    if POISON:
        nan_logs = [(i + 1, k) for i, l in enumerate(logs) for k, v in l.items() if v != v]
        nan_p = [k for k, v in list(model.netG.state_dict().items()) + list(model.netD.state_dict().items()) if torch.isnan(v.float()).any()]
        if nan_logs or nan_p:
            print("run %d: NaN in logs %s, params %s" % (r, nan_logs[:6], nan_p[:6]), flush=True)
    cur = (logs, OrderedDict((k, v.clone()) for k, v in model.netG.state_dict().items()),
           OrderedDict((k, v.clone()) for k, v in model.netD.state_dict().items()), trace)
    if first is None:
        first = cur
        continue
    if trace:
        d = [(a, b) for a, b in zip(first[3], trace) if a != b]
        if d:
            print("run %d trace: first differing entries:" % r)
            for a, b in d[:6]:
                print("    ", a, "|", b[3:], flush=True)
    step = next((i + 1 for i, (a, b) in enumerate(zip(first[0], cur[0])) if a != b), None)
    dg = [k for k, v in first[1].items() if not torch.equal(v, cur[1][k])]
    dd = [k for k, v in first[2].items() if not torch.equal(v, cur[2][k])]
    if step or dg or dd:
        bad += 1
        keys = [k for k in first[0][step - 1] if first[0][step - 1][k] != cur[0][step - 1][k]] if step else []
        print("run %d DIFFERS: first differing step %s (%s); %d G tensors (%s...), %d D tensors (%s...)" %
              (r, step, keys, len(dg), dg[:3], len(dd), dd[:3]), flush=True)
print("%d of %d runs differ from the first" % (bad, runs - 1))
