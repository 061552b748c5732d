"""Per-layer gradient parity report (GPU): engine vs fp32 oracle, next to the reference's own
bf16-autocast-vs-fp32 gap (the like-for-like yardstick).  Not part of the test-suite."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from oracle import esrgan_oracle as O
from ref_harness import seeded_state

def cos(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))
def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))

GOLD = os.path.join(ROOT, "tests", "golden")
which = sys.argv[1] if len(sys.argv) > 1 else "g"
if which == "g":
    from trainner_b200.architectures import RRDBNet_arch
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    net = RRDBNet_arch.RRDBNet(3, 3, 64, nb).cuda()
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    sd = seeded_state(shapes, 21)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 12, 20, generator=g)
    dy = torch.randn(2, 3, 48, 80, generator=torch.Generator().manual_seed(5))
    def oracle(dev, autocast):
        p = OrderedDict((k, v.clone().to(dev).requires_grad_(True)) for k, v in sd.items())
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = O.rrdbnet_forward(p, x.to(dev), nb)
        y.float().backward(dy.to(dev))
        return y.detach().float().cpu(), OrderedDict((k, v.grad.float().cpu()) for k, v in p.items())
    y32, g32 = oracle("cuda", False)
    y16, g16 = oracle("cuda", True)
    y = net(x.cuda()); y.backward(dy.cuda())
    print("fwd rel: engine %.4f  ref-bf16 %.4f" % (rel(y, y32), rel(y16, y32)))
    for k, p in net.named_parameters():
        print("%-40s engine cos %.5f rel %.4f | ref-bf16 cos %.5f rel %.4f" % (k, cos(p.grad, g32[k]), rel(p.grad, g32[k]), cos(g16[k], g32[k]), rel(g16[k], g32[k])))
elif which == "d":
    from trainner_b200.architectures import discriminators
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    net = discriminators.Discriminator_VGG(size, 3, 64).cuda(); net.train()
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    sd = seeded_state(shapes, 31); net.load_state_dict(sd)
    x = torch.rand(N, 3, size, size, generator=torch.Generator().manual_seed(1))
    dy = torch.randn(N, 1, generator=torch.Generator().manual_seed(6))
    def oracle(autocast):
        p = OrderedDict((k, (v.clone().cuda().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v.clone().cuda())) for k, v in sd.items())
        xo = x.cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = O.discriminator_vgg_forward(p, xo, size, training=True)
        y.float().backward(dy.cuda())
        return y.detach().float().cpu(), xo.grad.cpu(), OrderedDict((k, v.grad.float().cpu()) for k, v in p.items() if v.requires_grad)
    y32, dx32, g32 = oracle(False); y16, dx16, g16 = oracle(True)
    xc = x.cuda().requires_grad_(True); y = net(xc); y.backward(dy.cuda())
    print("logits engine", y.flatten().tolist(), "\n fp32", y32.flatten().tolist(), "\n bf16", y16.flatten().tolist())
    print("dx: engine cos %.5f rel %.4f | ref-bf16 cos %.5f rel %.4f" % (cos(xc.grad, dx32), rel(xc.grad, dx32), cos(dx16, dx32), rel(dx16, dx32)))
    for k, p in net.named_parameters():
        print("%-28s engine cos %.5f rel %.4f | ref-bf16 cos %.5f rel %.4f   |g| %.3g" % (k, cos(p.grad, g32[k]), rel(p.grad, g32[k]), cos(g16[k], g32[k]), rel(g16[k], g32[k]), float(g32[k].abs().max())))
