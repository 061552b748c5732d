"""In-kernel timeline of the whole-trunk chain kernel (csrc/rdb_chain.cu) at config-2 size.
    python tools/time_chain.py [nb] [fwd|bwd]
Prints, for a few CTAs, per stage s (first 15 stages) of tile 0: cycles since CTA entry at which the operand
slice was ready (R), the stage's MMAs were all issued (I), complete as seen by the epilogue (C) and the finished
slice was turned around incl. halo exchange (T); plus the event-timed launch duration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trainner_b200 import networks
from trainner_b200.architectures import RRDBNet_arch
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 23
torch.manual_seed(0)
net = RRDBNet_arch.RRDBNet(3, 3, 64, nb).cuda()
networks.init_weights(net, "kaiming", 0.3)
x = torch.rand(16, 3, 64, 64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for _ in range(5):
        net(x)
e1.record()
torch.cuda.synchronize()
print("G forward (eval), %d RRDBs, 16x64x64: %.3f ms per call" % (nb, e0.elapsed_time(e1) / 5))
dbg = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
os.environ["B200_CHAIN_DBG_PTR"] = str(dbg.data_ptr())
os.environ["B200_GRAPHS"] = "0"
from trainner_b200 import runtime
runtime.Plan.use_graphs = False
net2 = RRDBNet_arch.RRDBNet(3, 3, 64, nb).cuda()
net2.load_state_dict(net.state_dict())
with torch.no_grad():
    net2(x)
torch.cuda.synchronize()
d = dbg.view(148, 64).cpu()
for cta in (0, 1, 68, 135, 136):
    t0 = int(d[cta, 0])
    print("cta %3d: total %d cycles" % (cta, int(d[cta, 1]) - t0))
    for s in range(7):
        r, i, c, t, e, l, f, st = (int(d[cta, 2 + 8 * s + k]) - t0 for k in range(8))
        print("   s%-2d R %7d  I +%5d  C +%5d | tmem_ld +%5d  math +%5d  stores +%5d  loop-end(E) +%5d | T(arrive) +%5d" %
              (s, r, i - r, c - i, l - c, f - l, st - f, e - st, t - e))
