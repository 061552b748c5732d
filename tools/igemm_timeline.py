"""In-kernel timeline of conv_igemm256 (tap mode) for one layer: python tools/igemm_timeline.py <shape index>
CTA 0 stamps clock64() when its producer issues a stage, when MMA issuer 0 sees the stage full / has committed it,
and when the epilogue starts / ends a tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["B200_IGEMM_DBG_PTR"] = str(dbg.data_ptr())
import ctypes as C
from trainner_b200 import ops, _lib
from trainner_b200.runtime import make_conv_desc, taps_conv, stream_ptr
SHAPES = [(16, 256, 256, 64, 64), (16, 128, 128, 64, 128), (16, 128, 128, 128, 128), (16, 64, 64, 128, 256),
          (16, 64, 64, 256, 256), (16, 32, 32, 256, 512), (16, 32, 32, 512, 512), (16, 16, 16, 512, 512)]
n, h, w, cin, cout = SHAPES[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
x = (torch.randn(n, h, w, cin, device="cuda") * 0.5).to(torch.bfloat16)
wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.randn(cout, device="cuda") * 0.1
y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
wp = ops.pack_weight(wt, 0)
d = make_conv_desc(n, h, w, cin, 0, cin, h, w, h, w, cout, 0, cout, taps_conv(3, 1), 9, wp.shape[1], wp.shape[2])
P = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.lib.b200_conv_igemm(C.byref(d), P(x), P(wp), P(b), None, None, None, P(y), stream_ptr())
torch.cuda.synchronize()
e = dbg.cpu().tolist()
t0 = e[128]
print("conv %d -> %d @ %dx%dx%d ; cycles since the producer's first issue" % (cin, cout, n, h, w))
print(" it | producer issue | full seen   (+wait) | committed (+issue)")
for it in range(40):
    pi, f, c = e[128 + it] - t0, e[2 * it] - t0, e[2 * it + 1] - t0
    print("%3d | %8d       | %8d  (%+6d) | %8d (%+5d)" % (it, pi, f, f - pi, c, c - f))
for t in range(6):
    print("epilogue tile %d: start %d end %d (%d)" % (t, e[256 + 2 * t] - t0, e[257 + 2 * t] - t0, e[257 + 2 * t] - e[256 + 2 * t]))
