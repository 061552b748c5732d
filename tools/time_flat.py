import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trainner_b200 import ops
torch.manual_seed(0)
def bench(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N, H, W = 16, 64, 64
for (cin, cout, dgrad) in [(64, 32, False), (160, 32, False), (192, 64, False), (32, 160, True), (64, 192, True)]:
    x = torch.randn(N, H, W, 192, device="cuda").to(torch.bfloat16)
    xf = ops.to_flat(x)
    w = torch.randn((cin, cout, 3, 3) if dgrad else (cout, cin, 3, 3), device="cuda") * 0.05
    t0 = time.time()
    out = torch.zeros(N, H + 2, W + 2, 192, dtype=torch.bfloat16, device="cuda")
    ms = bench(lambda: ops.conv3x3_flat(xf, w, None, dgrad=dgrad, out=out, accumulate=dgrad))
    fl = 2.0 * N * H * W * cin * cout * 9
    print("flat  cin%d cout%d dgrad=%d: %.3f ms (incl. weight pack)  %.0f TF/s   wall %.1fs" % (cin, cout, dgrad, ms, fl / ms / 1e9, time.time() - t0), flush=True)
    xd = x[..., :cin].contiguous() if not dgrad else x[..., :cin].contiguous()
    if not dgrad:
        ms = bench(lambda: ops.conv2d(x, w, None, cin_off=0, cin=cin))
        print("igemm cin%d cout%d: %.3f ms  %.0f TF/s" % (cin, cout, ms, fl / ms / 1e9), flush=True)
