"""Launch a few trunk-shaped kernels back to back (pre-packed weights, static buffers) for ncu / timing."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trainner_b200 import ops, _lib
from trainner_b200._lib import lib
from trainner_b200.runtime import make_flat_desc, taps_conv, taps_dgrad_s1, stream_ptr
N, H, W, C_ = 16, 64, 64, 192
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
x = ops.to_flat(torch.randn(N, H, W, C_, device="cuda").to(torch.bfloat16))
g = ops.to_flat(torch.randn(N, H, W, C_, device="cuda").to(torch.bfloat16))
cases = {"fwd64_32": (64, 32, False), "fwd160_32": (160, 32, False), "fwd192_64": (192, 64, False),
         "dg32_160": (32, 160, True), "dg32_64": (32, 64, True), "dg64_192": (64, 192, True)}
for name, (cin, cout, dgrad) in cases.items():
    if which != "all" and which != name:
        continue
    w = torch.randn((cin, cout, 3, 3) if dgrad else (cout, cin, 3, 3), device="cuda") * 0.05
    wp = ops.pack_weight(w, 1 if dgrad else 0)
    if dgrad:
        d = make_flat_desc(N, H, W, C_, 64, cin, C_, 0, cout, taps_dgrad_s1(3, 1), 9, wp.shape[1], wp.shape[2],
                           accumulate=1, mask_c=C_, mask_lo=cout - 32, mask_hi=cout, mask_slope=0.2)
        args = (C.byref(d), g.data_ptr(), None, wp.data_ptr(), None, None, None, x.data_ptr(), g.data_ptr())
    else:
        d = make_flat_desc(N, H, W, C_, 0, cin, C_, 160, cout, taps_conv(3, 1), 9, wp.shape[1], wp.shape[2], act=1, slope=0.2)
        args = (C.byref(d), x.data_ptr(), None, wp.data_ptr(), None, None, None, None, x.data_ptr())
    s = stream_ptr()
    for _ in range(3):
        assert lib.b200_conv3x3_flat(*args, s) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.b200_conv3x3_flat(*args, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * N * H * W * cin * cout * 9
    print("%-10s %.1f us/launch  %.0f TF/s" % (name, ms * 1e3, fl / ms / 1e9), flush=True)
    if os.environ.get("TIMELINE"):
        dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
        os.environ["B200_FLAT_DBG_PTR"] = str(dbg.data_ptr())
        lib.b200_conv3x3_flat(*args, s); torch.cuda.synchronize()
        del os.environ["B200_FLAT_DBG_PTR"]
        d_ = dbg.view(148, 16).cpu()
        names = ["entry", "setup_done", "mma_first_A", "mma_tile0_issued", "mma_tile1_issued", "epi_tile0_tfull", "epi_tile0_done", "epi_tile1_tfull", "epi_tile1_done", "exit"]
        for cta in (0, 1, 100, 140):
            t0 = int(d_[cta, 0])
            print("  cta %3d: " % cta + "  ".join("%s=%d" % (n, int(d_[cta, i]) - t0 if int(d_[cta, i]) else -1) for i, n in enumerate(names)))
