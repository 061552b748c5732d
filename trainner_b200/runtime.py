"""Host runtime shared by the three network engines: static launch plans, weight packing,
flat gradient storage, workspace contexts.

Design (B200-first): a network forward/backward is a *static plan* -- a list of
C-ABI calls whose device pointers are resolved once, when the plan is built for an input shape.
Running a plan is a tight loop of ctypes calls on the current CUDA stream (no tensor objects, no
dispatcher), so the host stays far ahead of the GPU and the whole step is CUDA-graph capturable.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ConvDesc, FlatDesc, PackCatEntry, PackEntry, WgradDesc, lib

LRELU_SLOPE = 0.2


def require_device(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(
            "trainner_b200.%s runs only on a CUDA (sm_100a) device; got a %s tensor. "
            "There is no CPU fallback -- use oracle/ for CPU checks." % (what, "CPU" if isinstance(t, torch.Tensor) else type(t)))
    if not lib.b200_device_ok():
        raise RuntimeError("trainner_b200.%s: current CUDA device is not sm_100 (B200)" % what)


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def sm_count_hint():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def roundup(v, m):
    return (v + m - 1) // m * m


class Plan:
    """A list of (C function, argument tuple); the stream is appended at run time."""

    detail_sink = None  # set to a list to collect (tag, info, ms, flops) per call from run_timed()
    use_graphs = os.environ.get("B200_GRAPHS", "1") != "0"

    def __init__(self):
        self.calls = []
        self.meta = []   # (kernel tag, algorithmic FLOPs) per call, for the roofline report
        self._keep = []
        self._graph = None
        self._runs = 0

    def add(self, fn, *args, flops=0.0, tag=None, info=""):
        self.calls.append((fn, args))
        self.meta.append((tag or fn.__name__, float(flops), info))

    def keep(self, obj):
        self._keep.append(obj)
        return obj

    def _run_eager(self):
        s = stream_ptr()
        for fn, args in self.calls:
            if fn(*args, s):
                raise RuntimeError("trainner_b200 kernel call %s failed: %s" %
                                   (fn.__name__, lib.b200_last_error().decode()))

    def run(self):
        """Eager for the first two runs (lazy one-time initialisation inside the library), then the
        plan is captured once into a CUDA graph and replayed: every pointer in a plan is static, so
        the graph stays valid for the life of the context.  B200_GRAPHS=0 disables capture."""
        if not Plan.use_graphs or self._graph is False or len(self.calls) < 8:
            return self._run_eager()
        if self._graph is None:
            self._runs += 1
            if self._runs <= 2:
                return self._run_eager()
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_eager()
                self._graph = g
            except Exception as exc:   # not capturable on this setup: stay eager, but say so (a silent perf cliff otherwise)
                self._graph = False
                import warnings
                warnings.warn("trainner_b200: CUDA-graph capture of a %d-call plan failed (%s: %s); the plan stays eager"
                              % (len(self.calls), type(exc).__name__, exc), RuntimeWarning)
                torch.cuda.synchronize()
                return self._run_eager()
        self._graph.replay()

    def run_timed(self):
        """Run with a CUDA event pair around every call (on the launching stream); returns
        [(tag, milliseconds, flops)].  Used by bench.py's roofline leg, never in the timed step."""
        s = stream_ptr()
        evs = []
        for fn, args in self.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if fn(*args, s):
                raise RuntimeError("trainner_b200 kernel call %s failed: %s" %
                                   (fn.__name__, lib.b200_last_error().decode()))
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        rows = [(m[0], e0.elapsed_time(e1), m[1]) for m, (e0, e1) in zip(self.meta, evs)]
        if Plan.detail_sink is not None:
            Plan.detail_sink.extend((m[0], m[2], e0.elapsed_time(e1), m[1]) for m, (e0, e1) in zip(self.meta, evs))
        return rows

    def __len__(self):
        return len(self.calls)


def P(t):
    """device pointer (int) of a tensor or None"""
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
class ConvLayer:
    """A conv parameter holder (nn.Conv2d) plus its packed bf16 tensor-core weights."""

    def __init__(self, conv, name=""):
        self.conv = conv
        self.name = name
        self.cout, self.cin, self.kh, self.kw = conv.weight.shape
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]
        self.taps = self.kh * self.kw
        self.w_fwd = None   # bf16 [taps][cout_pad16][cin_pad64]
        self.w_dgr = None   # bf16 [taps][cin_pad16][cout_pad64]
        self.need_dgrad = True

    @property
    def weight(self):
        return self.conv.weight

    @property
    def bias(self):
        return self.conv.bias

    def alloc(self, device):
        self.fwd_rows, self.fwd_cols = roundup(self.cout, 16), roundup(self.cin, 64)
        self.dgr_rows, self.dgr_cols = roundup(self.cin, 16), roundup(self.cout, 64)
        self.w_fwd = torch.empty(self.taps, self.fwd_rows, self.fwd_cols, dtype=torch.bfloat16, device=device)
        if self.need_dgrad:
            self.w_dgr = torch.empty(self.taps, self.dgr_rows, self.dgr_cols, dtype=torch.bfloat16, device=device)

    def pack_entries(self):
        w = self.conv.weight
        out = [PackEntry(w.data_ptr(), self.w_fwd.data_ptr(), self.cout, self.cin, self.taps,
                         self.fwd_rows, self.fwd_cols, 0, 0, 0)]
        if self.need_dgrad:
            out.append(PackEntry(w.data_ptr(), self.w_dgr.data_ptr(), self.cout, self.cin, self.taps,
                                 self.dgr_rows, self.dgr_cols, 1, 0, 0))
        return out


class WeightPacker:
    """Repacks every tensor-core conv weight of a network with ONE kernel launch, only when a
    parameter changed (tracked through Tensor._version)."""

    def __init__(self, layers, device, cat_entries=None):
        self.layers = layers
        self.cat_count = 0
        if cat_entries:
            self.cat_count = len(cat_entries)
            arr = (PackCatEntry * self.cat_count)(*cat_entries)
            self.cat_table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
            self.cat_max = max(e.taps * e.n_rows * e.cout for e in cat_entries)
        entries = []
        for l in layers:
            l.alloc(device)
            entries += l.pack_entries()
        self.count = len(entries)
        self.max_elems = 0
        if self.count:
            arr = (PackEntry * self.count)(*entries)
            raw = bytes(arr)
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
            self.max_elems = max(e.taps * e.rows_pad * e.cols_pad for e in entries)
        self.version = None
        self.pack_count = 0
        self.dirty = True
        self.explicit = False
        self.was_trainable = False
        self.ptr_sig = tuple(l.weight.data_ptr() for l in layers)

    def stale_pointers(self):
        return self.ptr_sig != tuple(l.weight.data_ptr() for l in self.layers)

    def mark_dirty(self):
        self.dirty = True

    def ensure(self):
        """Repack before a forward.  Parameter updates are NOT reliably visible through
        Tensor._version (fused optimizers mutate through tensor lists), so by default every forward
        of a trainable network repacks (one ~20 us launch).  A caller that owns the optimizer can
        switch to explicit mode (`explicit = True`) and call mark_dirty() after each step; frozen
        networks repack only when their tensors are replaced or versions change."""
        if not self.count:
            return
        ver = sum(l.weight._version for l in self.layers)
        trainable = any(l.weight.requires_grad for l in self.layers) or self.was_trainable
        self.was_trainable = self.was_trainable or trainable
        need = self.dirty or ver != self.version or (trainable and not self.explicit)
        if need:
            _lib.check(lib.b200_pack_weights(self.table.data_ptr(), self.count, self.max_elems,
                                             stream_ptr()), "pack_weights")
            if self.cat_count:
                _lib.check(lib.b200_pack_cat(self.cat_table.data_ptr(), self.cat_count, self.cat_max,
                                             stream_ptr()), "pack_cat")
            self.version = ver
            self.dirty = False
            self.pack_count += 1


class FlatGrads:
    """All gradients of a network live in one flat fp32 buffer (param.grad are views): wgrad
    kernels accumulate straight into it and the data-parallel all-reduce is one NCCL call."""

    def __init__(self, params, device):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self._by_id = {id(p): v for p, v in zip(self.params, self.views)}

    def view(self, p):
        return self._by_id[id(p)]

    def attach(self):
        """Make param.grad point at the flat views for every param that requires grad.  If the
        optimizer dropped the grads (zero_grad(set_to_none=True)) the flat buffer is re-zeroed.
        Returns the list of params that take gradients this pass."""
        active = [p for p in self.params if p.requires_grad]
        fresh = [p for p in active if p.grad is None]
        if fresh:
            if len(fresh) == len(active) == len(self.params):
                self.flat.zero_()
            else:
                for p in fresh:
                    self.view(p).zero_()
            for p in fresh:
                p.grad = self.view(p)
        for p in active:
            if p.grad is not None and p.grad.data_ptr() != self.view(p).data_ptr():
                # foreign grad tensor (e.g. produced by plain autograd): fold it in and adopt ours
                v = self.view(p)
                v.copy_(p.grad)
                p.grad = v
        return active


class ContextPool:
    """Workspace contexts (activation buffers + bound plans).  A forward that needs a backward
    leases a context until its backward ran (or the autograd node died)."""

    def __init__(self, factory):
        self.factory = factory
        self.free = []
        self.all = []

    def acquire(self):
        if self.free:
            return self.free.pop()
        ctx = self.factory()
        self.all.append(ctx)
        return ctx

    def release(self, ctx):
        self.free.append(ctx)

    def in_use(self):
        return len(self.all) - len(self.free)


def pool_for(pools, key, factory, max_shapes=4):
    """Per-shape workspace pools with an LRU cap: validating over images of many sizes would otherwise
    keep one full activation (+ gradient) workspace per distinct shape forever.  Pools with a leased
    context are never evicted."""
    pool = pools.get(key)
    if pool is not None:
        pools[key] = pools.pop(key)    # most recently used last
        return pool
    while len(pools) >= max_shapes:
        victim = next((k for k, p in pools.items() if p.in_use() == 0), None)
        if victim is None:
            break
        del pools[victim]
    pool = pools[key] = ContextPool(factory)
    return pool


class Lease:
    def __init__(self, pool, ctx):
        self.pool, self.ctx = pool, ctx

    def release(self):
        if self.ctx is not None:
            self.pool.release(self.ctx)
            self.ctx = None

    def __del__(self):
        self.release()


# ------------------------------------------------------------------------------------------------
# descriptor builders

def taps_conv(k, pad):
    """forward taps of a k x k conv: input offset (ky - pad, kx - pad), weight index ky*k+kx"""
    return [(ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k)]


def taps_dgrad_s1(k, pad):
    """dgrad of a stride-1 conv: dX[y] = sum_k dY[y + pad - ky] * W[ky]"""
    return [(pad - ky, pad - kx, ky * k + kx) for ky in range(k) for kx in range(k)]


def taps_dgrad_s2_k4(py, px):
    """dgrad of the 4x4 stride-2 pad-1 conv restricted to output parity (py, px):
    y = 2a+py gets dY[a + d] * W[ky] for (ky, d) in the table below."""
    tab = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}
    return [(dy, dx, ky * 4 + kx) for (ky, dy) in tab[py] for (kx, dx) in tab[px]]


def make_conv_desc(n, h_in, w_in, cx, cin_off, cin, h_out, w_out, h_buf, w_buf, cy, cout_off, cout,
                   taps, w_taps, w_rows, w_cols, in_stride=1, in_off=(0, 0), out_mul=(1, 1),
                   out_off=(0, 0), upsample=0, alpha=1.0, act=0, slope=0.0, beta1=0.0, beta2=0.0,
                   res_nch=0, res1_c=0, res1_coff=0, res2_c=0, res2_coff=0, accumulate=0, mask_c=0,
                   mask_coff=0, mask_lo=0, mask_hi=0, mask_slope=0.0, parity_classes=0):
    d = ConvDesc()
    d.n, d.h_in, d.w_in, d.cx, d.cin_off, d.cin = n, h_in, w_in, cx, cin_off, cin
    d.h_out, d.w_out, d.h_buf, d.w_buf, d.cy, d.cout_off, d.cout = h_out, w_out, h_buf, w_buf, cy, cout_off, cout
    d.parity_classes = parity_classes
    d.ntaps = len(taps) // (parity_classes or 1)   # `taps` lists the classes' taps one class after the other
    for i, (dy, dx, wi) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_w[i] = dy, dx, wi
    d.in_stride, d.in_off_y, d.in_off_x = in_stride, in_off[0], in_off[1]
    d.out_mul_y, d.out_mul_x, d.out_off_y, d.out_off_x = out_mul[0], out_mul[1], out_off[0], out_off[1]
    d.upsample2x = upsample
    d.w_taps, d.w_cout_pad, d.w_cin_pad = w_taps, w_rows, w_cols
    d.alpha, d.act, d.slope = alpha, act, slope
    d.beta1, d.beta2, d.res_nch = beta1, beta2, res_nch
    d.res1_c, d.res1_coff, d.res2_c, d.res2_coff = res1_c, res1_coff, res2_c, res2_coff
    d.accumulate = accumulate
    d.mask_c, d.mask_coff, d.mask_lo, d.mask_hi, d.mask_slope = mask_c, mask_coff, mask_lo, mask_hi, mask_slope
    return d


def add_igemm(plan, desc, x, w, bias=None, res1=None, res2=None, mask=None, y=None):
    plan.keep(desc)
    ncls = desc.parity_classes or 1
    flops = 2.0 * desc.n * desc.h_out * desc.w_out * desc.cout * desc.cin * desc.ntaps * ncls
    info = "cin%d cout%d %dx%dx%d taps%d s%d%s%s%s" % (desc.cin, desc.cout, desc.n, desc.h_out, desc.w_out, desc.ntaps,
                                                     desc.in_stride, " acc" if desc.accumulate else "",
                                                     " up" if desc.upsample2x else "", " x4cls" if ncls == 4 else "")
    plan.add(lib.b200_conv_igemm, C.byref(desc), P(x), P(w), P(bias), P(res1), P(res2), P(mask), P(y),
             flops=flops, tag="conv_igemm", info=info)


def igemm_stat_rows(desc):
    """Rows of BatchNorm partial statistics the conv's epilogue would write (0: this shape is not served by the
    statistics epilogue -- run b200_bn_stats_finalize over the output instead)."""
    r = lib.b200_conv_igemm_stat_rows(C.byref(desc))
    if r < 0:
        raise RuntimeError("trainner_b200 conv_igemm_stat_rows failed: %s" % lib.b200_last_error().decode())
    return r


def add_igemm_stats(plan, desc, x, w, bias, y, stat_part):
    """conv + per-tile BatchNorm partial sums of its bf16 output in one launch (csrc/conv_igemm.cu, EPI = 3)."""
    plan.keep(desc)
    flops = 2.0 * desc.n * desc.h_out * desc.w_out * desc.cout * desc.cin * desc.ntaps
    info = "cin%d cout%d %dx%dx%d taps%d s%d +stats" % (desc.cin, desc.cout, desc.n, desc.h_out, desc.w_out, desc.ntaps,
                                                         desc.in_stride)
    plan.add(lib.b200_conv_igemm_stats, C.byref(desc), P(x), P(w), P(bias), P(y), P(stat_part),
             flops=flops, tag="conv_igemm", info=info)


def add_wgrad(plan, n, h_in, w_in, cx, x_coff, cin, h_out, w_out, cdy, dy_coff, cout, k, stride, pad,
              scale, x, dy, dw, db):
    d = WgradDesc(n, h_in, w_in, cx, x_coff, cin, h_out, w_out, cdy, dy_coff, cout, k, k, stride, pad,
                  scale)
    plan.keep(d)
    plan.add(lib.b200_conv_wgrad, C.byref(d), P(x), P(dy), P(dw), P(db),
             flops=2.0 * n * h_out * w_out * cout * cin * k * k, tag="conv_wgrad",
             info="cin%d cout%d %dx%dx%d k%d s%d" % (cin, cout, n, h_out, w_out, k, stride))


def make_flat_desc(n, h, w, cx, cin_off, cin, cy, cout_off, cout, taps, w_taps, w_rows, w_cols, out_mode=0,
                   cx2=0, cin2_off=0, cin2=0,
                   alpha=1.0, act=0, slope=0.0, beta1=0.0, beta2=0.0, res_nch=0, res1_c=0, res1_coff=0,
                   res2_c=0, res2_coff=0, accumulate=0, mask_c=0, mask_coff=0, mask_lo=0, mask_hi=0,
                   mask_slope=0.0):
    d = FlatDesc()
    d.n, d.h, d.w, d.cx, d.cin_off, d.cin, d.cy, d.cout_off, d.cout = n, h, w, cx, cin_off, cin, cy, cout_off, cout
    d.cx2, d.cin2_off, d.cin2 = cx2, cin2_off, cin2
    assert len(taps) == 9
    for i, (dy, dx, wi) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_w[i] = dy, dx, wi
    d.out_mode = out_mode
    d.w_taps, d.w_cout_pad, d.w_cin_pad = w_taps, w_rows, w_cols
    d.alpha, d.act, d.slope, d.beta1, d.beta2 = alpha, act, slope, beta1, beta2
    d.res_nch, d.res1_c, d.res1_coff, d.res2_c, d.res2_coff = res_nch, res1_c, res1_coff, res2_c, res2_coff
    d.accumulate = accumulate
    d.mask_c, d.mask_coff, d.mask_lo, d.mask_hi, d.mask_slope = mask_c, mask_coff, mask_lo, mask_hi, mask_slope
    return d


def add_flat(plan, desc, x, w, bias=None, res1=None, res2=None, mask=None, y=None, x2=None):
    plan.keep(desc)
    flops = 2.0 * desc.n * desc.h * desc.w * desc.cout * (desc.cin + desc.cin2) * 9
    info = "cin%d+%d cout%d %dx%dx%d%s" % (desc.cin, desc.cin2, desc.cout, desc.n, desc.h, desc.w,
                                          " acc" if desc.accumulate else "")
    plan.add(lib.b200_conv3x3_flat, C.byref(desc), P(x), P(x2), P(w), P(bias), P(res1), P(res2), P(mask), P(y),
             flops=flops, tag="conv_flat", info=info)
