"""Fused forward / input-gradient engine of the frozen VGG19 feature extractor
(reference: architectures/perceptual.py:73-214; torchvision vgg19.features[:35]).

conv1_1 folds the ImageNet input normalisation (x-mean)/std into its prologue (CUDA-core direct
kernel, K = 27); the other 15 convs are tcgen05 implicit GEMMs with a ReLU epilogue; MaxPool is one
NHWC pass.  Backward is dgrad only (the net is frozen): each dgrad epilogue applies the ReLU mask
of the layer below, MaxPool backward routes to the first maximum (torch semantics).
"""
import torch

from . import _lib
from ._lib import lib
from .runtime import (ConvLayer, ContextPool, pool_for, Lease, P, Plan, WeightPacker, add_igemm, make_conv_desc,
                      require_device, taps_conv, taps_dgrad_s1)

BF16 = torch.bfloat16


class _FContext:
    pass


class FeatureEngine:
    def __init__(self, net):
        self.net = net
        self.device = None
        self.pools = {}

    def _setup(self, device):
        net = self.net
        self.device = device
        self.ops = []  # ('conv', ConvLayer|nn.Conv2d, relu:bool, tap name) | ('pool', ...)
        self.unavailable = set()
        mods = list(net.feature_net._modules.items())
        i = 0
        first = True
        self.conv0 = None
        while i < len(mods):
            name, m = mods[i]
            if isinstance(m, torch.nn.Conv2d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1][1], torch.nn.ReLU)
                if relu:
                    # the stored tensor is the post-ReLU activation: it serves the 'reluX_Y' tap;
                    # the pre-ReLU 'convX_Y' tap of a non-final layer is not materialised
                    self.unavailable.add(name)
                    name = mods[i + 1][0]
                if first:
                    self.conv0 = m
                    self.ops.append(("conv0", m, relu, name))
                    first = False
                else:
                    self.ops.append(("conv", ConvLayer(m, name), relu, name))
                i += 2 if relu else 1
            elif isinstance(m, torch.nn.MaxPool2d):
                if m.kernel_size not in (2, (2, 2)) or m.stride not in (2, (2, 2)):
                    raise NotImplementedError("B200 FeatureExtractor: only MaxPool2d(2, 2)")
                self.ops.append(("pool", None, False, name))
                i += 1
            else:
                raise NotImplementedError("B200 FeatureExtractor: unsupported layer %s" % type(m).__name__)
        self.tc_layers = [o[1] for o in self.ops if o[0] == "conv"]
        self.packer = WeightPacker(self.tc_layers, device)
        self.mean = net.mean.reshape(-1).float().contiguous() if net.use_input_norm else None
        self.std = net.std.reshape(-1).float().contiguous() if net.use_input_norm else None
        self.inv_std = (1.0 / self.std) if self.std is not None else None
        self.pools = {}

    def _ensure(self, x):
        require_device(x, "FeatureExtractor")
        if self.device != x.device or self.packer.stale_pointers():
            self._setup(x.device)

    def _make_context(self, N, H, W):
        dev = self.device
        ctx = _FContext()
        ctx.shape = (N, H, W)
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        ctx.x = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev)
        ctx.T = []      # output tensor of every op
        ctx.dimsT = []
        f = Plan()
        h, w, c = H, W, 3
        prev = None
        ctx.taps = {}
        for kind, L, relu, name in self.ops:
            if kind == "conv0":
                c = L.out_channels
                t = e(N, h, w, c)
                f.add(lib.b200_conv3x3_thin_to_wide, P(ctx.x), P(L.weight), P(L.bias), P(t), N, h, w, 3, c, c, 0, 0,
                      P(self.mean), P(self.std), 1 if relu else 0, 0.0, None, 0, 0, 0.0)
            elif kind == "conv":
                t = e(N, h, w, L.cout)
                d = make_conv_desc(N, h, w, c, 0, c, h, w, h, w, L.cout, 0, L.cout, taps_conv(3, 1), L.taps,
                                   L.fwd_rows, L.fwd_cols, act=1 if relu else 0, slope=0.0)
                add_igemm(f, d, prev, L.w_fwd, L.bias, y=t)
                c = L.cout
            else:
                t = e(N, h // 2, w // 2, c)
                f.add(lib.b200_maxpool2x2, P(prev), P(t), N, h, w, c)
                h, w = h // 2, w // 2
            ctx.T.append(t)
            ctx.dimsT.append((h, w, c))
            ctx.taps[name] = t
            prev = t
        ctx.fwd = f
        ctx.bwd = None
        return ctx

    def _make_backward(self, ctx):
        dev = self.device
        N, H, W = ctx.shape
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        ctx.dT = [e(*t.shape) for t in ctx.T]
        ctx.dx = torch.empty_like(ctx.x)
        b = Plan()
        n_ops = len(self.ops)
        for j in range(n_ops - 1, 0, -1):
            kind, L, relu, name = self.ops[j]
            h, w, c = ctx.dimsT[j]
            hp, wp, cp = ctx.dimsT[j - 1]
            if kind == "conv":
                # dT[j] is the gradient wrt conv j's pre-activation; input of conv j is T[j-1]
                below = self.ops[j - 1]
                mask = ctx.T[j - 1] if (below[0] in ("conv", "conv0") and below[2]) else None
                mk = dict(mask_c=cp, mask_coff=0, mask_lo=0, mask_hi=cp, mask_slope=0.0) if mask is not None else {}
                d = make_conv_desc(N, h, w, c, 0, c, h, w, h, w, cp, 0, cp, taps_dgrad_s1(3, 1), L.taps, L.dgr_rows,
                                   L.dgr_cols, **mk)
                add_igemm(b, d, ctx.dT[j], L.w_dgr, mask=mask, y=ctx.dT[j - 1])
            elif kind == "pool":
                # gradient wrt the pooled tensor -> gradient wrt the pre-activation of the conv below
                b.add(lib.b200_maxpool2x2_bwd, P(ctx.T[j - 1]), P(ctx.dT[j]), P(ctx.dT[j - 1]), N, hp, wp, cp)
        L0 = self.ops[0][1]
        c0 = L0.out_channels
        b.add(lib.b200_conv3x3_wide_to_thin, P(ctx.dT[0]), P(L0.weight), None, P(ctx.dx), N, H, W, c0, c0, 0, 3, 1,
              P(self.inv_std), 1.0)
        ctx.bwd = b

    def forward(self, x, need_backward, listen):
        self._ensure(x)
        N, _, H, W = x.shape
        key = (N, H, W)
        pool = pool_for(self.pools, key, lambda: self._make_context(N, H, W))
        names = [o[3] for o in self.ops]
        for k in listen:   # validate BEFORE leasing a context (a raise must not leak it)
            if k in self.unavailable or k not in names:
                raise NotImplementedError("B200 FeatureExtractor: tap '%s' is not materialised (pre-ReLU taps "
                                          "exist only for the last extracted layer)" % k)
        ctx = pool.acquire()
        self.packer.ensure()
        ctx.x.copy_(x)
        ctx.fwd.run()
        outs = {k: ctx.taps[k].clone() for k in listen}
        if need_backward:
            return outs, Lease(pool, ctx)
        pool.release(ctx)
        return outs, None

    def backward(self, lease, grads):
        """grads: {layer name: NHWC bf16 gradient}; only the last layer may carry a gradient
        together with others being None (multi-tap gradients are summed into dT)."""
        ctx = lease.ctx
        if ctx.bwd is None:
            self._make_backward(ctx)
        names = [o[3] for o in self.ops]
        last = len(self.ops) - 1
        for k, gten in grads.items():
            if gten is None:
                continue
            if names.index(k) != last:
                raise NotImplementedError("B200 FeatureExtractor backward: gradient taps other than the last "
                                          "layer of the extracted stack are not implemented")
            ctx.dT[last].copy_(gten)
            if self.ops[last][2]:
                # the tap is a ReLU OUTPUT ('reluX_Y'): the incoming gradient is wrt the activation, the dgrad
                # chain starts from the pre-activation gradient -> apply this layer's own ReLU mask first
                t = ctx.T[last]
                _lib.check(lib.b200_lrelu_mask_mul(ctx.dT[last].data_ptr(), t.data_ptr(), ctx.dT[last].data_ptr(),
                                                   t.numel(), 0.0, torch.cuda.current_stream().cuda_stream),
                           "lrelu_mask_mul")
        ctx.bwd.run()
        dx = ctx.dx.clone()
        lease.release()
        return dx
