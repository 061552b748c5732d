"""Tensor-level wrappers around the C ABI (one call = one kernel), used by the unit tests and as
the small public functional API.  The network engines do NOT go through these (they run static
plans, runtime.py); both end in the same C entry points.

Layouts: activations NHWC bf16 (`[N, H, W, C]` contiguous), images NCHW fp32, conv weights fp32
OIHW exactly as nn.Conv2d holds them.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import PackEntry, WgradDesc, lib
from .runtime import (make_conv_desc, make_flat_desc, require_device, roundup, stream_ptr, taps_conv, taps_dgrad_s1,
                      taps_dgrad_s2_k4)

BF16 = torch.bfloat16


def _p(t):
    return None if t is None else t.data_ptr()


def pack_weight(w_oihw, mode, co_mul=0, co_off=0, cout=None):
    """fp32 [Cout, Cin, kh, kw] -> bf16 [taps][rows_pad][cols_pad] (mode 0: rows=co; 1: rows=ci)."""
    require_device(w_oihw, "pack_weight")
    w = w_oihw.detach().float().contiguous()
    co, ci, kh, kw = w.shape
    if cout is not None:
        co = cout
    taps = kh * kw
    rows, cols = (roundup(co, 16), roundup(ci, 64)) if mode == 0 else (roundup(ci, 16), roundup(co, 64))
    dst = torch.empty(taps, rows, cols, dtype=BF16, device=w.device)
    e = PackEntry(w.data_ptr(), dst.data_ptr(), co, ci, taps, rows, cols, mode, co_mul, co_off)
    table = torch.frombuffer(bytearray(bytes(e)), dtype=torch.uint8).to(w.device)
    _lib.check(lib.b200_pack_weights(table.data_ptr(), 1, taps * rows * cols, stream_ptr()), "pack_weights")
    return dst


def conv2d(x, w_oihw, bias=None, stride=1, padding=1, cin_off=0, cin=None, out=None, cout_off=0, act=0,
           slope=0.2, alpha=1.0, res1=None, res1_coff=0, beta1=0.0, res2=None, res2_coff=0, beta2=0.0,
           res_nch=0, upsample2x=False):
    """Implicit-GEMM conv forward on tcgen05.  x: [N,H,W,Cx] bf16 (uses channels
    [cin_off, cin_off+cin)); out: [N,Ho,Wo,Cy] bf16 (writes channels [cout_off, cout_off+Cout))."""
    require_device(x, "conv2d")
    N, H, W, Cx = x.shape
    Cout, Cin, kh, kw = w_oihw.shape
    cin = Cin if cin is None else cin
    assert cin == Cin
    Ho = (H + 2 * padding - kh) // stride + 1
    Wo = (W + 2 * padding - kw) // stride + 1
    Hb, Wb = (2 * Ho, 2 * Wo) if upsample2x else (Ho, Wo)
    if out is None:
        out = torch.empty(N, Hb, Wb, Cout, dtype=BF16, device=x.device)
    wp = pack_weight(w_oihw, 0)
    d = make_conv_desc(N, H, W, Cx, cin_off, cin, Ho, Wo, Hb, Wb, out.shape[3], cout_off, Cout,
                       taps_conv(kh, padding), kh * kw, wp.shape[1], wp.shape[2], in_stride=stride,
                       upsample=1 if upsample2x else 0, alpha=alpha, act=act, slope=slope, beta1=beta1,
                       beta2=beta2, res_nch=res_nch,
                       res1_c=res1.shape[3] if res1 is not None else 0, res1_coff=res1_coff,
                       res2_c=res2.shape[3] if res2 is not None else 0, res2_coff=res2_coff)
    b = bias.detach().float().contiguous() if bias is not None else None
    _lib.check(lib.b200_conv_igemm(C.byref(d), _p(x), _p(wp), _p(b), _p(res1), _p(res2), None, _p(out),
                                   stream_ptr()), "conv_igemm")
    return out


def conv2d_dgrad(dy, w_oihw, in_hw, stride=1, padding=1, dy_coff=0, out=None, dx_coff=0, alpha=1.0,
                 accumulate=False, mask=None, mask_coff=0, mask_lo=0, mask_hi=0, mask_slope=0.2, res1=None,
                 res1_coff=0, beta1=0.0, res_nch=0):
    """Input gradient of conv2d.  dy: [N,Ho,Wo,Cdy] bf16 (channels [dy_coff, dy_coff+Cout));
    result [N,H,W,Cdx] channels [dx_coff, dx_coff+Cin)."""
    require_device(dy, "conv2d_dgrad")
    N, Ho, Wo, Cdy = dy.shape
    Cout, Cin, kh, kw = w_oihw.shape
    H, W = in_hw
    if out is None:
        out = torch.zeros(N, H, W, Cin, dtype=BF16, device=dy.device)
    wp = pack_weight(w_oihw, 1)
    common = dict(alpha=alpha, accumulate=1 if accumulate else 0,
                  mask_c=mask.shape[3] if mask is not None else 0, mask_coff=mask_coff, mask_lo=mask_lo,
                  mask_hi=mask_hi, mask_slope=mask_slope, beta1=beta1, res_nch=res_nch,
                  res1_c=res1.shape[3] if res1 is not None else 0, res1_coff=res1_coff)
    descs = []
    if stride == 1:
        descs.append(make_conv_desc(N, Ho, Wo, Cdy, dy_coff, Cout, H, W, H, W, out.shape[3], dx_coff, Cin,
                                    taps_dgrad_s1(kh, padding), kh * kw, wp.shape[1], wp.shape[2], **common))
    else:
        assert stride == 2 and kh == 4 and padding == 1 and H % 2 == 0 and W % 2 == 0
        for py in (0, 1):
            for px in (0, 1):
                descs.append(make_conv_desc(N, Ho, Wo, Cdy, dy_coff, Cout, H // 2, W // 2, H, W, out.shape[3],
                                            dx_coff, Cin, taps_dgrad_s2_k4(py, px), 16, wp.shape[1],
                                            wp.shape[2], out_mul=(2, 2), out_off=(py, px), **common))
    for d in descs:
        _lib.check(lib.b200_conv_igemm(C.byref(d), _p(dy), _p(wp), None, _p(res1), None, _p(mask), _p(out),
                                       stream_ptr()), "conv_igemm(dgrad)")
    return out


def conv2d_wgrad(x, dy, w_shape, stride=1, padding=1, x_coff=0, dy_coff=0, scale=1.0, want_bias=True):
    """Weight (+bias) gradient, fp32 OIHW."""
    require_device(x, "conv2d_wgrad")
    N, H, W, Cx = x.shape
    _, Ho, Wo, Cdy = dy.shape
    Cout, Cin, kh, kw = w_shape
    dw = torch.zeros(w_shape, dtype=torch.float32, device=x.device)
    db = torch.zeros(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    d = WgradDesc(N, H, W, Cx, x_coff, Cin, Ho, Wo, Cdy, dy_coff, Cout, kh, kw, stride, padding, scale)
    _lib.check(lib.b200_conv_wgrad(C.byref(d), _p(x), _p(dy), _p(dw), _p(db), stream_ptr()), "conv_wgrad")
    return dw, db


def conv3x3_thin_to_wide(x, w, bias=None, transpose_w=False, cw=None, mean=None, std=None, act=0, slope=0.2,
                         mask=None, mask_slope=0.2):
    require_device(x, "conv3x3_thin_to_wide")
    N, cs, H, W = x.shape
    cw = (w.shape[1] if transpose_w else w.shape[0]) if cw is None else cw
    y = torch.empty(N, H, W, cw, dtype=BF16, device=x.device)
    _lib.check(lib.b200_conv3x3_thin_to_wide(_p(x), _p(w), _p(bias), _p(y), N, H, W, cs, cw, cw, 0,
                                             1 if transpose_w else 0, _p(mean), _p(std), act, slope, _p(mask),
                                             mask.shape[3] if mask is not None else 0, 0, mask_slope,
                                             stream_ptr()), "thin_to_wide")
    return y


def conv3x3_wide_to_thin(x, w, bias=None, transpose_w=False, cs=None, inv_std=None, out_scale=1.0):
    require_device(x, "conv3x3_wide_to_thin")
    N, H, W, cw = x.shape
    cs = (w.shape[1] if transpose_w else w.shape[0]) if cs is None else cs
    y = torch.empty(N, cs, H, W, dtype=torch.float32, device=x.device)
    _lib.check(lib.b200_conv3x3_wide_to_thin(_p(x), _p(w), _p(bias), _p(y), N, H, W, cw, cw, 0, cs,
                                             1 if transpose_w else 0, _p(inv_std), out_scale, stream_ptr()),
               "wide_to_thin")
    return y


def conv3x3_thin_wgrad(thin, wide, wide_is_out, want_bias_wide=False, want_bias_thin=False):
    require_device(thin, "conv3x3_thin_wgrad")
    N, cs, H, W = thin.shape
    cw = wide.shape[3]
    shape = (cw, cs, 3, 3) if wide_is_out else (cs, cw, 3, 3)
    dw = torch.zeros(shape, dtype=torch.float32, device=thin.device)
    dbw = torch.zeros(cw, dtype=torch.float32, device=thin.device) if want_bias_wide else None
    dbt = torch.zeros(cs, dtype=torch.float32, device=thin.device) if want_bias_thin else None
    _lib.check(lib.b200_conv3x3_thin_wgrad(_p(thin), _p(wide), _p(dw), _p(dbw), _p(dbt), N, H, W, cs, cw, cw, 0,
                                           1 if wide_is_out else 0, None, None, stream_ptr()), "thin_wgrad")
    return dw, dbw, dbt


def batchnorm_lrelu_train(z, gamma, beta, running_mean=None, running_var=None, momentum=0.1, eps=1e-5,
                          slope=0.2):
    """BatchNorm2d (batch statistics) + LeakyReLU on NHWC bf16. Returns (a, mean_invstd)."""
    require_device(z, "batchnorm_lrelu_train")
    c = z.shape[-1]
    npix = z.numel() // c
    stats = torch.empty(2 * c, dtype=torch.float32, device=z.device)
    mi = torch.empty(2 * c, dtype=torch.float32, device=z.device)
    a = torch.empty_like(z)
    s = stream_ptr()
    _lib.check(lib.b200_bn_stats_finalize(_p(z), _p(stats), _p(mi), _p(running_mean), _p(running_var), npix, c,
                                          momentum, eps, s), "bn_stats_finalize")
    _lib.check(lib.b200_bn_apply_lrelu(_p(z), _p(mi), _p(gamma), _p(beta), _p(a), npix, c, slope, s), "bn_apply")
    return a, mi


def batchnorm_lrelu_backward(z, da, mi, gamma, beta, slope=0.2, use_batch_stats=True):
    require_device(z, "batchnorm_lrelu_backward")
    c = z.shape[-1]
    npix = z.numel() // c
    sums = torch.empty(2 * c, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z)
    dgamma = torch.zeros(c, dtype=torch.float32, device=z.device)
    dbeta = torch.zeros(c, dtype=torch.float32, device=z.device)
    s = stream_ptr()
    _lib.check(lib.b200_bn_bwd_reduce(_p(z), _p(da), _p(mi), _p(gamma), _p(beta), _p(sums), _p(dgamma), _p(dbeta), npix,
                                      c, slope, s), "bn_bwd_reduce")
    _lib.check(lib.b200_bn_bwd_apply(_p(z), _p(da), _p(mi), _p(gamma), _p(beta), _p(sums), _p(dz), npix, c, slope,
                                     1 if use_batch_stats else 0, s), "bn_bwd_apply")
    return dz, dgamma, dbeta


def maxpool2x2(x):
    require_device(x, "maxpool2x2")
    N, H, W, Cc = x.shape
    y = torch.empty(N, H // 2, W // 2, Cc, dtype=BF16, device=x.device)
    _lib.check(lib.b200_maxpool2x2(_p(x), _p(y), N, H, W, Cc, stream_ptr()), "maxpool2x2")
    return y


def pixel_shuffle2(z, act=0, slope=0.2):
    """nn.PixelShuffle(2) (+ LeakyReLU) on NHWC bf16: [N,h,w,4c] -> [N,2h,2w,c] (block.py:383)."""
    require_device(z, "pixel_shuffle2")
    N, H, W, C4 = z.shape
    y = torch.empty(N, 2 * H, 2 * W, C4 // 4, dtype=BF16, device=z.device)
    _lib.check(lib.b200_pixel_shuffle2(_p(z), _p(y), N, H, W, C4 // 4, int(act), float(slope), stream_ptr()),
               "pixel_shuffle2")
    return y


def pixel_unshuffle2(dy):
    """Input gradient of pixel_shuffle2: [N,2h,2w,c] -> [N,h,w,4c]."""
    require_device(dy, "pixel_unshuffle2")
    N, H2, W2, Cc = dy.shape
    dz = torch.empty(N, H2 // 2, W2 // 2, 4 * Cc, dtype=BF16, device=dy.device)
    _lib.check(lib.b200_pixel_unshuffle2(_p(dy), _p(dz), N, H2 // 2, W2 // 2, Cc, stream_ptr()), "pixel_unshuffle2")
    return dz


def maxpool2x2_backward(x, dy):
    require_device(x, "maxpool2x2_backward")
    N, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    _lib.check(lib.b200_maxpool2x2_bwd(_p(x), _p(dy), _p(dx), N, H, W, Cc, stream_ptr()), "maxpool2x2_bwd")
    return dx


def sumpool2x2_mask(dy, mask_up=None, slope=0.2):
    require_device(dy, "sumpool2x2_mask")
    N, H2, W2, Cc = dy.shape
    dx = torch.empty(N, H2 // 2, W2 // 2, Cc, dtype=BF16, device=dy.device)
    _lib.check(lib.b200_sumpool2x2_mask(_p(dy), _p(mask_up), _p(dx), N, H2 // 2, W2 // 2, Cc, slope, stream_ptr()),
               "sumpool2x2_mask")
    return dx


def l1_loss_with_grad(a, b, weight=1.0):
    """mean |a - b| * weight and its gradient wrt a (same dtype/layout as a), one pass."""
    require_device(a, "l1_loss")
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=a.device)
    grad = torch.empty_like(a)
    fn = lib.b200_l1_loss_f32 if a.dtype == torch.float32 else lib.b200_l1_loss_bf16
    _lib.check(fn(_p(a), _p(b), _p(loss), _p(grad), a.numel(), float(weight), stream_ptr()), "l1_loss")
    return loss[0], grad


def to_flat(x_dense, c_total=None, coff=0):
    """dense [N,H,W,C] -> zero-bordered flat [N,H+2,W+2,c_total] with x at channels [coff, coff+C)."""
    require_device(x_dense, "to_flat")
    N, H, W, Cc = x_dense.shape
    ct = Cc if c_total is None else c_total
    out = torch.zeros(N, H + 2, W + 2, ct, dtype=BF16, device=x_dense.device)
    _lib.check(lib.b200_pad_copy(_p(out), ct, coff, _p(x_dense), Cc, 0, N, H, W, Cc, stream_ptr()), "pad_copy")
    return out


def from_flat(x_flat, coff=0, c=None, add=None):
    require_device(x_flat, "from_flat")
    N, Hp, Wp, Ct = x_flat.shape
    c = Ct - coff if c is None else c
    out = torch.empty(N, Hp - 2, Wp - 2, c, dtype=BF16, device=x_flat.device)
    _lib.check(lib.b200_unpad_add(_p(out), c, _p(x_flat), Ct, coff, _p(add), add.shape[3] if add is not None else 8,
                                  N, Hp - 2, Wp - 2, c, stream_ptr()), "unpad_add")
    return out


def conv3x3_flat(x_flat, w_oihw, bias=None, dgrad=False, cin_off=0, out=None, cout_off=0, out_mode=0, act=0,
                 slope=0.2, alpha=1.0, res1=None, res1_coff=0, beta1=0.0, res2=None, res2_coff=0, beta2=0.0,
                 res_nch=0, accumulate=False, mask=None, mask_coff=0, mask_lo=0, mask_hi=0, mask_slope=0.2):
    """3x3 s1 p1 conv (or its input gradient when dgrad=True) on zero-bordered flat tensors."""
    require_device(x_flat, "conv3x3_flat")
    N, Hp, Wp, Cx = x_flat.shape
    h, w = Hp - 2, Wp - 2
    Cout, Cin, kh, kw = w_oihw.shape
    assert kh == 3 and kw == 3
    K, Nout = (Cout, Cin) if dgrad else (Cin, Cout)
    wp = pack_weight(w_oihw, 1 if dgrad else 0)
    taps = taps_dgrad_s1(3, 1) if dgrad else taps_conv(3, 1)
    if out is None:
        shape = {0: (N, Hp, Wp, Nout), 1: (N, h, w, Nout), 2: (N, 2 * h, 2 * w, Nout)}[out_mode]
        out = torch.zeros(*shape, dtype=BF16, device=x_flat.device)
    d = make_flat_desc(N, h, w, Cx, cin_off, K, out.shape[3], cout_off, Nout, taps, 9, wp.shape[1], wp.shape[2],
                       out_mode=out_mode, alpha=alpha, act=act, slope=slope, beta1=beta1, beta2=beta2,
                       res_nch=res_nch, res1_c=res1.shape[3] if res1 is not None else 0, res1_coff=res1_coff,
                       res2_c=res2.shape[3] if res2 is not None else 0, res2_coff=res2_coff,
                       accumulate=1 if accumulate else 0, mask_c=mask.shape[3] if mask is not None else 0,
                       mask_coff=mask_coff, mask_lo=mask_lo, mask_hi=mask_hi, mask_slope=mask_slope)
    b = bias.detach().float().contiguous() if bias is not None else None
    _lib.check(lib.b200_conv3x3_flat(C.byref(d), _p(x_flat), None, _p(wp), _p(b), _p(res1), _p(res2), _p(mask), _p(out),
                                     stream_ptr()), "conv3x3_flat")
    return out
