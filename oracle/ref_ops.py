"""CPU oracle for the VFM-VAE decoder hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with plain torch CPU tensor arithmetic, what the reference's
``impl='ref'`` code path computes for the four ops on the hot path.  It is *not*
part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``vfm_vae_b200``) never does and fails loudly without its CUDA library.

Pinning: every function below is checked against outputs of the unmodified
reference (imported from /root/reference in the build container) that are
committed under ``tests/golden/`` together with the script that generated them
(``tools/make_golden.py``).  See ``tests/test_oracle_golden.py``.

Reference lines each function follows (paths relative to the reference repo):

* ``bias_act``             torch_utils/ops/bias_act.py:91-120 (+ table :21-31)
* ``setup_filter``         torch_utils/ops/upfirdn2d.py:70-114
* ``upfirdn2d``            torch_utils/ops/upfirdn2d.py:167-211
* ``filter2d/upsample2d/downsample2d``  torch_utils/ops/upfirdn2d.py:277-387
* ``filtered_lrelu``       torch_utils/ops/filtered_lrelu.py:121-153
* ``conv2d_resample``      torch_utils/ops/conv2d_resample.py:46-141
* ``modulated_conv2d``     networks/generator.py:46-103

The dense contraction itself lives in a third-party dependency of the reference
(``torch.nn.functional.conv2d`` / ``conv_transpose2d``; the reference pins
``torch==2.4.0`` in README.md:82 and calls them from
torch_utils/ops/conv2d_gradfix.py:40,45).  It is restated here from its
published definition (cross-correlation, zero padding; transposed convolution =
scatter of input pixels at ``stride`` spacing) as a sum of shifted matrix
products, so the oracle does not route through the cuDNN/MKLDNN conv at all.

The formulations are deliberately *different* from the reference's (shift-and-add
FIR instead of a depthwise ``conv2d``, tap-wise matmuls instead of a grouped
conv) so that agreement with the golden vectors is evidence, not tautology.
"""

import math

import numpy as np
import torch

# ---------------------------------------------------------------------------
# bias_act

_SELU_SCALE = 1.0507009873554804934193349852946
_SELU_ALPHA = 1.6732632423543772848170429916717

#: name -> (default alpha, default gain, plugin index, which tensor the gradient refers to, has 2nd-order grad)
ACTIVATIONS = {
    'linear':   (0.0, 1.0,          1, '',  False),
    'relu':     (0.0, math.sqrt(2), 2, 'y', False),
    'lrelu':    (0.2, math.sqrt(2), 3, 'y', False),
    'tanh':     (0.0, 1.0,          4, 'y', True),
    'sigmoid':  (0.0, 1.0,          5, 'y', True),
    'elu':      (0.0, 1.0,          6, 'y', True),
    'selu':     (0.0, 1.0,          7, 'y', True),
    'softplus': (0.0, 1.0,          8, 'y', True),
    'swish':    (0.0, math.sqrt(2), 9, 'x', True),
}


def _act(x, name, alpha):
    if name == 'linear':
        return x
    if name == 'relu':
        return torch.where(x > 0, x, torch.zeros_like(x))
    if name == 'lrelu':
        return torch.where(x > 0, x, x * alpha)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    if name == 'elu':
        return torch.where(x >= 0, x, torch.expm1(x))
    if name == 'selu':
        return _SELU_SCALE * torch.where(x >= 0, x, _SELU_ALPHA * torch.expm1(x))
    if name == 'softplus':
        return torch.nn.functional.softplus(x)
    if name == 'swish':
        return torch.sigmoid(x) * x
    raise KeyError(name)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """y = clamp(act(x + b) * gain, +-clamp)."""
    def_alpha, def_gain = ACTIVATIONS[act][0], ACTIVATIONS[act][1]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        assert b.ndim == 1 and b.shape[0] == x.shape[dim]
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    y = _act(x, act, alpha)
    if gain != 1:
        y = y * gain
    if clamp >= 0:
        y = y.clamp(-clamp, clamp)
    return y


# ---------------------------------------------------------------------------
# upfirdn2d

def _pair(v):
    if isinstance(v, int):
        return v, v
    a, b = v
    return int(a), int(b)


def _pad4(p):
    if isinstance(p, int):
        return p, p, p, p
    p = [int(v) for v in p]
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    f = f.clone()
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def _fir_1d_pass(x, taps, axis, up, down, pad0, pad1):
    """Zero-insert, pad/crop, correlate with ``taps`` and decimate along one axis."""
    n = x.shape[axis]
    ft = taps.numel()
    # zero-inserted signal of length n*up
    shape = list(x.shape)
    shape[axis] = n * up
    xu = x.new_zeros(shape)
    idx = [slice(None)] * x.ndim
    idx[axis] = slice(0, n * up, up)
    xu[tuple(idx)] = x
    # pad / crop
    total = n * up + pad0 + pad1
    shape[axis] = max(total, 0)
    xp = x.new_zeros(shape)
    src_lo, src_hi = max(-pad0, 0), n * up - max(-pad1, 0)
    dst_lo = max(pad0, 0)
    if src_hi > src_lo:
        s = [slice(None)] * x.ndim
        d = [slice(None)] * x.ndim
        s[axis] = slice(src_lo, src_hi)
        d[axis] = slice(dst_lo, dst_lo + (src_hi - src_lo))
        xp[tuple(d)] = xu[tuple(s)]
    full = total - ft + 1
    assert full >= 1, 'filter larger than padded signal'
    out_n = (full + down - 1) // down
    shape[axis] = out_n
    y = x.new_zeros(shape)
    for t in range(ft):
        s = [slice(None)] * x.ndim
        s[axis] = slice(t, t + (out_n - 1) * down + 1, down)
        y = y + xp[tuple(s)] * taps[t].to(x.dtype)
    return y


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Per channel: zero-insert by ``up`` -> pad/crop -> FIR (true convolution unless
    ``flip_filter``) * gain -> keep every ``down``-th sample."""
    assert x.ndim == 4
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    f = f.to(torch.float32)
    if f.ndim == 1:
        g = f * (gain ** 0.5)
        g = g if flip_filter else g.flip(0)
        y = _fir_1d_pass(x, g.to(x.dtype), 3, upx, downx, px0, px1)
        y = _fir_1d_pass(y, g.to(x.dtype), 2, upy, downy, py0, py1)
        return y
    # Non-separable: do the zero-insert/pad with unit 1-tap passes, then a 2-D shift-and-add.
    one = torch.ones([1], dtype=x.dtype)
    xp = _fir_1d_pass(x, one, 3, upx, 1, px0, px1)
    xp = _fir_1d_pass(xp, one, 2, upy, 1, py0, py1)
    g = (f * gain).to(x.dtype)
    if not flip_filter:
        g = g.flip([0, 1])
    fh, fw = g.shape
    fullh, fullw = xp.shape[2] - fh + 1, xp.shape[3] - fw + 1
    assert fullh >= 1 and fullw >= 1
    oh, ow = (fullh + downy - 1) // downy, (fullw + downx - 1) // downx
    y = x.new_zeros([x.shape[0], x.shape[1], oh, ow])
    for ty in range(fh):
        for tx in range(fw):
            y = y + g[ty, tx] * xp[:, :, ty: ty + (oh - 1) * downy + 1: downy, tx: tx + (ow - 1) * downx + 1: downx]
    return y


def _fsize(f):
    if f is None:
        return 1, 1
    return int(f.shape[-1]), int(f.shape[0])


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# ---------------------------------------------------------------------------
# filtered_lrelu

def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=math.sqrt(2), slope=0.2,
                   clamp=None, flip_filter=False):
    """down_fd( clamp( lrelu( up_fu(x + b) * up^2 * gain ) ) )."""
    px0, px1, py0, py1 = _pad4(padding)
    y = bias_act(x, b)
    y = upfirdn2d(y, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    y = bias_act(y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    y = upfirdn2d(y, fd, down=down, flip_filter=flip_filter)
    return y


def filtered_lrelu_signs(x, fu=None, b=None, up=1, padding=0, gain=math.sqrt(2), slope=0.2, clamp=None,
                         flip_filter=False):
    """2-bit code per element of the upsampled intermediate (before the down filter):
    1 = negative (slope applied), 2 = clamped (overrides 1).  Mirrors the semantics of
    torch_utils/ops/filtered_lrelu.cu:1140-1153 (act kernel, sign-write mode)."""
    px0, px1, py0, py1 = _pad4(padding)
    y = bias_act(x, b)
    y = upfirdn2d(y, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    v = y * gain
    code = (v < 0).to(torch.uint8)
    v = torch.where(v < 0, v * slope, v)
    if clamp is not None:
        code = torch.where(v.abs() > clamp, torch.full_like(code, 2), code)
    return code


# ---------------------------------------------------------------------------
# dense conv restated (third-party: torch.nn.functional.conv2d / conv_transpose2d)

def _corr2d_grouped(x, w, pad_y=0, pad_x=0, stride=1):
    """Cross-correlation.  x [N,I,H,W]; w [N,O,I,kh,kw] (per-sample weights) or [O,I,kh,kw]."""
    n, i, h, ww = x.shape
    per_sample = (w.ndim == 5)
    kh, kw = w.shape[-2:]
    xp = torch.nn.functional.pad(x, [pad_x, pad_x, pad_y, pad_y])
    oh = (h + 2 * pad_y - kh) // stride + 1
    ow = (ww + 2 * pad_x - kw) // stride + 1
    o = w.shape[-4]
    y = x.new_zeros([n, o, oh, ow])
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, :, ky: ky + (oh - 1) * stride + 1: stride, kx: kx + (ow - 1) * stride + 1: stride]
            patch = patch.reshape(n, i, oh * ow)
            if per_sample:
                y = y + torch.bmm(w[:, :, :, ky, kx], patch).reshape(n, o, oh, ow)
            else:
                y = y + torch.matmul(w[:, :, ky, kx], patch).reshape(n, o, oh, ow)
    return y


def _conv_transpose2d_grouped(x, w, stride, pad_y=0, pad_x=0):
    """Transposed conv: out[n,o, iy*stride+ky-pad, ix*stride+kx-pad] += x[n,i,iy,ix] * w[(n,)o,i,ky,kx]."""
    n, i, h, ww = x.shape
    per_sample = (w.ndim == 5)
    kh, kw = w.shape[-2:]
    o = w.shape[-4]
    fh, fw = (h - 1) * stride + kh, (ww - 1) * stride + kw
    full = x.new_zeros([n, o, fh, fw])
    xf = x.reshape(n, i, h * ww)
    for ky in range(kh):
        for kx in range(kw):
            if per_sample:
                contrib = torch.bmm(w[:, :, :, ky, kx], xf)
            else:
                contrib = torch.matmul(w[:, :, ky, kx], xf)
            full[:, :, ky: ky + (h - 1) * stride + 1: stride, kx: kx + (ww - 1) * stride + 1: stride] += contrib.reshape(n, o, h, ww)
    return full[:, :, pad_y: fh - pad_y, pad_x: fw - pad_x]


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv with optional up/down-sampling.  ``w`` may be [O,I,kh,kw] or per-sample [N,O,I,kh,kw]
    (the latter restates the reference's groups=N grouped conv on a [1,N*I,H,W] view).

    flip_weight=True = cross-correlation (what conv2d does); False = true convolution.
    groups > 1 (torch_utils/ops/conv2d_resample.py:113-118): a grouped conv is `groups` independent convs on channel slices and every
    FIR stage is per channel, so the whole op is evaluated slice by slice."""
    if groups > 1:
        ig, og = x.shape[1] // groups, w.shape[0] // groups
        return torch.cat([conv2d_resample(x[:, g * ig:(g + 1) * ig], w[g * og:(g + 1) * og], f=f, up=up, down=down, padding=padding,
                                          flip_weight=flip_weight, flip_filter=flip_filter) for g in range(groups)], dim=1)
    kh, kw = w.shape[-2:]
    fw, fh = _fsize(f)
    px0, px1, py0, py1 = _pad4(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2

    def corr(xx, ww, **kw_):
        if not flip_weight and (kw > 1 or kh > 1):
            ww = ww.flip([-2, -1])
        return _corr2d_grouped(xx, ww, **kw_)

    if kw == 1 and kh == 1 and down > 1 and up == 1:
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return corr(x, w)
    if kw == 1 and kh == 1 and up > 1 and down == 1:
        x = corr(x, w)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return corr(x, w, stride=down)
    if up > 1:
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        # conv_transpose2d scatters with the weight as given == true convolution of the zero-inserted signal;
        # the reference passes flip_weight=(not flip_weight) to its wrapper, i.e. the taps are reversed only
        # when the caller asked for correlation.
        wt = w.flip([-2, -1]) if (flip_weight and (kw > 1 or kh > 1)) else w
        x = _conv_transpose2d_grouped(x, wt, stride=up, pad_y=pyt, pad_x=pxt)
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return corr(x, w, pad_y=py0, pad_x=px0)
    x = upfirdn2d(x, None, padding=[px0, px1, py0, py1])
    return corr(x, w)


# ---------------------------------------------------------------------------
# modulated conv

def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True):
    """w'[n,o,i,k] = weight[o,i,k] * styles[n,i]; d[n,o] = rsqrt(sum_{i,k} w'^2 + 1e-8);
    y[n] = conv2d_resample(x[n], w'[n] * d[n]) (+ noise)."""
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    assert styles.shape == (n, i)
    w = weight.unsqueeze(0) * styles.reshape(n, 1, i, 1, 1)
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        w = w * d.reshape(n, o, 1, 1, 1)
    y = conv2d_resample(x, w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
    if noise is not None:
        y = y + noise.to(y.dtype)
    return y


def modulated_pointwise_conv2d(x, weight, style, bias=None, demodulate=True):
    """networks/utils/convnext_utils.py:36-57: 1x1 modulated / demodulated conv + broadcast bias (fp32/fp64 semantics; the
    fp16-only pre-normalisation of the reference cancels under demodulation)."""
    y = modulated_conv2d(x, weight, style, noise=None, up=1, padding=0, demodulate=demodulate, flip_weight=True)
    if bias is not None:
        y = y + bias
    return y


# ---------------------------------------------------------------------------
# one legacy synthesis layer (networks/generator.py:240-276) -- used by the decoder-level tests/bench

def synthesis_layer(x, weight, bias, styles, noise, up, resample_filter, act_gain, act_clamp):
    y = modulated_conv2d(x, weight, styles, noise=noise, up=up, padding=weight.shape[-1] // 2,
                         resample_filter=resample_filter, flip_weight=(up == 1))
    return bias_act(y, bias.to(y.dtype), act='lrelu', gain=act_gain, clamp=act_clamp)


def replicate_blur_edges(dy, f, dx):
    """CPU model of ``vfm_replicate_blur_edges`` (include/vfm_ops.h): the border rows / columns of the data gradient of
    ``conv2d(pad(x, k//2, mode='replicate'), f, groups=C)`` (networks/utils/convnext_utils.py:250-255), written into ``dx`` [N,C,H,W]
    (whose interior the caller has filled with the zero-padded stencil over ``dy``).  For output j and input i of one axis the taps
    that land on j after clamping form a range -- [0, p-i] on the first row / column, [n-1+p-i, k-1] on the last, the single tap j-i+p
    elsewhere -- so each weight is a range sum of ``f``.  Plain loops over the 2W + 2(H-2) border elements: small cases only."""
    k = f.shape[0]
    p = k // 2
    h, w = dy.shape[2:]

    def taps(j, i, n):
        if j == 0:
            return range(0, min(k - 1, p - i) + 1)
        if j == n - 1:
            return range(max(0, n - 1 + p - i), k)
        t = j - i + p
        return range(t, t + 1) if 0 <= t < k else range(0)

    border = [(0, x) for x in range(w)] + [(h - 1, x) for x in range(w)] + [(y, x) for y in range(1, h - 1) for x in (0, w - 1)]
    for jy, jx in border:
        acc = torch.zeros(dy.shape[:2], dtype=dy.dtype)
        for iy in range(max(0, jy - p), min(h, jy + p + 1)):
            for ix in range(max(0, jx - p), min(w, jx + p + 1)):
                wgt = sum(float(f[ty, tx]) for ty in taps(jy, iy, h) for tx in taps(jx, ix, w))
                acc = acc + wgt * dy[:, :, iy, ix]
        dx[:, :, jy, jx] = acc
    return dx
