"""2-D convolution with optional up / down sampling -- drop-in for the reference entry point torch_utils/ops/conv2d_resample.py:46
(same signature, same result for every argument combination).

The reference composes a cuDNN conv / transposed conv (through its ``conv2d_gradfix`` pass-through) with its ``upfirdn2d`` plugin.
Here:

* the decoder's cases -- ``groups == 1, down == 1, up in {1, 2}``, one symmetric padding, 1x1 / 3x3 kernels -- run on the sm_100a
  implicit-GEMM kernel of ``modulated_conv2d`` (an unmodulated conv is the modulated one with unit styles and demodulation off), so the
  up=2 semantics (transposed conv, then the FIR with the padding arithmetic of reference lines 82-126) live in exactly one place
  (csrc/modconv_api.cu);
* everything else (``down > 1``, ``groups > 1``, asymmetric / per-axis padding, ``flip_filter``, separable filters with down-sampling)
  is the same *composition* the reference defines, with this package's ``upfirdn2d`` kernels for every FIR / resampling stage and the stock
  dense convolution (``torch.nn.functional.conv2d`` / ``conv_transpose2d``, what the reference itself calls) for the contraction.

Stage plan (what the reference's branch ladder amounts to; ``P`` = the caller's padding plus the filter's own support):

    kernel 1x1, down only      FIR(down, P)              -> conv
    kernel 1x1, up only        conv                      -> FIR(up, P, gain up^2)
    down only                  FIR(P)                    -> conv(stride down)
    up (any down)              convT(stride up, crop)    -> FIR(P', gain up^2)   [-> FIR(down)]
    neither, P symmetric >= 0  conv(padding P)
    otherwise                  FIR(up, P, gain up^2)     -> conv                 [-> FIR(down)]
"""
import torch
import torch.nn.functional as F

from . import upfirdn2d as _fir
from .modulated_conv2d import modulated_conv2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _dense(x, w, stride=1, padding=(0, 0), groups=1, transposed=False, correlate=True):
    """The contraction itself.  ``correlate=False`` flips the taps (true convolution); F.conv2d correlates."""
    if not correlate and (w.shape[2] > 1 or w.shape[3] > 1):
        w = w.flip([2, 3])
    op = F.conv_transpose2d if transposed else F.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def _fast_path_ok(x, w, f, up, down, groups, pads, flip_filter):
    """Can the tensor-core / generic modulated-conv kernel take this call?  (what the decoder uses)"""
    if not x.is_cuda or groups != 1 or down != 1 or up not in (1, 2) or flip_filter:
        return False
    if len(set(pads)) != 1 or pads[0] < 0:
        return False
    kh, kw = w.shape[2], w.shape[3]
    if kh != kw or kh % 2 == 0:
        return False
    if up == 2 and (f is None or (kh == 1)):
        return False
    return x.dtype in (torch.float16, torch.float32, torch.float64)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """Arguments as in the reference (conv2d_resample.py:46-66): ``padding`` is relative to the up-sampled image and may be an int,
    ``[x, y]`` or ``[x0, x1, y0, y1]``; ``flip_weight=True`` correlates (F.conv2d), ``False`` convolves; ``f`` comes from ``setup_filter``."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and isinstance(groups, int) and groups >= 1
    user_pads = _parse_padding(padding)
    if _fast_path_ok(x, w, f, up, down, groups, user_pads, flip_filter):
        f2 = f.ger(f) if (up > 1 and f.ndim == 1) else f
        ones = torch.ones([x.shape[0], x.shape[1]], dtype=torch.float32, device=x.device)
        return modulated_conv2d(x, w, ones, noise=None, up=up, padding=user_pads[0], resample_filter=f2, demodulate=False, flip_weight=flip_weight)

    cout, cin_g, kh, kw = (int(v) for v in w.shape)
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = user_pads
    # the FIR's own support joins the caller's padding: (taps + factor - 1) // 2 in front, (taps - factor) // 2 behind, per resampling stage
    if up > 1:
        px0, px1, py0, py1 = px0 + (fw + up - 1) // 2, px1 + (fw - up) // 2, py0 + (fh + up - 1) // 2, py1 + (fh - up) // 2
    if down > 1:
        px0, px1, py0, py1 = px0 + (fw - down + 1) // 2, px1 + (fw - down) // 2, py0 + (fh - down + 1) // 2, py1 + (fh - down) // 2
    P = [px0, px1, py0, py1]
    fir = _fir.upfirdn2d
    pointwise = kh == 1 and kw == 1

    if pointwise and up == 1 and down > 1:          # decimate first: the 1x1 conv then runs on a quarter of the pixels
        return _dense(fir(x, f, down=down, padding=P, flip_filter=flip_filter), w, groups=groups, correlate=flip_weight)
    if pointwise and up > 1 and down == 1:          # 1x1 conv first: it then runs on the small image
        return fir(_dense(x, w, groups=groups, correlate=flip_weight), f, up=up, padding=P, gain=up ** 2, flip_filter=flip_filter)
    if up == 1 and down > 1:                        # low-pass at full rate, decimation folded into the conv's stride
        return _dense(fir(x, f, padding=P, flip_filter=flip_filter), w, stride=down, groups=groups, correlate=flip_weight)
    if up > 1:
        # zero-insertion folded into a transposed conv: its weight is [in, out/groups, kh, kw], its taps flipped w.r.t. a plain conv
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).reshape(groups * cin_g, cout // groups, kh, kw)
        qx0, qx1, qy0, qy1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        cx = max(min(-qx0, -qx1), 0)                # the part of the (negative) padding the transposed conv can crop itself
        cy = max(min(-qy0, -qy1), 0)
        y = _dense(x, wt, stride=up, padding=(cy, cx), groups=groups, transposed=True, correlate=not flip_weight)
        y = fir(y, f, padding=[qx0 + cx, qx1 + cx, qy0 + cy, qy1 + cy], gain=up ** 2, flip_filter=flip_filter)
        return fir(y, f, down=down, flip_filter=flip_filter) if down > 1 else y
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:      # up == down == 1: a plain padded conv
        return _dense(x, w, padding=(py0, px0), groups=groups, correlate=flip_weight)
    # up == down == 1 with asymmetric or negative padding: pad / crop with the FIR stage (f = None is the identity filter), then convolve
    return _dense(fir(x, None, padding=P, flip_filter=flip_filter), w, groups=groups, correlate=flip_weight)
