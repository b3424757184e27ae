"""Plugin objects with the same callable surface as the reference's three JIT-built pybind modules, backed by
libvfmops.so through the C ABI (include/vfm_ops.h).

Reference signatures mirrored (positional, same argument meaning, RuntimeError where the reference TORCH_CHECKs):

* ``bias_act_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp) -> Tensor``
  (torch_utils/ops/bias_act.cpp:32-90)
* ``upfirdn2d_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain) -> Tensor``
  (torch_utils/ops/upfirdn2d.cpp:16-98)
* ``filtered_lrelu_plugin.filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope,
  clamp, flip_filters, writeSigns) -> (y, so, return_code)`` and ``.filtered_lrelu_act_(x, si, sx, sy, gain, slope,
  clamp, writeSigns) -> so``  (torch_utils/ops/filtered_lrelu.cpp:16-290)
* new: ``modconv_plugin.forward / .backward`` for networks/generator.py:46 ``modulated_conv2d`` (the reference has no
  native boundary for it).

"Absent" optional tensors are zero-element tensors, exactly like the reference (``_null_tensor``).  Outputs are
allocated here with torch (same device, layout preserved); the library itself never allocates.
"""
import ctypes as C

import torch

from . import _lib

_DT = {torch.float16: _lib.VFM_F16, torch.float32: _lib.VFM_F32, torch.float64: _lib.VFM_F64}


def _ptr(t):
    return None if (t is None or t.numel() == 0) else t.data_ptr()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _dense(t):
    """non-overlapping and dense (what x.is_non_overlapping_and_dense() checks in C++)."""
    if t.numel() == 0:
        return True
    dims = sorted(((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1))
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


def _same_layout(a, b):
    if a.dim() != b.dim():
        return False
    for i in range(a.dim()):
        if a.size(i) != b.size(i):
            return False
        if a.size(i) >= 2 and a.stride(i) != b.stride(i):
            return False
    return True


def _dtype_code(t, what):
    _check(t.dtype in _DT, f'{what}: unsupported dtype {t.dtype} (float16/float32/float64 only)')
    return _DT[t.dtype]


# ---------------------------------------------------------------------------------------------------------------

class _BiasActPlugin:
    @staticmethod
    def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, db=None):
        """``db`` (extension): optional zero-initialised fp32 [C] tensor that receives the fused bias gradient."""
        _check(x.is_cuda, 'x must reside on CUDA device')
        _check(b.numel() == 0 or (b.dtype == x.dtype and b.device == x.device), 'b must have the same dtype and device as x')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            _check(t.numel() == 0 or (t.shape == x.shape and t.dtype == x.dtype and t.device == x.device),
                   f'{name} must have the same shape, dtype, and device as x')
        _check(b.dim() == 1, 'b must have rank 1')
        _check(b.numel() == 0 or (0 <= dim < x.dim()), 'dim is out of bounds')
        _check(b.numel() == 0 or b.numel() == x.size(dim), 'b has wrong number of elements')
        _check(grad >= 0, 'grad must be non-negative')
        _check(_dense(x), 'x must be non-overlapping and dense')
        _check(b.is_contiguous(), 'b must be contiguous')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            _check(t.numel() == 0 or _same_layout(t, x), f'{name} must have the same layout as x')
        y = torch.empty_like(x)
        _check(_same_layout(y, x), 'y must have the same layout as x')
        if x.numel() == 0:
            return y    # the reference launches an empty grid here
        p = _lib.BiasActParams()
        p.x, p.b, p.xref, p.yref, p.dy, p.y = _ptr(x), _ptr(b), _ptr(xref), _ptr(yref), _ptr(dy), _ptr(y)
        p.db = None
        p.dtype = _dtype_code(x, 'bias_act')
        p.grad, p.act, p.alpha, p.gain, p.clamp = int(grad), int(act), float(alpha), float(gain), float(clamp)
        p.size_x = x.numel()
        p.size_b = b.numel()
        p.step_b = x.stride(dim) if b.numel() else 1
        if db is not None:
            _check(db.dtype == torch.float32 and db.is_contiguous() and db.device == x.device, 'db must be contiguous float32 on the device of x')
            _check(0 <= dim < x.dim() and db.numel() == x.size(dim), 'db has wrong number of elements')
            p.db = _ptr(db)
            p.size_b = db.numel()
            p.step_b = x.stride(dim)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().vfm_bias_act(C.byref(p), _stream(x)), 'bias_act')
        return y


class _Upfirdn2dPlugin:
    @staticmethod
    def pixel_shuffle2(x, inverse=False):
        """``F.pixel_shuffle(x, 2)`` / with ``inverse`` ``F.pixel_unshuffle(x, 2)`` (extension; contiguous fp16/fp32 NCHW, width of
        the 4-plane tensor % 4 == 0)."""
        _check(x.is_cuda and x.dim() == 4 and x.is_contiguous(), 'pixel_shuffle2: x must be a contiguous NCHW CUDA tensor')
        if inverse:
            n, c, h2, w2 = x.shape
            _check(h2 % 2 == 0 and w2 % 8 == 0, 'pixel_unshuffle2: bad input shape')
            c_out, h, w = c, h2 // 2, w2 // 2
            y = torch.empty([n, 4 * c, h, w], dtype=x.dtype, device=x.device)
        else:
            n, c4, h, w = x.shape
            _check(c4 % 4 == 0 and w % 4 == 0, 'pixel_shuffle2: bad input shape')
            c_out = c4 // 4
            y = torch.empty([n, c_out, 2 * h, 2 * w], dtype=x.dtype, device=x.device)
        p = _lib.PixelShuffle2Params()
        p.x, p.y, p.dtype = _ptr(x), _ptr(y), _dtype_code(x, 'pixel_shuffle2')
        p.batch, p.out_channels, p.in_h, p.in_w, p.inverse = n, c_out, h, w, int(bool(inverse))
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().vfm_pixel_shuffle2(C.byref(p), _stream(x)), 'pixel_shuffle2')
        return y

    @staticmethod
    def replicate_blur_edges(dy, dx, f):
        """Overwrites the border rows / columns of ``dx`` (the zero-padded stencil pass over ``dy``) with the data gradient of the
        replicate-padded blur (extension; ``vfm_replicate_blur_edges``).  In place; returns ``dx``."""
        _check(dy.is_cuda and dy.dim() == 4 and dy.is_contiguous() and dx.is_contiguous() and dx.shape == dy.shape and dx.dtype == dy.dtype,
               'replicate_blur_edges: dy / dx must be contiguous NCHW CUDA tensors of the same shape and dtype')
        _check(f.dtype == torch.float32 and f.is_contiguous() and f.dim() == 2 and f.shape[0] == f.shape[1] and f.device == dy.device,
               'replicate_blur_edges: f must be a contiguous float32 [k,k] tensor on the device of dy')
        p = _lib.ReplicateBlurEdgesParams()
        p.dy, p.dx, p.f, p.dtype, p.k = _ptr(dy), _ptr(dx), _ptr(f), _dtype_code(dy, 'replicate_blur_edges'), f.shape[0]
        p.planes, p.h, p.w = dy.shape[0] * dy.shape[1], dy.shape[2], dy.shape[3]
        with torch.cuda.device(dy.device):
            _lib.check(_lib.load().vfm_replicate_blur_edges(C.byref(p), _stream(dy)), 'replicate_blur_edges')
        return dx

    @staticmethod
    def depthwise_wgrad(x, dy, k, want_bias=True):
        """-> (dweight [C,k,k] fp32, dbias [C] fp32 | None) of the depthwise k x k conv (extension), or None where no kernel applies."""
        _check(x.is_cuda and x.dim() == 4 and x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype,
               'depthwise_wgrad: x and dy must be contiguous NCHW CUDA tensors of the same shape and dtype')
        n, c, h, w = x.shape
        dw = torch.zeros([c, k, k], dtype=torch.float32, device=x.device)
        db = torch.zeros([c], dtype=torch.float32, device=x.device) if want_bias else None
        p = _lib.DepthwiseWgradParams()
        p.x, p.dy, p.dweight, p.dbias, p.dtype = _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), _dtype_code(x, 'depthwise_wgrad')
        p.batch, p.channels, p.h, p.w, p.k = n, c, h, w, int(k)
        with torch.cuda.device(x.device):
            st = _lib.load().vfm_depthwise_wgrad(C.byref(p), _stream(x))
        if st == _lib.VFM_ERR_NO_KERNEL:
            return None
        _lib.check(st, 'depthwise_wgrad')
        return dw, db

    @staticmethod
    def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain, add=None, pad_mode=0, bias=None):
        """``pad_mode=1`` (extension): replicate padding; a rank-3 ``f`` [C,fh,fw] (extension) is one filter per channel = a depthwise
        conv, optionally with ``bias`` [C]; both return None where no kernel implements them."""
        _check(x.is_cuda, 'x must reside on CUDA device')
        _check(f.device == x.device, 'f must reside on the same device as x')
        _check(f.dtype == torch.float32, 'f must be float32')
        _check(x.numel() > 0, 'x has zero size')
        _check(f.numel() > 0, 'f has zero size')
        _check(x.dim() == 4, 'x must be rank 4')
        per_channel = f.dim() == 3
        _check(f.dim() == 2 or (per_channel and f.size(0) == x.size(1) and f.is_contiguous()), 'f must be rank 2 (or [C,fh,fw] contiguous)')
        _check(upx >= 1 and upy >= 1, 'upsampling factor must be at least 1')
        _check(downx >= 1 and downy >= 1, 'downsampling factor must be at least 1')
        out_w = (x.size(3) * upx + padx0 + padx1 - f.size(-1) + downx) // downx
        out_h = (x.size(2) * upy + pady0 + pady1 - f.size(-2) + downy) // downy
        _check(out_w >= 1 and out_h >= 1, 'output must be at least 1x1')
        cl = x.dim() == 4 and x.stride(1) == 1 and x.size(1) > 1 and not x.is_contiguous()
        y = torch.empty([x.size(0), x.size(1), out_h, out_w], dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        p = _lib.Upfirdn2dParams()
        p.x, p.f, p.y = _ptr(x), _ptr(f), _ptr(y)
        p.dtype = _dtype_code(x, 'upfirdn2d')
        p.upx, p.upy, p.downx, p.downy = int(upx), int(upy), int(downx), int(downy)
        p.padx0, p.pady0, p.flip, p.gain = int(padx0), int(pady0), int(bool(flip)), float(gain)
        p.in_w, p.in_h, p.channels, p.batch = x.size(3), x.size(2), x.size(1), x.size(0)
        p.in_stride_w, p.in_stride_h, p.in_stride_c, p.in_stride_n = x.stride(3), x.stride(2), x.stride(1), x.stride(0)
        p.fw, p.fh, p.f_stride_w, p.f_stride_h = f.size(-1), f.size(-2), f.stride(-1), f.stride(-2)
        p.f_stride_c = f.stride(0) if per_channel else 0
        if bias is not None:
            p.ep_enable, p.ep_act, p.ep_alpha, p.ep_gain, p.ep_clamp, p.ep_bias = 1, 1, 0.0, 1.0, -1.0, _ptr(bias)
        p.out_w, p.out_h = out_w, out_h
        p.out_stride_w, p.out_stride_h, p.out_stride_c, p.out_stride_n = y.stride(3), y.stride(2), y.stride(1), y.stride(0)
        p.add, p.add_stride_h, p.add_stride_n = None, 0, 0
        if add is not None:
            _check(add.dtype == torch.float32 and add.is_contiguous() and add.shape[-2:] == (out_h, out_w), 'add must be fp32 [..,out_h,out_w]')
            p.add, p.add_stride_h = _ptr(add), out_w
            p.add_stride_n = out_h * out_w if add.numel() == x.size(0) * out_h * out_w and add.dim() > 2 and x.size(0) > 1 else 0
        p.pad_mode = int(pad_mode)
        with torch.cuda.device(x.device):
            st = _lib.load().vfm_upfirdn2d(C.byref(p), _stream(x))
        if (pad_mode or per_channel) and st == _lib.VFM_ERR_NO_KERNEL:
            return None
        _lib.check(st, 'upfirdn2d')
        return y


class _FilteredLreluPlugin:
    @staticmethod
    def filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, writeSigns, y_sum=None):
        """``y_sum`` (extension): optional zero-initialised fp32 [C] tensor that receives the per-channel sum of the output (the fused bias
        gradient when this call is the backward pass)."""
        _check(x.is_cuda, 'x must reside on CUDA device')
        _check(fu.device == x.device and fd.device == x.device and b.device == x.device, 'all input tensors must reside on the same device')
        _check(fu.dtype == torch.float32 and fd.dtype == torch.float32, 'fu and fd must be float32')
        _check(b.dtype == x.dtype, 'x and b must have the same dtype')
        _check(x.dtype in (torch.float16, torch.float32), 'x and b must be float16 or float32')
        _check(x.dim() == 4, 'x must be rank 4')
        _check(x.numel() > 0, 'x is empty')
        _check(fu.dim() in (1, 2) and fd.dim() in (1, 2), 'fu and fd must be rank 1 or 2')
        _check(fu.numel() > 0, 'fu is empty')
        _check(fd.numel() > 0, 'fd is empty')
        _check(b.dim() == 1 and b.size(0) == x.size(1), 'b must be a vector with the same number of channels as x')
        _check(up >= 1 and down >= 1, 'up and down must be at least 1')
        xw, xh = x.size(3), x.size(2)
        fut_w, fut_h = fu.size(-1) - 1, fu.size(0) - 1
        fdt_w, fdt_h = fd.size(-1) - 1, fd.size(0) - 1
        cw = xw * up + (px0 + px1) - fut_w
        ch = xh * up + (py0 + py1) - fut_h
        _check(cw > fdt_w and ch > fdt_h, 'upsampled buffer must be at least the size of downsampling filter')
        yw = (cw - fdt_w + (down - 1)) // down
        yh = (ch - fdt_h + (down - 1)) // down
        _check(yw > 0 and yh > 0, 'output must be at least 1x1')
        read_signs = si.numel() > 0
        p = _lib.FilteredLreluParams()
        p.dtype = _DT[x.dtype]
        p.up, p.down = int(up), int(down)
        p.fu_w, p.fu_h = fu.size(-1), (fu.size(0) if fu.dim() == 2 else 0)
        p.fd_w, p.fd_h = fd.size(-1), (fd.size(0) if fd.dim() == 2 else 0)
        p.fu_stride_w, p.fu_stride_h = fu.stride(-1), (fu.stride(0) if fu.dim() == 2 else 0)
        p.fd_stride_w, p.fd_stride_h = fd.stride(-1), (fd.stride(0) if fd.dim() == 2 else 0)
        p.pad_x0, p.pad_y0 = int(px0), int(py0)
        p.gain, p.slope, p.clamp = float(gain), float(slope), float(clamp)
        p.flip, p.write_signs, p.read_signs = int(bool(flip_filters)), int(bool(writeSigns)), int(read_signs)
        p.x_w, p.x_h, p.channels, p.batch = xw, xh, x.size(1), x.size(0)
        # Probe the kernel envelope before allocating anything, like the reference's test_spec (filtered_lrelu.cpp:47-56).
        lib = _lib.load()
        if (x.dtype not in (torch.float16, torch.float32) or up not in (1, 2, 4) or down not in (1, 2, 4)
                or max(fu.size(-1), fu.size(0), fd.size(-1), fd.size(0)) > 32):
            return None, None, -1
        cl = x.stride(1) == 1 and x.size(1) > 1 and not x.is_contiguous()
        y = torch.empty([x.size(0), x.size(1), yh, yw], dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        so = None
        s = si
        sw_active = 0
        if writeSigns:
            sw_active = yw * down - (down - 1) + fdt_w
            sh = yh * down - (down - 1) + fdt_h
            sw = (sw_active + 15) & ~15
            s = so = torch.empty([x.size(0), x.size(1), sh, sw >> 2], dtype=torch.uint8, device=x.device)
        elif read_signs:
            sw_active = s.size(3) << 2
        if read_signs or writeSigns:
            _check(s.is_contiguous(), 'signs must be contiguous')
            _check(s.dtype == torch.uint8, 'signs must be uint8')
            _check(s.device == x.device, 'signs must reside on the same device as x')
            _check(s.dim() == 4, 'signs must be rank 4')
            _check(s.size(0) == x.size(0) and s.size(1) == x.size(1), 'signs must have same batch & channels as x')
        p.x, p.y, p.b, p.fu, p.fd = _ptr(x), _ptr(y), _ptr(b), _ptr(fu), _ptr(fd)
        p.s = _ptr(s) if (read_signs or writeSigns) else None
        p.x_stride_w, p.x_stride_h, p.x_stride_c, p.x_stride_n = x.stride(3), x.stride(2), x.stride(1), x.stride(0)
        p.y_w, p.y_h = yw, yh
        p.y_stride_w, p.y_stride_h, p.y_stride_c, p.y_stride_n = y.stride(3), y.stride(2), y.stride(1), y.stride(0)
        p.b_stride = b.stride(0)
        p.s_w_bytes, p.s_h = (s.size(3), s.size(2)) if (read_signs or writeSigns) else (0, 0)
        p.s_ofs_x, p.s_ofs_y = int(sx), int(sy)
        p.s_w_active = int(sw_active)
        if y_sum is not None:
            _check(y_sum.dtype == torch.float32 and y_sum.is_contiguous() and y_sum.device == x.device and y_sum.numel() == x.size(1),
                   'y_sum must be a contiguous float32 [C] tensor on the device of x')
            p.y_sum = _ptr(y_sum)
        with torch.cuda.device(x.device):
            st = lib.vfm_filtered_lrelu(C.byref(p), _stream(x))
        if st == _lib.VFM_ERR_NO_KERNEL:
            return None, None, -1
        _lib.check(st, 'filtered_lrelu')
        return y, so, 0

    @staticmethod
    def filtered_lrelu_act_(x, si, sx, sy, gain, slope, clamp, writeSigns):
        _check(x.is_cuda, 'x must reside on CUDA device')
        _check(x.dim() == 4, 'x must be rank 4')
        _check(x.numel() > 0, 'x is empty')
        _check(x.dtype in _DT, 'x must be float16, float32 or float64')
        so = None
        s = si
        read_signs = s.numel() > 0
        if writeSigns:
            sw = (x.size(3) + 15) & ~15
            s = so = torch.empty([x.size(0), x.size(1), x.size(2), sw >> 2], dtype=torch.uint8, device=x.device)
        if read_signs or writeSigns:
            _check(s.is_contiguous(), 'signs must be contiguous')
            _check(s.dtype == torch.uint8, 'signs must be uint8')
            _check(s.device == x.device, 'signs must reside on the same device as x')
            _check(s.dim() == 4, 'signs must be rank 4')
            _check(s.size(0) == x.size(0) and s.size(1) == x.size(1), 'signs must have same batch & channels as x')
        p = _lib.FilteredLreluActParams()
        p.x = _ptr(x)
        p.s = _ptr(s) if (read_signs or writeSigns) else None
        p.dtype = _DT[x.dtype]
        p.gain, p.slope, p.clamp = float(gain), float(slope), float(clamp)
        p.write_signs, p.read_signs = int(bool(writeSigns)), int(read_signs and not writeSigns)
        p.x_w, p.x_h, p.channels, p.batch = x.size(3), x.size(2), x.size(1), x.size(0)
        p.x_stride_w, p.x_stride_h, p.x_stride_c, p.x_stride_n = x.stride(3), x.stride(2), x.stride(1), x.stride(0)
        p.s_w, p.s_h = (s.size(3) << 2, s.size(2)) if (read_signs or writeSigns) else (0, 0)
        p.s_ofs_x, p.s_ofs_y = int(sx), int(sy)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().vfm_filtered_lrelu_act(C.byref(p), _stream(x)), 'filtered_lrelu_act_')
        return so


class _ModconvPlugin:
    """New native boundary for networks/generator.py:46 modulated_conv2d."""

    @staticmethod
    def _desc(x, weight, up, padding, demodulate, flip_weight, noise, resample_filter, force_generic):
        d = _lib.ModconvDesc()
        d.dtype = _dtype_code(x, 'modulated_conv2d')
        d.batch, d.in_channels, d.in_h, d.in_w = x.shape
        d.out_channels, _, d.kh, d.kw = weight.shape
        d.up, d.padding = int(up), int(padding)
        d.demodulate, d.flip_weight = int(bool(demodulate)), int(bool(flip_weight))
        if noise is None:
            d.noise_mode = _lib.NOISE_NONE
        elif noise.dim() == 2:
            d.noise_mode = _lib.NOISE_HW
        else:
            d.noise_mode = _lib.NOISE_N1HW
        d.resample_filter = _ptr(resample_filter) if up > 1 else None
        d.fw, d.fh = (resample_filter.size(1), resample_filter.size(0)) if (up > 1 and resample_filter is not None) else (1, 1)
        if up == 1:
            d.out_h = d.in_h + 2 * padding - d.kh + 1
            d.out_w = d.in_w + 2 * padding - d.kw + 1
        else:
            px0, px1 = padding + (d.fw + 1) // 2, padding + (d.fw - 2) // 2
            py0, py1 = padding + (d.fh + 1) // 2, padding + (d.fh - 2) // 2
            d.out_w = d.in_w * 2 + px0 + px1 - d.fw + 1 - (d.kw - 1)
            d.out_h = d.in_h * 2 + py0 + py1 - d.fh + 1 - (d.kh - 1)
        d.force_generic = int(bool(force_generic))
        return d

    @staticmethod
    def _common_checks(x, weight, styles, noise, up, resample_filter):
        _check(x.is_cuda, 'x must reside on CUDA device')
        _check(x.dim() == 4 and weight.dim() == 4 and styles.dim() == 2, 'x must be [N,I,H,W], weight [O,I,kh,kw], styles [N,I]')
        _check(weight.size(1) == x.size(1) and tuple(styles.shape) == (x.size(0), x.size(1)), 'shape mismatch between x, weight and styles')
        _check(x.is_contiguous(), 'x must be contiguous (NCHW)')
        _check(weight.dtype == torch.float32 and weight.is_contiguous(), 'weight must be contiguous float32')
        _check(styles.dtype == torch.float32 and styles.is_contiguous(), 'styles must be contiguous float32')
        _check(weight.device == x.device and styles.device == x.device, 'all tensors must reside on the same device')
        _check(noise is None or (noise.dtype == torch.float32 and noise.is_contiguous() and noise.device == x.device), 'noise must be contiguous float32')
        _check(up == 1 or (resample_filter is not None and resample_filter.dim() == 2 and resample_filter.dtype == torch.float32
                           and resample_filter.is_contiguous() and resample_filter.device == x.device),
               'up=2 needs a contiguous float32 2-D resample filter on the device of x')

    @staticmethod
    def uses_tensor_cores(x, weight, up=1, padding=0, demodulate=True, flip_weight=True, noise=None, resample_filter=None):
        d = _ModconvPlugin._desc(x, weight, up, padding, demodulate, flip_weight, noise, resample_filter, False)
        return bool(_lib.load().vfm_modconv_uses_tensor_cores(C.byref(d)))

    @staticmethod
    def fused_backward_supported(x, weight, up=1, padding=0, demodulate=True, flip_weight=True, noise=None, resample_filter=None):
        """Can ``backward(epilogue=...)`` (fused training layer) take this call?"""
        d = _ModconvPlugin._desc(x, weight, up, padding, demodulate, flip_weight, noise, resample_filter, False)
        return bool(_lib.load().vfm_modconv_fused_backward_supported(C.byref(d)))

    @staticmethod
    def group_norm_affine(x, weight, bias, num_groups, eps=1e-5):
        """GroupNorm statistics of x [N,C,H,W] as the per-(sample, channel) affine map (scale, shift), both fp32 [N,C]:
        group_norm(x) == x * scale[:, :, None, None] + shift[:, :, None, None]  (nn.GroupNorm evaluated in fp32)."""
        _check(x.is_cuda and x.dim() == 4 and x.is_contiguous(), 'x must be a contiguous NCHW CUDA tensor')
        n, c = x.shape[0], x.shape[1]
        _check(c % int(num_groups) == 0, 'channels must be divisible by num_groups')
        gam = weight.detach().to(torch.float32).contiguous() if weight is not None else None
        bet = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        scale = torch.empty([n, c], dtype=torch.float32, device=x.device)
        shift = torch.empty([n, c], dtype=torch.float32, device=x.device)
        p = _lib.GroupNormAffineParams()
        p.x, p.gamma, p.beta, p.scale, p.shift = _ptr(x), _ptr(gam), _ptr(bet), _ptr(scale), _ptr(shift)
        p.dtype = _dtype_code(x, 'group_norm_affine')
        p.batch, p.channels, p.groups, p.hw, p.eps = n, c, int(num_groups), x.shape[2] * x.shape[3], float(eps)
        with torch.cuda.device(x.device):
            st = _lib.load().vfm_group_norm_affine(C.byref(p), _stream(x))
        _lib.check(st, 'group_norm_affine')
        return scale, shift

    @staticmethod
    def forward(x, weight, styles, noise, up, padding, resample_filter, demodulate, flip_weight, force_generic=False, epilogue=None,
                x_affine=None, keep_operand=False):
        """-> (y [N,O,Hout,Wout] in x.dtype, dcoefs [N,O] fp32)   [, saved] with ``keep_operand``.

        ``keep_operand`` (training): the tcgen05 path leaves its NHWC activation operand in the workspace; the third return value is then
        ``(workspace tensor, hi pointer, lo pointer)`` to hand to ``backward(saved_operand=...)`` (or None when the path keeps none).

        ``x_affine`` (inference only): (scale, shift) fp32 [N,I] from ``group_norm_affine``: the conv sees x*scale+shift; with
        ``epilogue['residual_affine']`` the same map is applied to the epilogue's residual (which then is the raw x).

        ``epilogue`` (inference only): dict(act='linear'|'lrelu'|'gelu', alpha, gain, clamp, bias, residual, gamma, res_scale) fused into the
        kernel that writes y; returns None instead of a tuple if no kernel can fuse it for this call (caller composes)."""
        _ModconvPlugin._common_checks(x, weight, styles, noise, up, resample_filter)
        d = _ModconvPlugin._desc(x, weight, up, padding, demodulate, flip_weight, noise, resample_filter, force_generic)
        if noise is not None:
            exp = (d.out_h, d.out_w) if noise.dim() == 2 else (d.batch, 1, d.out_h, d.out_w)
            _check(tuple(noise.shape) == exp, f'noise must have shape {exp}, got {tuple(noise.shape)}')
        lib = _lib.load()
        y = torch.empty([d.batch, d.out_channels, d.out_h, d.out_w], dtype=x.dtype, device=x.device)
        dcoefs = torch.empty([d.batch, d.out_channels], dtype=torch.float32, device=x.device)
        nbytes = lib.vfm_modconv_workspace_bytes(C.byref(d), 0)
        ws = torch.empty([nbytes], dtype=torch.uint8, device=x.device)
        p = _lib.ModconvFwdParams()
        p.d = d
        p.x, p.weight, p.styles, p.noise, p.y, p.dcoefs = _ptr(x), _ptr(weight), _ptr(styles), _ptr(noise), _ptr(y), _ptr(dcoefs)
        p.workspace, p.workspace_bytes = _ptr(ws), nbytes
        keep = []
        if epilogue is not None:
            bias, res, gamma = epilogue.get('bias'), epilogue.get('residual'), epilogue.get('gamma')
            _check(bias is None or (bias.dtype == x.dtype and bias.is_contiguous() and bias.numel() == d.out_channels), 'epilogue bias must be [O] in the dtype of x')
            _check(res is None or (res.dtype == x.dtype and res.is_contiguous() and tuple(res.shape) == tuple(y.shape)), 'epilogue residual must match y')
            if gamma is not None:
                gamma = gamma.detach().to(torch.float32).reshape(-1).contiguous()
                _check(gamma.numel() == d.out_channels, 'epilogue gamma must have O elements')
            keep = [bias, res, gamma]
            p.ep_enable = 1
            p.ep_act = {'linear': 1, 'lrelu': 3, 'gelu': 10}[epilogue.get('act', 'linear')]
            p.ep_alpha = float(epilogue.get('alpha', 0.2))
            p.ep_gain = float(epilogue.get('gain', 1.0))
            clamp = epilogue.get('clamp')
            p.ep_clamp = float(clamp) if clamp is not None else -1.0
            p.ep_bias, p.ep_residual, p.ep_gamma = _ptr(bias), _ptr(res), _ptr(gamma)
            p.ep_res_scale = float(epilogue.get('res_scale', 1.0))
            p.ep_res_affine = int(bool(epilogue.get('residual_affine', False)))
        if x_affine is not None:
            xs, xb = x_affine
            _check(all(t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (d.batch, d.in_channels) and t.device == x.device
                       for t in (xs, xb)), 'x_affine must be two contiguous fp32 [N,I] tensors')
            keep += [xs, xb]
            p.x_scale, p.x_shift = _ptr(xs), _ptr(xb)
        p.keep_operand = int(bool(keep_operand))
        with torch.cuda.device(x.device):
            st = lib.vfm_modconv_forward(C.byref(p), _stream(x))
        del keep
        if (epilogue is not None or x_affine is not None) and st == _lib.VFM_ERR_NO_KERNEL:
            return None
        _lib.check(st, 'modulated_conv2d')
        if keep_operand:
            hi, lo = C.c_void_p(), C.c_void_p()
            _lib.check(lib.vfm_modconv_forward_operand(C.byref(d), _ptr(ws), nbytes, C.byref(hi), C.byref(lo)), 'modconv_forward_operand')
            return y, dcoefs, ((ws, hi.value, lo.value) if hi.value else None)
        return y, dcoefs

    @staticmethod
    def backward(dy, x, y, weight, styles, noise, dcoefs, up, padding, resample_filter, demodulate, flip_weight,
                 need_dx=True, need_dweight=True, need_dstyles=True, need_dnoise=False, force_generic=False, saved_operand=None, epilogue=None):
        """-> (dx or None, dweight fp32 or None, dstyles fp32 or None, dnoise fp32 or None)   [, dbias_no fp32 [N,O]] with ``epilogue``.

        ``saved_operand``: what ``forward(keep_operand=True)`` returned for the same x / styles (its workspace must still be alive and untouched).
        ``epilogue`` (fused training layer): dict(act, alpha, gain, clamp, bias) of the forward's fused epilogue; ``dy`` / ``y`` then refer to the
        ACTIVATED output and the bias_act backward is folded into the kernels (the fifth return value is the per-sample bias gradient)."""
        _ModconvPlugin._common_checks(x, weight, styles, noise, up, resample_filter)
        d = _ModconvPlugin._desc(x, weight, up, padding, demodulate, flip_weight, noise, resample_filter, force_generic)
        _check(dy.is_contiguous() and dy.dtype == x.dtype and tuple(dy.shape) == (d.batch, d.out_channels, d.out_h, d.out_w),
               'dy must be contiguous, of the dtype of x and of the shape of the forward output')
        _check(not demodulate or (y is not None and y.is_contiguous() and y.shape == dy.shape and y.dtype == x.dtype), 'y must match dy')
        lib = _lib.load()
        need_dx = need_dx or need_dstyles
        dx = torch.empty_like(x) if need_dx else None
        dweight = torch.empty_like(weight) if need_dweight else None
        dstyles = torch.empty_like(styles) if need_dstyles else None
        dnoise = torch.empty_like(noise) if (need_dnoise and noise is not None) else None
        nbytes = lib.vfm_modconv_workspace_bytes(C.byref(d), 2 if epilogue is not None else 1)
        ws = torch.empty([nbytes], dtype=torch.uint8, device=x.device)
        p = _lib.ModconvBwdParams()
        p.d = d
        p.dy, p.x, p.y, p.weight, p.styles, p.noise, p.dcoefs = _ptr(dy), _ptr(x), _ptr(y), _ptr(weight), _ptr(styles), _ptr(noise), _ptr(dcoefs)
        p.dx, p.dweight, p.dstyles, p.dnoise = _ptr(dx), _ptr(dweight), _ptr(dstyles), _ptr(dnoise)
        p.workspace, p.workspace_bytes = _ptr(ws), nbytes
        if saved_operand is not None:
            p.saved_operand, p.saved_operand_lo = saved_operand[1], saved_operand[2]
        dbias = ebias = None
        if epilogue is not None:
            ebias = epilogue.get('bias')
            _check(ebias is None or (ebias.dtype == x.dtype and ebias.is_contiguous() and ebias.numel() == d.out_channels), 'epilogue bias must be [O] in the dtype of x')
            _check(y is not None and y.is_contiguous() and y.shape == dy.shape and y.dtype == x.dtype, 'the fused backward needs the activated output y')
            dbias = torch.empty([d.batch, d.out_channels], dtype=torch.float32, device=x.device)
            p.ep_enable = 1
            p.ep_act = {'linear': 1, 'lrelu': 3}[epilogue.get('act', 'linear')]
            p.ep_alpha, p.ep_gain = float(epilogue.get('alpha', 0.2)), float(epilogue.get('gain', 1.0))
            clamp = epilogue.get('clamp')
            p.ep_clamp = float(clamp) if clamp is not None else -1.0
            p.ep_bias, p.dbias_no = _ptr(ebias), _ptr(dbias)
        with torch.cuda.device(x.device):
            _lib.check(lib.vfm_modconv_backward(C.byref(p), _stream(x)), 'modulated_conv2d backward')
        if epilogue is not None:
            return dx, dweight, dstyles, dnoise, dbias
        return dx, dweight, dstyles, dnoise


bias_act_plugin = _BiasActPlugin()
upfirdn2d_plugin = _Upfirdn2dPlugin()
filtered_lrelu_plugin = _FilteredLreluPlugin()
modconv_plugin = _ModconvPlugin()

PLUGINS = {
    'bias_act_plugin': bias_act_plugin,
    'upfirdn2d_plugin': upfirdn2d_plugin,
    'filtered_lrelu_plugin': filtered_lrelu_plugin,
    'modconv_plugin': modconv_plugin,
}
