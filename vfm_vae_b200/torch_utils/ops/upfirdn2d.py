"""Pad / upsample / FIR-filter / downsample of 2-D images on sm_100a.  Host-side mirror of the reference wrapper
torch_utils/ops/upfirdn2d.py (public functions ``setup_filter``, ``upfirdn2d``, ``filter2d``, ``upsample2d``,
``downsample2d`` with identical signatures and padding arithmetic; the gradient is the same op with up/down swapped
and the filter flipped, reference lines 251-269).  The kernels come from libvfmops.so; no ref/CPU path here."""
import numpy as np
import torch

from ... import custom_ops

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='upfirdn2d_plugin')
    return True


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, (int, np.integer)) for v in padding)
    padding = [int(v) for v in padding]
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Build the fp32 FIR filter tensor ([taps] separable or [h,w]); same rules as reference lines 70-114."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Arguments as in the reference (upfirdn2d.py:118-157)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError('vfm_vae_b200.upfirdn2d has no reference/CPU implementation: CUDA tensors only '
                           '(the CPU oracle is oracle/ref_ops.py, for tests).')
    _init()
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


_upfirdn2d_cuda_cache = dict()


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _upfirdn2d_cuda_cache:
        return _upfirdn2d_cuda_cache[key]

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            if f.ndim == 1 and f.shape[0] == 1:
                f = f.square().unsqueeze(0)   # separable 1-tap == full 1x1
            assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
            if f.ndim == 2:
                y = _plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:   # separable: a horizontal pass then a vertical pass, gain applied once
                y = _plugin.upfirdn2d(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, 1.0)
                y = _plugin.upfirdn2d(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, gain)
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _upfirdn2d_cuda_cache[key] = Upfirdn2dCuda
    return Upfirdn2dCuda


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


class _ReplicateBlur(torch.autograd.Function):
    """replicate-pad + fixed depthwise blur (k = 3, 5) as one pass of the streaming kernel; backward = one pass of the zero-padded
    stencil kernel over dy (exact away from the border) + ``vfm_replicate_blur_edges``, which rewrites the outermost row / column of every
    plane with the sums that include the taps the clamp folds back onto them."""

    @staticmethod
    def forward(ctx, x, f32):
        k = f32.shape[0]
        ctx.f = f32
        return _plugin.upfirdn2d(x, f32, 1, 1, 1, 1, k // 2, k // 2, k // 2, k // 2, True, 1.0, pad_mode=1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        f32 = ctx.f
        k, c = f32.shape[0], dy.shape[1]
        dy = dy.contiguous()
        fc = f32[None].repeat(c, 1, 1)
        dx = _dw_call(dy, fc, k, False)
        return _plugin.replicate_blur_edges(dy, dx, f32), None


def blur2d_replicate(x, f, padding):
    """``F.conv2d(F.pad(x, padding, mode='replicate'), f[None, None].repeat(C, 1, 1, 1), groups=C)`` for a fixed 2-D kernel ``f``
    (<= 5x5) whose padding keeps the size -- the blur behind the pixel-shuffle upsampler (networks/utils/convnext_utils.py:250-255) --
    as one pass of the streaming blur kernel with clamp-to-edge addressing.  ``padding`` = (left, right, top, bottom) as for ``F.pad``.
    With gradients (odd square kernels, k = 3, 5) it is an autograd function whose backward is one pass of the zero-padded stencil
    kernel over dy plus the folding of the pad rows / columns onto the edges.  Returns None when the kernels do not apply (caller
    composes the stock ops)."""
    if x.device.type != 'cuda' or x.dtype not in (torch.float16, torch.float32):
        return None
    pl, pr, pt, pb = [int(v) for v in padding]
    fh, fw = f.shape
    if pl + pr != fw - 1 or pt + pb != fh - 1 or pl != pt or not x.is_contiguous():
        return None
    _init()
    f32 = f.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if torch.is_grad_enabled() and x.requires_grad:
        if (fh != fw or fh not in (3, 5) or pl != pr or pt != pb or x.dim() != 4 or min(x.shape[2:]) < fh
                or (x.shape[3] * x.element_size()) % 16 != 0 or x.shape[3] % 8 != 0):
            return None
        return _ReplicateBlur.apply(x, f32)
    return _plugin.upfirdn2d(x, f32, 1, 1, 1, 1, pl, pr, pt, pb, True, 1.0, pad_mode=1)


def _dw_call(x, f, k, flip, bias=None, add=None):
    return _plugin.upfirdn2d(x, f, 1, 1, 1, 1, k // 2, k // 2, k // 2, k // 2, flip, 1.0, add=add, bias=bias)


class _DepthwiseConv2d(torch.autograd.Function):
    """nn.Conv2d(C, C, k, padding=k//2, groups=C): forward and data gradient on the streaming stencil kernel (the data gradient is the
    same conv with the taps flipped), weight / bias gradient by vfm_depthwise_wgrad.  First-order autograd."""

    @staticmethod
    def forward(ctx, x, weight, bias, noise):
        k = weight.shape[2]
        f = weight.detach().to(torch.float32).reshape(weight.shape[0], k, k).contiguous()
        b = bias.detach().to(x.dtype).contiguous() if bias is not None else None
        add = noise.detach().to(torch.float32).reshape(x.shape[2], x.shape[3]).contiguous() if noise is not None else None
        y = _dw_call(x, f, k, True, bias=b, add=add)
        if y is None:
            raise RuntimeError('depthwise_conv2d: no kernel for this input (checked by the caller)')
        ctx.save_for_backward(x, f)
        ctx.cfg = (k, weight.shape, weight.dtype, bias.dtype if bias is not None else None,
                   (noise.shape, noise.dtype) if noise is not None else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, f = ctx.saved_tensors
        k, wshape, wdt, bdt, nz = ctx.cfg
        dy = dy.contiguous()
        dx = dw = db = dn = None
        if ctx.needs_input_grad[0]:
            dx = _dw_call(dy, f, k, False)
        if ctx.needs_input_grad[1] or (bdt is not None and ctx.needs_input_grad[2]):
            out = _plugin.depthwise_wgrad(x, dy, k, bdt is not None)
            if out is None:
                raise RuntimeError('depthwise_conv2d: no weight-gradient kernel for this input (checked by the caller)')
            dw = out[0].reshape(wshape).to(wdt) if ctx.needs_input_grad[1] else None
            db = out[1].to(bdt) if (bdt is not None and ctx.needs_input_grad[2]) else None
        if nz is not None and ctx.needs_input_grad[3]:
            dn = dy.sum(dim=(0, 1), dtype=torch.float32).reshape(nz[0]).to(nz[1])       # the addend is broadcast over samples and channels
        return dx, dw, db, dn


class _PixelShuffle2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _plugin.pixel_shuffle2(x, False)

    @staticmethod
    def backward(ctx, dy):
        return _plugin.pixel_shuffle2(dy.contiguous(), True)


def depthwise_conv2d(x, weight, bias=None, noise=None):
    """``F.conv2d(x, weight, bias, padding=k // 2, groups=C) (+ noise [1,1,H,W])`` for ``weight`` [C,1,k,k], k in {3, 5, 7}, fp16 / fp32
    contiguous NCHW CUDA ``x`` -- the dwconv of the ConvNeXt synthesis layers (networks/utils/convnext_utils.py:99,128) and of
    SeparableUpsampleWithFixedBlur -- on the streaming stencil kernel with the channel's taps in registers.  The noise add is
    folded into the kernel (one rounding of conv + bias + noise to ``x.dtype``; the reference rounds the conv output to fp16 under
    autocast and then adds the fp32 noise); with gradients the op is an autograd function (data gradient on the same kernel, weight /
    bias gradient by ``vfm_depthwise_wgrad``, noise gradient = the sum of dy over samples and channels).  Returns None when the kernels do
    not apply (the caller uses the stock module)."""
    if (x.device.type != 'cuda' or x.dtype not in (torch.float16, torch.float32) or x.dim() != 4 or weight.dim() != 4 or weight.shape[1] != 1
            or weight.shape[0] != x.shape[1] or weight.shape[2] != weight.shape[3] or weight.shape[2] not in (3, 5, 7) or not x.is_contiguous()
            or (x.shape[3] * x.element_size()) % 16 != 0 or x.numel() == 0):
        return None
    _init()
    k = weight.shape[2]
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)
                                              or (noise is not None and noise.requires_grad))
    if needs_grad:
        if x.shape[3] % 8 != 0 or x.shape[3] > 1024 or x.shape[0] > 65535 or x.shape[1] > 65535:
            return None
        return _DepthwiseConv2d.apply(x, weight, bias, noise)
    f = weight.detach().to(torch.float32).reshape(weight.shape[0], k, k).contiguous()
    b = bias.detach().to(x.dtype).contiguous() if bias is not None else None
    add = noise.detach().to(torch.float32).reshape(x.shape[2], x.shape[3]).contiguous() if noise is not None else None
    return _dw_call(x, f, k, True, bias=b, add=add)


def pixel_shuffle2(x):
    """``F.pixel_shuffle(x, 2)`` for contiguous fp16 / fp32 NCHW CUDA tensors with W % 4 == 0 (the upsampling step of
    SeparableUpsampleWithFixedBlur, networks/utils/convnext_utils.py:197-257) as a vectorised copy at the HBM rate; the backward is
    the inverse permutation on the same kernel.  Returns None when the kernel does not apply (the caller uses the stock op)."""
    if (x.device.type != 'cuda' or x.dtype not in (torch.float16, torch.float32) or x.dim() != 4 or x.shape[1] % 4 != 0 or x.shape[3] % 4 != 0
            or not x.is_contiguous() or x.numel() == 0):
        return None
    _init()
    if torch.is_grad_enabled() and x.requires_grad:
        return _PixelShuffle2.apply(x)
    return _plugin.pixel_shuffle2(x)
