"""Filtered leaky ReLU on sm_100a: ``y = down_fd( clamp( lrelu( up_fu(x + b) * up^2 * gain ), +-clamp ) )`` in one kernel.

Public entry point and semantics are the reference's (torch_utils/ops/filtered_lrelu.py:56-116): same signature, same output size
``(in*up + pad0 + pad1 - (fu-1) - (fd-1) + (down-1)) // down``, the reference's 2-bit sign tensor handed from forward to backward,
and the same "no fused kernel -> compose upfirdn2d + activation + upfirdn2d" contract.

How this file differs from the reference wrapper:

* ONE autograd Function parameterised by an immutable ``_Op`` record (the reference builds and caches a Function class per
  parameter tuple).  The backward pass is the same Function applied to the transposed record (up <-> down, fu <-> fd, flipped
  filters, gain * up^2 / down^2, no clamp, signs read at the shifted offset), so derivatives of any order work.
* the bias gradient is fused: the backward launch accumulates ``sum(dx)`` per channel inside the kernel (``y_sum`` of
  ``vfm_filtered_lrelu``) instead of a second pass ``dx.sum([0, 2, 3])`` over the gradient tensor (reference line 266).
* stream-safe: the filters are kernel arguments, not a global ``__constant__`` buffer, so the reference's warning about
  non-default streams (lines 215-216) has no counterpart; the fused envelope is wider (any separable / full mix up to 32 taps).
"""
import warnings
from collections import namedtuple

import numpy as np
import torch

from ... import custom_ops
from . import upfirdn2d

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='filtered_lrelu_plugin')
    return True


#: everything that is not a tensor: resampling factors, the four paddings, activation constants, filter orientation
_Op = namedtuple('_Op', 'up down px0 px1 py0 py1 gain slope clamp flip')


def _pad4(padding):
    """int | [x, y] | [x0, x1, y0, y1] -> (x0, x1, y0, y1)"""
    if isinstance(padding, (int, np.integer)):
        padding = [padding] * 4
    padding = [int(v) for v in padding]
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    assert len(padding) == 4, 'padding must be an int, [x, y] or [x0, x1, y0, y1]'
    return tuple(padding)


def _taps(f):
    """(taps along x, taps along y) of a 1-D (separable) or 2-D filter"""
    return f.shape[-1], f.shape[0]


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    """Arguments as in the reference (filtered_lrelu.py:56-110)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError('vfm_vae_b200.filtered_lrelu has no reference/CPU implementation: CUDA tensors only '
                           '(the CPU oracle is oracle/ref_ops.py, for tests).')
    assert x.ndim == 4
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert float(gain) > 0 and float(slope) >= 0 and (clamp is None or float(clamp) >= 0)
    _init()
    op = _Op(up, down, *_pad4(padding), float(gain), float(slope), float(clamp) if clamp is not None else float('inf'), bool(flip_filter))
    return _FilteredLRelu.apply(x, fu, fd, b, op, None, (0, 0), False)[0]


def _canonical_filter(f, factor, device):
    """None -> identity; a 1-tap separable filter without resampling becomes the equivalent 1x1 2-D filter (as the reference does)."""
    if f is None:
        return torch.ones([1, 1], dtype=torch.float32, device=device)
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    if factor == 1 and f.ndim == 1 and f.shape[0] == 1:
        return f.square()[None]
    return f


class _FilteredLRelu(torch.autograd.Function):
    """forward(x, fu, fd, b, op, signs, (sx, sy), want_sum) -> (y, y_sum)

    ``signs`` None: a training forward writes the sign tensor; given: this call is (part of) a backward pass and reads it at offset
    (sx, sy).  ``want_sum``: also return the per-channel sum of y (fp32), accumulated by the kernel."""

    @staticmethod
    def forward(ctx, x, fu, fd, b, op, signs, sofs, want_sum):
        fu = _canonical_filter(fu, op.up, x.device)
        fd = _canonical_filter(fd, op.down, x.device)
        bias = b if b is not None else torch.zeros([x.shape[1]], dtype=x.dtype, device=x.device)
        reading = signs is not None
        writing = (not reading) and (x.requires_grad or (b is not None and b.requires_grad))
        si = signs if reading else torch.empty([0])
        y_sum = torch.zeros([x.shape[1]], dtype=torch.float32, device=x.device) if want_sum else None
        pads = (op.px0, op.px1, op.py0, op.py1)
        y = so = None
        rc = -1
        if x.dtype in (torch.float16, torch.float32):
            y, so, rc = _plugin.filtered_lrelu(x, fu, fd, bias, si, op.up, op.down, *pads, sofs[0], sofs[1], op.gain, op.slope, op.clamp,
                                               op.flip, writing, y_sum=y_sum)
        if rc < 0:
            # outside the fused kernel's envelope: the reference's generic composition (filtered_lrelu.py:223-229)
            warnings.warn('filtered_lrelu called with parameters that have no fused CUDA kernel, composing it from upfirdn2d + activation', RuntimeWarning)
            y = upfirdn2d.upfirdn2d(x=x + bias.reshape(1, -1, 1, 1), f=fu, up=op.up, padding=list(pads), gain=op.up ** 2, flip_filter=op.flip)
            so = _plugin.filtered_lrelu_act_(y, si, sofs[0], sofs[1], op.gain, op.slope, op.clamp, writing)
            y = upfirdn2d.upfirdn2d(x=y, f=fd, down=op.down, flip_filter=op.flip)
            if want_sum:
                y_sum = y.sum([0, 2, 3])                     # in y's own dtype: fp64 inputs take this route and keep their precision
        ctx.save_for_backward(fu, fd, signs if reading else so)
        ctx.op, ctx.sofs = op, sofs
        ctx.x_hw, ctx.y_hw = tuple(x.shape[2:]), tuple(y.shape[2:])
        ctx.bias_dtype = b.dtype if b is not None else None
        if y_sum is None:
            y_sum = torch.empty([0], device=x.device)
        ctx.mark_non_differentiable(y_sum)
        return y, y_sum

    @staticmethod
    def backward(ctx, dy, _d_sum):
        fu, fd, signs = ctx.saved_tensors
        op = ctx.op
        need_x, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[3]
        assert not (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]), 'filtered_lrelu: no gradients for the filters (as in the reference)'
        dx = db = None
        if need_x or need_b:
            (xh, xw), (yh, yw) = ctx.x_hw, ctx.y_hw
            fuw, fuh = _taps(fu)
            fdw, fdh = _taps(fd)
            # adjoint op: swap the resampling factors and the filters, mirror the filters, and pad so that the result has x's size
            adj = _Op(up=op.down, down=op.up,
                      px0=(fuw - 1) + (fdw - 1) - op.px0, px1=xw * op.up - yw * op.down + op.px0 - (op.up - 1),
                      py0=(fuh - 1) + (fdh - 1) - op.py0, py1=xh * op.up - yh * op.down + op.py0 - (op.up - 1),
                      gain=op.gain * (op.up ** 2) / (op.down ** 2), slope=op.slope, clamp=float('inf'), flip=not op.flip)
            sofs = (ctx.sofs[0] - (fuw - 1) + op.px0, ctx.sofs[1] - (fuh - 1) + op.py0)
            # the kernel's fused sum is not differentiable: under create_graph (higher-order gradients) db is a plain reduction of dx
            fused_sum = bool(need_b) and not torch.is_grad_enabled()
            dx, dx_sum = _FilteredLRelu.apply(dy, fd, fu, None, adj, signs, sofs, fused_sum)
            if need_b:
                db = dx_sum.to(ctx.bias_dtype) if fused_sum else dx.sum([0, 2, 3])
        return dx, None, None, db, None, None, None, None
