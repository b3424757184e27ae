"""Fused bias + activation (+ gain, clamp) on sm_100a.  Host-side mirror of the reference wrapper
torch_utils/ops/bias_act.py:52-207: same public function, same defaults table, same autograd structure (forward
Function + a gradient Function that is itself differentiable once), with two differences that do not change results:

* the kernels come from libvfmops.so via ``custom_ops.get_plugin('bias_act_plugin')`` instead of a JIT build;
* the bias gradient is produced by the same kernel launch that computes ``dx`` (warp-shuffle/block reduction, fp32
  accumulation) instead of a separate ``dx.sum(...)`` pass over HBM (reference: bias_act.py:170).

There is no ``impl='ref'`` path and no CPU path here: the oracle lives in ``oracle/`` and is test-only.
"""
from types import SimpleNamespace

import numpy as np
import torch

from ... import custom_ops

activation_funcs = {
    'linear':   SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     SimpleNamespace(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    SimpleNamespace(def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': SimpleNamespace(def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    SimpleNamespace(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

_plugin = None
_null_tensor = torch.empty([0])


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='bias_act_plugin')
    return True


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain, +-clamp).  Arguments as in the reference (bias_act.py:52-82)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError('vfm_vae_b200.bias_act has no reference/CPU implementation: it runs the sm_100a kernel on CUDA '
                           'tensors only (the CPU oracle is oracle/ref_ops.py, for tests).')
    _init()
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


_bias_act_cuda_cache = dict()

#: Gradient of a clamped *linear* bias_act (ToRGB with conv_clamp) where |y| >= clamp.  The reference disagrees with itself here:
#: its impl='ref' path (``x.clamp``, bias_act.py:118-119) zeroes it, its CUDA path keeps y only for activations whose derivative is
#: written in terms of y (bias_act.py:151-154) and therefore lets it through.  'ref' (default) follows impl='ref' -- the path
#: north_star defines parity against and the golden vectors pin; 'cuda' reproduces the reference's CUDA training behaviour, which is
#: also what the reference's own wrapper does when it runs on these kernels through integration.install().
linear_clamp_grad = 'ref'


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    assert linear_clamp_grad in ('ref', 'cuda')
    key = (dim, act, alpha, gain, clamp, linear_clamp_grad)
    if key in _bias_act_cuda_cache:
        return _bias_act_cuda_cache[key]

    trivial = (act == 'linear' and gain == 1 and clamp < 0)
    needs_x = ('x' in spec.ref) or spec.has_2nd_grad
    # The clamp mask of the gradient needs y; see ``linear_clamp_grad`` above for the one case where the reference's two paths differ.
    needs_y = ('y' in spec.ref) or (clamp >= 0 and 'x' not in spec.ref and linear_clamp_grad == 'ref')

    def _mem_format(t):
        return torch.channels_last if t.ndim > 2 and t.stride(1) == 1 else torch.contiguous_format

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = _mem_format(x)
            x = x.contiguous(memory_format=ctx.memory_format)
            b = b.contiguous() if b is not None else _null_tensor
            y = x
            if not trivial or b is not _null_tensor:
                y = _plugin.bias_act(x, b, _null_tensor, _null_tensor, _null_tensor, 0, dim, spec.cuda_idx, alpha, gain, clamp)
            ctx.save_for_backward(x if needs_x else _null_tensor, b if needs_x else _null_tensor,
                                  y if needs_y else _null_tensor)
            ctx.b_numel = b.numel()
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                if trivial:
                    dx = dy
                    if ctx.needs_input_grad[1]:
                        db = dx.sum([i for i in range(dx.ndim) if i != dim])
                else:
                    # fused fp32-accumulated db for fp16/fp32; fp64 (test-only dtype) keeps full precision via a plain sum
                    fuse_db = bool(ctx.needs_input_grad[1]) and dy.dtype != torch.float64
                    dx, db = BiasActCudaGrad.apply(dy, x, b, y, fuse_db)
                    if not ctx.needs_input_grad[1]:
                        db = None
                    elif not fuse_db:
                        db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y, want_db):
            ctx.memory_format = _mem_format(dy)
            db32 = torch.zeros([dy.shape[dim]], dtype=torch.float32, device=dy.device) if want_db else None
            dx = _plugin.bias_act(dy, b, x, y, _null_tensor, 1, dim, spec.cuda_idx, alpha, gain, clamp, db=db32)
            ctx.save_for_backward(dy if spec.has_2nd_grad else _null_tensor, x, b, y)
            ctx.want_db = want_db
            db = db32.to(dy.dtype) if want_db else _null_tensor.to(dy.device)
            ctx.mark_non_differentiable(*([] if want_db else [db]))
            return dx, db

        @staticmethod
        def backward(ctx, d_dx, d_db):
            dy, x, b, y = ctx.saved_tensors
            # db = sum(dx): its cotangent broadcasts back onto dx's
            if ctx.want_db and d_db is not None:
                shape = [1] * d_dx.ndim
                shape[dim] = -1
                d_dx = d_dx + d_db.reshape(shape)
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy, _ = BiasActCudaGrad.apply(d_dx, x, b, y, False)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                fuse_db = bool(ctx.needs_input_grad[2]) and d_dx.dtype != torch.float64
                db32 = torch.zeros([d_dx.shape[dim]], dtype=torch.float32, device=d_dx.device) if fuse_db else None
                d_x = _plugin.bias_act(d_dx, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp, db=db32)
                if ctx.needs_input_grad[2]:
                    d_b = db32.to(d_dx.dtype) if fuse_db else d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None, None

    _bias_act_cuda_cache[key] = BiasActCuda
    return BiasActCuda
