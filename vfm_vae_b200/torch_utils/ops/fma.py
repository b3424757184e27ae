"""a * b + c with broadcasting-aware gradients.  Mirror of torch_utils/ops/fma.py:15 (only reached on the reference's
unfused modconv branch, which this package never takes; kept so callers importing it keep working).  Tested on CPU
(tests/test_fma_cpu.py, incl. gradcheck) and on the device (tests/test_ops_gpu.py)."""
import torch


def fma(a, b, c):
    return _Fma.apply(a, b, c)


def _reduce_to(g, shape):
    """Sum a gradient over the dimensions that were broadcast to reach g.shape from `shape` (leading dimensions that `shape` lacks,
    and size-1 dimensions of `shape` that g expanded)."""
    lead = g.ndim - len(shape)
    dims = [i for i in range(g.ndim) if i < lead or (shape[i - lead] == 1 and g.shape[i] > 1)]
    if dims:
        g = g.sum(dim=dims, keepdim=True)
    return g.reshape(shape)


class _Fma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return torch.addcmul(c, a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = _reduce_to(g * b, a.shape) if ctx.needs_input_grad[0] else None
        gb = _reduce_to(g * a, b.shape) if ctx.needs_input_grad[1] else None
        gc = _reduce_to(g, ctx.c_shape) if ctx.needs_input_grad[2] else None
        return ga, gb, gc
