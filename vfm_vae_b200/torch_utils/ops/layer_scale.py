"""Layer-scaled residual of the residual synthesis layers, ``(gamma * y + x) * scale`` (networks/generator.py:272-274:
``y = (self.gamma * y).to(dtype).add_(x).mul(np.sqrt(2))``), forward and backward in one elementwise pass each
(``vfm_rows_affine``) plus one row-wise dot product for the gradient of gamma (``vfm_rows_dot``) -- instead of three
broadcasting PyTorch kernels and their autograd graph (an fp32 round trip of the full activation at 256x256)."""
import ctypes as C

import torch

from ... import _lib
from ...plugins import _dtype_code, _ptr, _stream


def _rows(a, b, P, Q, R, out, rows, hw):
    p = _lib.RowsParams()
    p.a, p.b, p.P, p.Q, p.R, p.out = _ptr(a), _ptr(b), _ptr(P), _ptr(Q), _ptr(R), _ptr(out)
    p.dtype, p.rows, p.hw = _dtype_code(a, 'layer_scale_residual'), rows, hw
    return p


class _LayerScaleResidual(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, x, gamma, scale):
        n, c = y.shape[0], y.shape[1]
        hw = y[0, 0].numel()
        g32 = gamma.detach().to(torch.float32).reshape(-1)
        P = (g32 * scale).repeat(n).contiguous()                      # [N*C]
        Q = torch.full([n * c], float(scale), dtype=torch.float32, device=y.device)
        R = torch.zeros([n * c], dtype=torch.float32, device=y.device)
        out = torch.empty_like(y)
        lib = _lib.load()
        with torch.cuda.device(y.device):
            _lib.check(lib.vfm_rows_affine(C.byref(_rows(y, x, P, Q, R, out, n * c, hw)), _stream(y)), 'layer_scale_residual')
        ctx.save_for_backward(y, P, Q, R)
        ctx.cfg = (float(scale), gamma.shape, gamma.dtype)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        y, P, Q, R = ctx.saved_tensors
        scale, gshape, gdtype = ctx.cfg
        n, c = y.shape[0], y.shape[1]
        hw = y[0, 0].numel()
        dout = dout.contiguous()
        lib = _lib.load()
        dy = dx = dgamma = None
        with torch.cuda.device(y.device):
            if ctx.needs_input_grad[0]:
                dy = torch.empty_like(y)
                _lib.check(lib.vfm_rows_affine(C.byref(_rows(dout, None, P, None, R, dy, n * c, hw)), _stream(y)), 'layer_scale_residual backward')
            if ctx.needs_input_grad[1]:
                dx = torch.empty_like(y)
                _lib.check(lib.vfm_rows_affine(C.byref(_rows(dout, None, Q, None, R, dx, n * c, hw)), _stream(y)), 'layer_scale_residual backward')
            if ctx.needs_input_grad[2]:
                dots = torch.empty([n, c], dtype=torch.float32, device=y.device)
                _lib.check(lib.vfm_rows_dot(C.byref(_rows(dout, y, None, None, None, dots, n * c, hw)), _stream(y)), 'layer_scale_residual backward')
                dgamma = (dots.sum(0) * scale).reshape(gshape).to(gdtype)
        return dy, dx, dgamma, None


def supported(y, x, gamma):
    return (y.is_cuda and y.dim() == 4 and y.dtype in (torch.float16, torch.float32) and x.dtype == y.dtype and x.shape == y.shape
            and y.is_contiguous() and x.is_contiguous() and gamma.numel() == y.shape[1])


def layer_scale_residual(y, x, gamma, scale):
    """``((gamma * y).to(y.dtype) + x) * scale`` with gamma ``[1,C,1,1]`` (fp32 parameter)."""
    return _LayerScaleResidual.apply(y, x, gamma, float(scale))
