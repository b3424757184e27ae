"""GroupNorm32 (networks/utils/shared.py: ``nn.GroupNorm`` evaluated in fp32, result in ``x.dtype``) on sm_100a, forward and
backward, as one statistics pass + one elementwise pass each way (``vfm_group_norm_forward`` / ``_backward``) instead of the
reference's cast -> moments -> normalise -> cast chain.  Same semantics: biased variance, eps inside the square root, affine
parameters in fp32, fp16 or fp32 activations (contiguous NCHW).  First-order autograd only."""
import ctypes as C

import torch

from ... import _lib
from ...plugins import _check, _dtype_code, _ptr, _stream


def _params(x, weight, bias, groups, eps, mean, rstd, scratch):
    p = _lib.GroupNormParams()
    p.x, p.gamma, p.beta = _ptr(x), _ptr(weight), _ptr(bias)
    p.mean, p.rstd, p.scratch = _ptr(mean), _ptr(rstd), _ptr(scratch)
    p.dtype = _dtype_code(x, 'group_norm')
    p.batch, p.channels, p.groups, p.hw, p.eps = x.shape[0], x.shape[1], int(groups), x[0, 0].numel(), float(eps)
    return p


class _GroupNorm32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        n, c = x.shape[0], x.shape[1]
        w32 = weight.detach().to(torch.float32).contiguous() if weight is not None else None
        b32 = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        y = torch.empty_like(x)
        mean = torch.empty([n, groups], dtype=torch.float32, device=x.device)
        rstd = torch.empty([n, groups], dtype=torch.float32, device=x.device)
        scratch = torch.empty([3, n, c], dtype=torch.float32, device=x.device)
        p = _params(x, w32, b32, groups, eps, mean, rstd, scratch)
        p.y = _ptr(y)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().vfm_group_norm_forward(C.byref(p), _stream(x)), 'group_norm')
        ctx.save_for_backward(x, w32 if w32 is not None else torch.empty([0]), mean, rstd)
        ctx.cfg = (groups, eps, weight is not None, bias is not None, weight.dtype if weight is not None else None,
                   bias.dtype if bias is not None else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w32, mean, rstd = ctx.saved_tensors
        groups, eps, has_w, has_b, wdt, bdt = ctx.cfg
        n, c = x.shape[0], x.shape[1]
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = torch.empty([n, c], dtype=torch.float32, device=x.device)
        db = torch.empty([n, c], dtype=torch.float32, device=x.device)
        scratch = torch.empty([3, n, c], dtype=torch.float32, device=x.device)
        p = _params(x, w32 if has_w else None, None, groups, eps, mean, rstd, scratch)
        p.dy, p.dx, p.dgamma_nc, p.dbeta_nc = _ptr(dy), _ptr(dx), _ptr(dg), _ptr(db)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().vfm_group_norm_backward(C.byref(p), _stream(x)), 'group_norm backward')
        dw = dg.sum(0).to(wdt) if (has_w and ctx.needs_input_grad[1]) else None
        dbias = db.sum(0).to(bdt) if (has_b and ctx.needs_input_grad[2]) else None
        return (dx if ctx.needs_input_grad[0] else None), dw, dbias, None, None


def supported(x, num_groups):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.float32) and x.is_contiguous()
            and x.shape[1] % num_groups == 0 and x.shape[1] // num_groups <= 64 and x.numel() > 0)


def group_norm32(x, num_groups, weight=None, bias=None, eps=1e-5):
    """``GroupNorm32(num_groups, C)(x)`` of the reference: statistics and affine in fp32, output in ``x.dtype``."""
    _check(supported(x, num_groups), 'group_norm32: needs a contiguous fp16/fp32 NCHW CUDA tensor with <= 64 channels per group')
    return _GroupNorm32.apply(x, weight, bias, int(num_groups), float(eps))
