"""Replicate-pad blur backward without a GPU: the oracle's model of ``vfm_replicate_blur_edges`` (border rows / columns as range sums of
the filter) against stock autograd of ``F.conv2d(F.pad(x, mode='replicate'), f, groups=C)`` (networks/utils/convnext_utils.py:250-255)
in fp64, with the interior ("core") supplied by ``conv_transpose2d`` in place of the CUDA stencil pass.  The CUDA kernel is held to the
same stock autograd in tests/test_ops_gpu.py::test_blur2d_replicate_autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle.ref_ops import replicate_blur_edges


@pytest.mark.parametrize('k', [3, 5])
@pytest.mark.parametrize('shape', [(2, 3, 9, 12), (1, 2, 5, 5), (2, 1, 16, 8), (1, 4, 7, 33)], ids=lambda s: 'x'.join(map(str, s)))
@pytest.mark.parametrize('symmetric', [True, False], ids=['binomial', 'random'])
def test_replicate_blur_edges_model_matches_autograd(k, shape, symmetric):
    g = torch.Generator().manual_seed(7)
    n, c, h, w = shape
    p = k // 2
    if symmetric:
        t = torch.tensor({3: [1, 2, 1], 5: [1, 4, 6, 4, 1]}[k], dtype=torch.float64)
        f = torch.outer(t, t) / t.sum() ** 2
    else:
        f = torch.randn(k, k, generator=g, dtype=torch.float64)
    fw = f[None, None].repeat(c, 1, 1, 1)
    x = torch.randn(shape, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(F.pad(x, (p, p, p, p), mode='replicate'), fw, groups=c)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (want,) = torch.autograd.grad(y, [x], dy)
    core = F.conv_transpose2d(dy, fw, groups=c)[:, :, p:h + p, p:w + p].clone()      # what the zero-padded stencil pass over dy computes
    got = replicate_blur_edges(dy, f, core)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-12 * max(1.0, want.abs().max().item())


def test_extension_ops_decline_cpu_tensors():
    """The inference / training extensions of the upsampler and ConvNeXt layers only apply to CUDA tensors: on CPU they return None
    (the decoder mirror then runs the stock module, which is what the CPU oracle runs use) -- they never compute on the CPU."""
    from vfm_vae_b200.torch_utils.ops import upfirdn2d as U
    x = torch.randn(1, 4, 8, 8)
    assert U.depthwise_conv2d(x, torch.randn(4, 1, 3, 3), None) is None
    assert U.pixel_shuffle2(x) is None
    assert U.blur2d_replicate(x, torch.ones(3, 3) / 9, (1, 1, 1, 1)) is None
