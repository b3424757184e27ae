"""fma.fma (reference torch_utils/ops/fma.py:15-58): a * b + c with un-broadcast gradients.  Pure torch arithmetic (torch.addcmul), so
its host logic is checked on CPU; tests/test_ops_gpu.py repeats it on the device."""
import torch

from vfm_vae_b200.torch_utils.ops import fma as Fm


def test_fma_forward_and_unbroadcast_gradients_cpu():
    g = torch.Generator().manual_seed(1)
    for shapes in [((2, 3, 4, 5), (2, 3, 1, 1), (2, 1, 4, 5)), ((4, 5), (5,), (1,)), ((3, 1, 2), (1, 4, 2), (3, 4, 1)), ((2, 2), (2, 2), (2, 2)),
                   ((2, 3, 4, 4), (2, 3, 1, 1), ())]:
        a, b, c = (torch.randn(s, generator=g, dtype=torch.float64).requires_grad_(True) for s in shapes)
        y = Fm.fma(a, b, c)
        ref = a * b + c
        assert y.shape == ref.shape and torch.allclose(y, ref, rtol=1e-12, atol=1e-12)
        dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        for u, v, t in zip(torch.autograd.grad(y, [a, b, c], dy), torch.autograd.grad(ref, [a, b, c], dy), (a, b, c)):
            assert u.shape == t.shape and torch.allclose(u, v, rtol=1e-12, atol=1e-12)


def test_fma_gradcheck():
    a = torch.randn(2, 3, 1, 1, dtype=torch.float64, requires_grad=True)
    b = torch.randn(2, 3, 4, 4, dtype=torch.float64, requires_grad=True)
    c = torch.randn(1, 1, 4, 4, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(Fm.fma, (a, b, c))
