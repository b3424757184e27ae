"""The CPU oracle (oracle/ref_ops.py) against golden vectors produced by the unmodified reference.

This is what pins the oracle: if these pass, parity-vs-oracle on the GPU is parity-vs-reference."""
import pytest
import torch

from conftest import DT, golden, rel_err
from oracle import ref_ops as O

TOL = {'float32': 2e-6, 'float64': 1e-12}


def _cases(name):
    return golden(name).meta['cases']


@pytest.mark.parametrize('case', _cases('bias_act'), ids=lambda c: f"{c['key']}-{c['act']}-{c['dtype']}")
def test_bias_act(case):
    G = golden('bias_act')
    k, tol = case['key'], TOL[case['dtype']]
    x = G.t(k + '_x').requires_grad_(True)
    b = G.t(k + '_b').requires_grad_(True) if case['use_b'] else None
    y = O.bias_act(x, b, dim=case['dim'], act=case['act'], alpha=case['alpha'], gain=case['gain'], clamp=case['clamp'])
    assert y.dtype == DT[case['dtype']]
    assert rel_err(y, G.t(k + '_y')) <= tol
    grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), G.t(k + '_dy'))
    assert rel_err(grads[0], G.t(k + '_dx')) <= 10 * tol
    if b is not None:
        assert rel_err(grads[1], G.t(k + '_db')) <= 10 * tol


@pytest.mark.parametrize('case', _cases('upfirdn2d'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_upfirdn2d(case):
    G = golden('upfirdn2d')
    k, tol = case['key'], TOL[case['dtype']]
    x = G.t(k + '_x').requires_grad_(True)
    f = G.t(k + '_f') if case['has_f'] else None
    y = O.upfirdn2d(x, f, up=case['up'], down=case['down'], padding=case['padding'], flip_filter=case['flip'], gain=case['gain'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= tol
    dx, = torch.autograd.grad(y, x, G.t(k + '_dy'))
    assert rel_err(dx, G.t(k + '_dx')) <= tol


def test_upfirdn2d_helpers():
    G = golden('upfirdn2d')
    x, f = G.t('h_x'), G.t('h_f')
    assert rel_err(O.filter2d(x, f), G.t('h_filter2d')) <= 2e-6
    assert rel_err(O.upsample2d(x, f), G.t('h_upsample2d')) <= 2e-6
    assert rel_err(O.downsample2d(x, f), G.t('h_downsample2d')) <= 2e-6
    assert torch.equal(O.setup_filter([1, 3, 3, 1]), G.t('sf_a'))
    assert torch.allclose(O.setup_filter([1, 2, 1], gain=4), G.t('sf_b'), rtol=1e-6)
    assert torch.equal(O.setup_filter(list(range(1, 13))), G.t('sf_c'))
    assert torch.equal(O.setup_filter([1, 3, 3, 1], flip_filter=True, normalize=False), G.t('sf_d'))


@pytest.mark.parametrize('case', _cases('filtered_lrelu'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_filtered_lrelu(case):
    G = golden('filtered_lrelu')
    k, tol = case['key'], TOL[case['dtype']]
    x = G.t(k + '_x').requires_grad_(True)
    b = G.t(k + '_b').requires_grad_(True) if case['use_b'] else None
    fu = G.t(k + '_fu') if case['has_fu'] else None
    fd = G.t(k + '_fd') if case['has_fd'] else None
    y = O.filtered_lrelu(x, fu, fd, b, up=case['up'], down=case['down'], padding=case['padding'], gain=case['gain'],
                         slope=case['slope'], clamp=case['clamp'], flip_filter=case['flip'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 4 * tol
    grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), G.t(k + '_dy'))
    assert rel_err(grads[0], G.t(k + '_dx')) <= 4 * tol
    if b is not None:
        assert rel_err(grads[1], G.t(k + '_db')) <= 4 * tol


@pytest.mark.parametrize('case', _cases('conv2d_resample'), ids=lambda c: c['key'])
def test_conv2d_resample(case):
    G = golden('conv2d_resample')
    k = case['key']
    y = O.conv2d_resample(G.t(k + '_x'), G.t(k + '_w'), f=G.t('f'), up=case['up'], down=case['down'],
                          padding=case['padding'], flip_weight=case['flip_weight'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 1e-12


@pytest.mark.parametrize('case', _cases('conv2d_resample_ext'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_conv2d_resample_ext(case):
    """groups, down-sampling, up + down, per-axis / asymmetric / negative padding, flip_filter, separable filters."""
    G = golden('conv2d_resample_ext')
    k = case['key']
    y = O.conv2d_resample(G.t(k + '_x'), G.t(k + '_w'), f=G.t('f::' + case['filter']), up=case['up'], down=case['down'], padding=case['padding'],
                          groups=case['groups'], flip_weight=case['flip_weight'], flip_filter=case['flip_filter'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert rel_err(y, ref) <= {'float32': 2e-6, 'float64': 1e-12}[case['dtype']]


@pytest.mark.parametrize('case', _cases('modulated_conv2d'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_modulated_conv2d(case):
    G = golden('modulated_conv2d')
    k = case['key']
    tol = {'float32': 1e-5, 'float64': 1e-11}[case['dtype']]
    x = G.t(k + '_x').requires_grad_(True)
    w = G.t(k + '_weight').requires_grad_(True)
    s = G.t(k + '_styles').requires_grad_(True)
    noise = G.t(k + '_noise').requires_grad_(True) if case['noise'] else None
    f = G.t(k + '_f') if case['use_f'] else None
    y = O.modulated_conv2d(x, w, s, noise=noise, up=case['up'], padding=case['k'] // 2, resample_filter=f,
                           demodulate=case['demodulate'], flip_weight=case['flip_weight'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= tol
    leaves = [x, w, s] + ([noise] if noise is not None else [])
    grads = torch.autograd.grad(y, leaves, G.t(k + '_dy'))
    for g, name in zip(grads, ['dx', 'dweight', 'dstyles', 'dnoise']):
        assert rel_err(g, G.t(k + '_' + name)) <= tol, name
