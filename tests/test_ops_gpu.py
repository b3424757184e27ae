"""Parity of the sm_100a kernels (called through the C ABI / plugin layer) against
  (1) the committed golden vectors produced by the unmodified reference, and
  (2) the CPU oracle on seeded inputs at decoder-like sizes.
Tolerances are the ones north_star states: max|a-b|/max|b| <= 1e-5 in fp32, <= 2e-3 in fp16, outputs and gradients
(fp64 is held to 1e-6 for the modulated conv, whose kernel interface takes fp32 weights/styles, and 1e-10 elsewhere)."""
import math

import pytest
import torch

from conftest import DT, golden, rel_err
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu

TOL = {'float32': 1e-5, 'float64': 1e-10, 'float16': 2e-3}
DEV = 'cuda'


def _ops():
    import vfm_vae_b200 as V
    return V


def _cases(name):
    return golden(name).meta['cases']


# ----------------------------------------------------------------------------------------------------- bias_act

@pytest.mark.parametrize('case', _cases('bias_act'), ids=lambda c: f"{c['key']}-{c['act']}-{c['dtype']}")
def test_bias_act_golden(case):
    V = _ops()
    G = golden('bias_act')
    k, tol = case['key'], TOL[case['dtype']]
    x = G.t(k + '_x', DEV).requires_grad_(True)
    b = G.t(k + '_b', DEV).requires_grad_(True) if case['use_b'] else None
    kw = dict(dim=case['dim'], act=case['act'], alpha=case['alpha'], gain=case['gain'], clamp=case['clamp'])
    y = V.bias_act.bias_act(x, b, **kw)
    assert y.dtype == DT[case['dtype']] and y.shape == x.shape
    assert rel_err(y, G.t(k + '_y')) <= tol
    dy = G.t(k + '_dy', DEV)
    grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), dy)
    assert rel_err(grads[0], G.t(k + '_dx')) <= tol
    if b is not None:
        assert rel_err(grads[1], G.t(k + '_db')) <= tol
    # second order
    dy_leaf = dy.clone().requires_grad_(True)
    y2 = V.bias_act.bias_act(x, b, **kw)
    dx2, = torch.autograd.grad(y2, x, dy_leaf, create_graph=True)
    g2 = torch.autograd.grad(dx2, [dy_leaf, x], G.t(k + '_ddx', DEV), allow_unused=True)
    assert rel_err(g2[0], G.t(k + '_g2_dy')) <= tol
    ref_g2x = G.t(k + '_g2_x')
    got_g2x = g2[1] if g2[1] is not None else torch.zeros_like(x)
    assert rel_err(got_g2x, ref_g2x) <= tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('channels_last', [False, True])
@pytest.mark.parametrize('shape', [(4, 128, 64, 64), (3, 6, 33, 17), (2, 512, 8, 8)])
def test_bias_act_decoder_shapes(dtype, channels_last, shape):
    V = _ops()
    g = torch.Generator().manual_seed(10)
    x32 = torch.randn(shape, generator=g) * 3
    b32 = torch.randn(shape[1], generator=g)
    dy32 = torch.randn(shape, generator=g)
    # the oracle sees exactly the values the kernel sees (after the cast), computed in fp32
    xq, bq, dyq = x32.to(dtype).float(), b32.to(dtype).float(), dy32.to(dtype).float()
    xr = xq.clone().requires_grad_(True)
    br = bq.clone().requires_grad_(True)
    yr = O.bias_act(xr, br, act='lrelu', gain=math.sqrt(2), clamp=2.0)
    dxr, dbr = torch.autograd.grad(yr, [xr, br], dyq)
    x = xq.to(DEV, dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    b = bq.to(DEV, dtype).requires_grad_(True)
    y = V.bias_act.bias_act(x, b, act='lrelu', gain=math.sqrt(2), clamp=2.0)
    assert y.stride() == x.stride()
    tol = TOL[str(dtype).split('.')[-1]]
    assert rel_err(y, yr) <= tol
    dx, db = torch.autograd.grad(y, [x, b], dyq.to(DEV, dtype))
    if dtype == torch.float16:
        # the gradient mask is decided from the stored (fp16-rounded) y, exactly as in the reference kernels: an element
        # whose exact y lies within one fp16 ulp of a decision boundary (0 or +-clamp) may legitimately land on the other
        # side.  Exclude that band from the max-abs comparison.
        y_unclamped = O.bias_act(xq, bq, act='lrelu', gain=math.sqrt(2), clamp=None)
        band = ((y_unclamped.abs() - 2.0).abs() < 4e-3) | (y_unclamped.abs() < 1e-3)
        # the fused bias gradient must equal the sum of the kernel's own dx (fp32 accumulation inside the kernel)
        assert rel_err(db, dx.float().sum([0, 2, 3])) <= 2e-3
        dx = torch.where(band.to(DEV), torch.zeros_like(dx), dx)
        dxr = torch.where(band, torch.zeros_like(dxr), dxr)
        assert band.float().mean() < 0.01
        assert rel_err(dx, dxr) <= tol
    else:
        assert rel_err(dx, dxr) <= tol
        assert rel_err(db, dbr) <= tol


def test_bias_act_errors():
    V = _ops()
    x = torch.randn(2, 3, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        V.bias_act.bias_act(x, torch.randn(5, device=DEV))          # wrong bias length
    with pytest.raises(RuntimeError):
        V.bias_act.bias_act(x.cpu(), None)                           # no CPU path
    with pytest.raises(RuntimeError):
        V.bias_act.bias_act(x, torch.randn(3, device=DEV).double())  # dtype mismatch
    assert V.bias_act.bias_act(x, None) is not None                  # identity short-circuit
    e = torch.empty(0, 3, 4, 4, device=DEV)
    assert V.bias_act.bias_act(e, torch.zeros(3, device=DEV), act='lrelu').shape == e.shape   # empty input


# ---------------------------------------------------------------------------------------------------- upfirdn2d

@pytest.mark.parametrize('case', _cases('upfirdn2d'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_upfirdn2d_golden(case):
    V = _ops()
    G = golden('upfirdn2d')
    k, tol = case['key'], TOL[case['dtype']]
    x = G.t(k + '_x', DEV).requires_grad_(True)
    f = G.t(k + '_f', DEV) if case['has_f'] else None
    y = V.upfirdn2d.upfirdn2d(x, f, up=case['up'], down=case['down'], padding=case['padding'], flip_filter=case['flip'], gain=case['gain'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert rel_err(y, ref) <= tol
    dx, = torch.autograd.grad(y, x, G.t(k + '_dy', DEV))
    assert rel_err(dx, G.t(k + '_dx')) <= tol


def test_upfirdn2d_helpers_golden():
    V = _ops()
    G = golden('upfirdn2d')
    x, f = G.t('h_x', DEV), G.t('h_f', DEV)
    assert rel_err(V.upfirdn2d.filter2d(x, f), G.t('h_filter2d')) <= 1e-5
    assert rel_err(V.upfirdn2d.upsample2d(x, f), G.t('h_upsample2d')) <= 1e-5
    assert rel_err(V.upfirdn2d.downsample2d(x, f), G.t('h_downsample2d')) <= 1e-5


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('cfg', [
    dict(shape=(1, 64, 65, 65), up=1, down=1, padding=[1, 1, 1, 1], gain=4.0),      # post-convT blur (tiled 1/1)
    dict(shape=(2, 8, 257, 257), up=1, down=1, padding=[1, 1, 1, 1], gain=4.0),
    dict(shape=(2, 16, 64, 64), up=2, down=1, padding=[2, 1, 2, 1], gain=4.0),       # upsample2d (tiled 2/1)
    dict(shape=(2, 16, 128, 128), up=1, down=2, padding=[1, 1, 1, 1], gain=1.0),     # downsample2d (tiled 1/2)
    dict(shape=(2, 4, 40, 52), up=2, down=1, padding=[3, 2, 1, 4], gain=1.0),        # odd phase alignment
    dict(shape=(2, 4, 40, 52), up=1, down=1, padding=[-3, 2, 5, -1], gain=1.0),      # crop + pad
])
@pytest.mark.parametrize('channels_last', [False, True])
def test_upfirdn2d_decoder_shapes(dtype, cfg, channels_last):
    V = _ops()
    g = torch.Generator().manual_seed(11)
    xq = torch.randn(cfg['shape'], generator=g).to(dtype).float()
    f = O.setup_filter([1, 3, 3, 1])
    kw = dict(up=cfg['up'], down=cfg['down'], padding=cfg['padding'], gain=cfg['gain'])
    xr = xq.clone().requires_grad_(True)
    yr = O.upfirdn2d(xr, f, **kw)
    dyq = torch.randn(yr.shape, generator=g).to(dtype).float()
    dxr, = torch.autograd.grad(yr, xr, dyq)
    x = xq.to(DEV, dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = V.upfirdn2d.upfirdn2d(x, f.to(DEV), **kw)
    tol = TOL[str(dtype).split('.')[-1]]
    assert y.shape == yr.shape
    assert rel_err(y, yr) <= tol
    dx, = torch.autograd.grad(y, x, dyq.to(DEV, dtype))
    assert rel_err(dx, dxr) <= tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('cfg', [
    # (N, C, H, W, filter, padding [x0,x1,y0,y1], flip)  -- all with 16-byte aligned rows => the streaming blur kernel
    dict(shape=(2, 5, 24, 64), f='binom4', pad=[1, 1, 1, 1], flip=False),            # the decoder's post-convT blur geometry
    dict(shape=(2, 3, 33, 264), f='binom4', pad=[2, 2, 2, 2], flip=True),            # its backward (out = in + 1)
    dict(shape=(1, 4, 16, 8), f='binom4', pad=[0, 3, 3, 0], flip=False),             # one column group, lopsided padding
    dict(shape=(3, 2, 9, 16), f='rand4', pad=[3, 0, 0, 3], flip=False),              # non-separable filter
    dict(shape=(1, 2, 40, 512), f='rand4', pad=[1, 2, 2, 1], flip=True),             # rows wider than one warp (halo across warps)
    dict(shape=(2, 2, 20, 40), f='rand3', pad=[1, 1, 1, 1], flip=False),             # 3x3 taps, 5 column groups (non power of two)
    dict(shape=(1, 3, 12, 48), f='binom4', pad=[2, -1, 1, -2], flip=False),          # negative padding on the far sides (crop)
])
def test_upfirdn2d_streaming_blur(dtype, cfg):
    V = _ops()
    g = torch.Generator().manual_seed(11)
    f = {'binom4': O.setup_filter([1, 3, 3, 1]), 'rand4': torch.randn(4, 4, generator=g), 'rand3': torch.randn(3, 3, generator=g)}[cfg['f']]
    xq = torch.randn(cfg['shape'], generator=g).to(dtype).float()
    kw = dict(padding=cfg['pad'], flip_filter=cfg['flip'], gain=4.0)
    xr = xq.clone().requires_grad_(True)
    yr = O.upfirdn2d(xr, f, **kw)
    dyq = torch.randn(yr.shape, generator=g).to(dtype).float()
    dxr, = torch.autograd.grad(yr, xr, dyq)
    x = xq.to(DEV, dtype).requires_grad_(True)
    y = V.upfirdn2d.upfirdn2d(x, f.to(DEV), **kw)
    tol = TOL[str(dtype).split('.')[-1]]
    assert y.shape == yr.shape
    assert rel_err(y, yr) <= tol
    dx, = torch.autograd.grad(y, x, dyq.to(DEV, dtype))
    assert rel_err(dx, dxr) <= tol
    # pitched rows (what the modulated conv hands to the blur): a W-slice of a wider tensor
    wide = torch.zeros(*cfg['shape'][:3], cfg['shape'][3] + 16, device=DEV, dtype=dtype)
    wide[..., :cfg['shape'][3] - 3] = xq[..., :cfg['shape'][3] - 3].to(DEV, dtype)
    wide[..., cfg['shape'][3] - 3:] = 7.0     # junk in the pitch padding must never be read as data
    xs = wide[..., :cfg['shape'][3] - 3]
    ys = V.upfirdn2d.upfirdn2d(xs, f.to(DEV), **kw)
    assert rel_err(ys, O.upfirdn2d(xq[..., :cfg['shape'][3] - 3], f, **kw)) <= tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('cfg', [dict(shape=(2, 5, 16, 32), taps=[1, 2, 1]), dict(shape=(2, 3, 24, 64), taps=[1, 4, 6, 4, 1]),
                                 dict(shape=(1, 4, 9, 8), taps=[1, 4, 6, 4, 1]), dict(shape=(1, 2, 40, 512), taps=[1, 4, 6, 4, 1]),
                                 dict(shape=(2, 2, 12, 24), taps='rand5')], ids=lambda c: f"{c['shape'][2]}x{c['shape'][3]}-{c['taps']}")
def test_blur2d_replicate(cfg, dtype):
    """replicate-pad + fixed-kernel depthwise blur of the pixel-shuffle upsampler (convnext_utils.py:250-255) in one kernel."""
    V = _ops()
    g = torch.Generator().manual_seed(13)
    if cfg['taps'] == 'rand5':
        k2 = torch.randn(5, 5, generator=g)
    else:
        k = torch.tensor(cfg['taps'], dtype=torch.float32)
        k2 = torch.outer(k, k)
        k2 = k2 / k2.sum()
    kh, kw = k2.shape
    pad = ((kw - 1) // 2, (kw - 1) // 2 + int(kw % 2 == 0), (kh - 1) // 2, (kh - 1) // 2 + int(kh % 2 == 0))
    xq = torch.randn(cfg['shape'], generator=g).to(dtype).float()
    Cc = cfg['shape'][1]
    yr = torch.nn.functional.conv2d(torch.nn.functional.pad(xq, pad, mode='replicate'), k2[None, None].repeat(Cc, 1, 1, 1), groups=Cc)
    with torch.no_grad():
        y = V.upfirdn2d.blur2d_replicate(xq.to(DEV, dtype), k2.to(DEV), pad)
    assert y is not None and y.shape == yr.shape and y.dtype == dtype
    assert rel_err(y, yr) <= TOL[str(dtype).split('.')[-1]]


@pytest.mark.parametrize('cfg', [dict(shape=(2, 6, 16, 32), k=5), dict(shape=(2, 5, 24, 64), k=7), dict(shape=(1, 3, 9, 8), k=7),
                                 dict(shape=(1, 2, 70, 512), k=7), dict(shape=(3, 4, 12, 24), k=5)], ids=lambda c: f"k{c['k']}-{c['shape'][2]}x{c['shape'][3]}")
@pytest.mark.parametrize('bias', [True, False])
@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
def test_depthwise_conv2d(cfg, bias, dtype):
    """k x k depthwise conv + bias of the ConvNeXt layers (convnext_utils.py:99,128) against F.conv2d in fp64."""
    V = _ops()
    g = torch.Generator().manual_seed(14)
    Cc, k = cfg['shape'][1], cfg['k']
    xq = torch.randn(cfg['shape'], generator=g).to(dtype).double()
    w = (torch.randn(Cc, 1, k, k, generator=g) * 0.2).double()
    b = (torch.randn(Cc, generator=g) * 0.3).double() if bias else None
    noise = torch.randn(1, 1, *cfg['shape'][2:], generator=g) if bias else None       # the legacy layers' broadcast noise addend
    yr = torch.nn.functional.conv2d(xq, w, b, padding=k // 2, groups=Cc)
    if noise is not None:
        yr = yr + noise.double()
    with torch.no_grad():
        y = V.upfirdn2d.depthwise_conv2d(xq.to(DEV, dtype), w.float().to(DEV), b.float().to(DEV) if bias else None,
                                         noise.to(DEV) if noise is not None else None)
    assert y is not None and y.shape == yr.shape and y.dtype == dtype
    assert rel_err(y, yr) <= (2e-3 if dtype == torch.float16 else 1e-5)


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
@pytest.mark.parametrize('shape', [(2, 8, 6, 8), (1, 4, 5, 4), (3, 12, 16, 64), (2, 16, 33, 128)], ids=lambda s: 'x'.join(map(str, s)))
def test_pixel_shuffle2(shape, dtype):
    """PixelShuffle(2) of SeparableUpsampleWithFixedBlur (convnext_utils.py:197-257): a pure permutation, bit-exact."""
    V = _ops()
    g = torch.Generator().manual_seed(15)
    x = torch.randn(shape, generator=g).to(dtype).to(DEV)
    with torch.no_grad():
        y = V.upfirdn2d.pixel_shuffle2(x)
    assert y is not None and torch.equal(y, torch.nn.functional.pixel_shuffle(x, 2))
    assert V.upfirdn2d.pixel_shuffle2(x[:, :, :, :3].contiguous()) is None        # W % 4 != 0: the caller keeps the stock op


def test_depthwise_conv2d_k3():
    """3x3 depthwise conv (no bias) of SeparableUpsampleWithFixedBlur against F.conv2d."""
    V = _ops()
    g = torch.Generator().manual_seed(16)
    for dtype, tol in ((torch.float16, 2e-3), (torch.float32, 1e-5)):
        xq = torch.randn(2, 6, 20, 40, generator=g).to(dtype).double()
        w = (torch.randn(6, 1, 3, 3, generator=g) * 0.3).double()
        yr = torch.nn.functional.conv2d(xq, w, None, padding=1, groups=6)
        with torch.no_grad():
            y = V.upfirdn2d.depthwise_conv2d(xq.to(DEV, dtype), w.float().to(DEV))
        assert y is not None and rel_err(y, yr) <= tol


@pytest.mark.parametrize('cfg', [dict(shape=(2, 6, 16, 32), k=5), dict(shape=(2, 5, 24, 64), k=7), dict(shape=(3, 4, 12, 24), k=3),
                                 dict(shape=(1, 3, 70, 256), k=7), dict(shape=(2, 4, 8, 8), k=5)], ids=lambda c: f"k{c['k']}-{c['shape'][2]}x{c['shape'][3]}")
@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
def test_depthwise_conv2d_autograd(cfg, dtype):
    """Gradients of the depthwise conv (data gradient = the same kernel with flipped taps, weight / bias gradient by
    vfm_depthwise_wgrad) against stock autograd of F.conv2d in fp64."""
    V = _ops()
    g = torch.Generator().manual_seed(17)
    Cc, k = cfg['shape'][1], cfg['k']
    x0 = torch.randn(cfg['shape'], generator=g).to(dtype)
    w0 = torch.randn(Cc, 1, k, k, generator=g) * 0.2
    b0 = torch.randn(Cc, generator=g) * 0.3
    dy0 = torch.randn(cfg['shape'], generator=g).to(dtype)
    xr, wr, br = x0.double().requires_grad_(True), w0.double().requires_grad_(True), b0.double().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, br, padding=k // 2, groups=Cc)
    x, w, b = x0.to(DEV).requires_grad_(True), w0.to(DEV).requires_grad_(True), b0.to(DEV).requires_grad_(True)
    tol = 2e-3 if dtype == torch.float16 else 1e-5
    nz0 = torch.randn(1, 1, *cfg['shape'][2:], generator=g)
    nzr = nz0.double().requires_grad_(True)
    yr = yr + nzr
    gr = torch.autograd.grad(yr, [xr, wr, br, nzr], dy0.double())
    nz = nz0.to(DEV).requires_grad_(True)
    y = V.upfirdn2d.depthwise_conv2d(x, w, b, nz)
    assert y is not None and y.requires_grad and y.dtype == dtype
    gx, gw, gb, gn = torch.autograd.grad(y, [x, w, b, nz], dy0.to(DEV))
    assert gn.shape == nz.shape and rel_err(gn, gr[3]) <= tol
    assert rel_err(y, yr) <= tol
    assert gx.dtype == dtype and gw.dtype == torch.float32 and gw.shape == w.shape
    assert rel_err(gx, gr[0]) <= tol
    assert rel_err(gw, gr[1]) <= tol
    assert rel_err(gb, gr[2]) <= tol


def test_pixel_shuffle2_autograd():
    V = _ops()
    g = torch.Generator().manual_seed(18)
    for dtype in (torch.float16, torch.float32):
        x = torch.randn(2, 8, 6, 16, generator=g).to(dtype).to(DEV).requires_grad_(True)
        dy = torch.randn(2, 2, 12, 32, generator=g).to(dtype).to(DEV)
        y = V.upfirdn2d.pixel_shuffle2(x)
        assert torch.equal(y, torch.nn.functional.pixel_shuffle(x, 2))
        (gx,) = torch.autograd.grad(y, [x], dy)
        assert torch.equal(gx, torch.nn.functional.pixel_unshuffle(dy, 2))


@pytest.mark.parametrize('k', [3, 5])
@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
@pytest.mark.parametrize('shape', [(2, 3, 7, 10), (1, 2, 5, 5), (3, 1, 9, 16), (1, 4, 12, 6)], ids=lambda s: 'x'.join(map(str, s)))
def test_replicate_blur_edges_entry_point(k, dtype, shape):
    """``vfm_replicate_blur_edges`` on its own (the autograd wrapper only reaches it with widths that are a multiple of 8): border elements against
    the oracle's model -- widths that are not a multiple of 4 take the scalar column loads, the others the vector ones -- and every interior
    element of dx left untouched."""
    from vfm_vae_b200.plugins import upfirdn2d_plugin as P
    g = torch.Generator().manual_seed(23)
    f = torch.randn(k, k, generator=g)
    dyq = torch.randn(shape, generator=g).to(dtype)
    marker = torch.full(shape, 7.0, dtype=dtype)
    dx = P.replicate_blur_edges(dyq.to(DEV), marker.to(DEV).clone(), f.to(DEV))
    want = O.replicate_blur_edges(dyq.double(), f.double(), torch.full(shape, 7.0, dtype=torch.float64))
    tol = 2e-3 if dtype == torch.float16 else 1e-5
    assert rel_err(dx, want) <= tol
    assert torch.equal(dx[:, :, 1:-1, 1:-1].cpu(), marker[:, :, 1:-1, 1:-1])


@pytest.mark.parametrize('taps', ['random', 'binomial'])
@pytest.mark.parametrize('k', [3, 5])
@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
def test_blur2d_replicate_autograd(k, dtype, taps):
    """Backward of replicate-pad + fixed blur (SeparableUpsampleWithFixedBlur, convnext_utils.py:250-255): one zero-padded stencil pass
    over dy + the border row / column rewritten by vfm_replicate_blur_edges, against stock autograd in fp64.  Random (non-symmetric, full-rank)
    taps take the dense stencil body, the decoder's binomial taps (rank 1) the separable one."""
    V = _ops()
    g = torch.Generator().manual_seed(19)
    p = k // 2
    for shape in ((2, 3, 16, 24), (1, 2, 40, 64), (2, 2, 8, 8)):
        if taps == 'random':
            f = torch.randn(k, k, generator=g) * 0.3
        else:
            t = torch.tensor({3: [1., 2., 1.], 5: [1., 4., 6., 4., 1.]}[k])
            f = torch.outer(t, t) / t.sum() ** 2
        x0 = torch.randn(shape, generator=g).to(dtype)
        dy0 = torch.randn(shape, generator=g).to(dtype)
        xr = x0.double().requires_grad_(True)
        yr = torch.nn.functional.conv2d(torch.nn.functional.pad(xr, (p, p, p, p), mode='replicate'), f.double()[None, None].repeat(shape[1], 1, 1, 1), groups=shape[1])
        (gr,) = torch.autograd.grad(yr, [xr], dy0.double())
        x = x0.to(DEV).requires_grad_(True)
        y = V.upfirdn2d.blur2d_replicate(x, f.to(DEV), (p, p, p, p))
        assert y is not None and y.requires_grad
        (gx,) = torch.autograd.grad(y, [x], dy0.to(DEV))
        tol = 2e-3 if dtype == torch.float16 else 1e-5
        assert rel_err(y, yr) <= tol and rel_err(gx, gr) <= tol


def test_upfirdn2d_errors():
    V = _ops()
    x = torch.randn(1, 1, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        V.upfirdn2d.upfirdn2d(x, torch.ones(8, 8, device=DEV))                    # output smaller than 1x1
    with pytest.raises(RuntimeError):
        V.upfirdn2d.upfirdn2d(x, torch.ones(2, 2, device=DEV, dtype=torch.float64))   # f must be fp32


# ----------------------------------------------------------------------------------------------- filtered_lrelu

@pytest.mark.parametrize('case', _cases('filtered_lrelu'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_filtered_lrelu_golden(case):
    V = _ops()
    G = golden('filtered_lrelu')
    k = case['key']
    tol = {'float32': 2e-5, 'float64': 1e-10}[case['dtype']]
    x = G.t(k + '_x', DEV).requires_grad_(True)
    b = G.t(k + '_b', DEV).requires_grad_(True) if case['use_b'] else None
    fu = G.t(k + '_fu', DEV) if case['has_fu'] else None
    fd = G.t(k + '_fd', DEV) if case['has_fd'] else None
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y = V.filtered_lrelu.filtered_lrelu(x, fu, fd, b, up=case['up'], down=case['down'], padding=case['padding'],
                                            gain=case['gain'], slope=case['slope'], clamp=case['clamp'], flip_filter=case['flip'])
        ref = G.t(k + '_y')
        assert y.shape == ref.shape
        assert rel_err(y, ref) <= tol
        grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), G.t(k + '_dy', DEV))
    assert rel_err(grads[0], G.t(k + '_dx')) <= tol
    if b is not None:
        assert rel_err(grads[1], G.t(k + '_db')) <= tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('shape', [(2, 16, 64, 64), (1, 8, 37, 50)])
def test_filtered_lrelu_stylegan3_shape(dtype, shape):
    """separable 12-tap up2/down2, padding chosen so out == in."""
    V = _ops()
    g = torch.Generator().manual_seed(12)
    xq = torch.randn(shape, generator=g).to(dtype).float()
    bq = (torch.randn(shape[1], generator=g) * 0.3).to(dtype).float()
    fu = O.setup_filter([1, 4, 8, 12, 14, 16, 16, 14, 12, 8, 4, 1])
    fd = O.setup_filter([1, 3, 6, 10, 14, 16, 16, 14, 10, 6, 3, 1])
    kw = dict(up=2, down=2, padding=[10, 11, 10, 11], gain=math.sqrt(2), slope=0.2, clamp=0.8)
    xr, br = xq.clone().requires_grad_(True), bq.clone().requires_grad_(True)
    yr = O.filtered_lrelu(xr, fu, fd, br, **kw)
    assert yr.shape == xq.shape
    dyq = torch.randn(yr.shape, generator=g).to(dtype).float()
    dxr, dbr = torch.autograd.grad(yr, [xr, br], dyq)
    x = xq.to(DEV, dtype).requires_grad_(True)
    b = bq.to(DEV, dtype).requires_grad_(True)
    y = V.filtered_lrelu.filtered_lrelu(x, fu.to(DEV), fd.to(DEV), b, **kw)
    tol = {torch.float32: 2e-5, torch.float16: 2e-3}[dtype]
    assert rel_err(y, yr) <= tol
    dx, db = torch.autograd.grad(y, [x, b], dyq.to(DEV, dtype))
    assert rel_err(dx, dxr) <= tol
    assert rel_err(db, dbr) <= (tol if dtype == torch.float32 else 6e-3)


def test_filtered_lrelu_sign_tensor_format():
    """The 2-bit sign tensor written by the fused kernel follows the reference layout and matches the oracle's codes."""
    from vfm_vae_b200.plugins import filtered_lrelu_plugin as P
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 3, 10, 9, generator=g)
    b = torch.randn(3, generator=g) * 0.2
    fu = O.setup_filter(list(range(1, 9)))
    fd = O.setup_filter([1, 3, 3, 1])
    up, down, pad, gain, slope, clamp = 2, 2, [5, 5, 4, 6], 1.1, 0.2, 0.6
    y, so, rc = P.filtered_lrelu(x.to(DEV), fu.to(DEV), fd.to(DEV), b.to(DEV), torch.empty([0]), up, down, *pad, 0, 0, gain, slope, clamp, False, True)
    assert rc == 0 and so.dtype == torch.uint8
    codes = O.filtered_lrelu_signs(x, fu, b, up=up, padding=pad, gain=gain, slope=slope, clamp=clamp)
    yw, yh = y.shape[3], y.shape[2]
    sw_active = yw * down - (down - 1) + fd.shape[-1] - 1
    sh = yh * down - (down - 1) + fd.shape[0] - 1
    assert so.shape == (2, 3, sh, ((sw_active + 15) & ~15) >> 2)
    so = so.cpu()
    unpacked = torch.stack([(so >> (2 * k)) & 3 for k in range(4)], dim=-1).reshape(2, 3, sh, -1)
    assert torch.equal(unpacked[..., :sw_active], codes[:, :, :sh, :sw_active])


def test_filtered_lrelu_composed_fallback():
    """More than 32 taps -> plugin answers return_code -1 -> wrapper composes upfirdn2d + act + upfirdn2d."""
    V = _ops()
    g = torch.Generator().manual_seed(14)
    x = torch.randn(1, 2, 40, 40, generator=g)
    fu = O.setup_filter(list(range(1, 41)), separable=True)
    xr = x.clone().requires_grad_(True)
    yr = O.filtered_lrelu(xr, fu, None, None, up=2, padding=[20, 19, 20, 19], clamp=0.5)
    dy = torch.randn(yr.shape, generator=g)
    dxr, = torch.autograd.grad(yr, xr, dy)
    xc = x.to(DEV).requires_grad_(True)
    with pytest.warns(RuntimeWarning):
        y = V.filtered_lrelu.filtered_lrelu(xc, fu.to(DEV), None, None, up=2, padding=[20, 19, 20, 19], clamp=0.5)
    assert rel_err(y, yr) <= 2e-5
    with pytest.warns(RuntimeWarning):
        dx, = torch.autograd.grad(y, xc, dy.to(DEV))
    assert rel_err(dx, dxr) <= 2e-5


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [dict(up=2, down=2, taps=12, pad=[10, 11, 10, 11]), dict(up=1, down=2, taps=8, pad=[3, 4, 3, 4]),
                                 dict(up=2, down=1, taps=5, pad=2, two_d=True)], ids=lambda c: f"up{c['up']}down{c['down']}t{c['taps']}")
def test_filtered_lrelu_fused_bias_gradient(cfg, dtype):
    """db comes out of the backward kernel itself (y_sum accumulation) -- it must equal dx.sum([0, 2, 3]) (reference filtered_lrelu.py:266);
    under create_graph the wrapper switches to the differentiable reduction and second-order gradients flow."""
    V = _ops()
    g = torch.Generator().manual_seed(60)
    x = torch.randn(3, 6, 24, 20, generator=g).to(dtype).to(DEV).requires_grad_(True)
    b = (torch.randn(6, generator=g) * 0.5).to(dtype).to(DEV).requires_grad_(True)
    f1 = O.setup_filter(list(range(1, cfg['taps'] + 1)))
    f = (torch.outer(f1, f1) if cfg.get('two_d') and f1.ndim == 1 else f1).to(DEV)
    y = V.filtered_lrelu.filtered_lrelu(x, f, f, b, up=cfg['up'], down=cfg['down'], padding=cfg['pad'], clamp=0.8)
    dy = torch.randn(y.shape, generator=g).to(dtype).to(DEV)
    dx, db = torch.autograd.grad(y, [x, b], dy, retain_graph=True)
    want = dx.float().sum([0, 2, 3])
    assert db.dtype == b.dtype and rel_err(db, want) <= (1e-5 if dtype == torch.float32 else 1e-3)
    if dtype == torch.float32:
        # higher-order use (create_graph): db is then the differentiable reduction of dx and gradients flow back to the incoming dy
        dy2 = dy.clone().requires_grad_(True)
        dx2, db2 = torch.autograd.grad(y, [x, b], dy2, create_graph=True)
        assert rel_err(db2, want) <= 1e-5 and db2.requires_grad
        (g_dy,) = torch.autograd.grad(db2.sum(), [dy2])
        assert g_dy.shape == dy.shape and torch.isfinite(g_dy).all() and g_dy.abs().max() > 0


# --------------------------------------------------------------------------------------------- modulated_conv2d

@pytest.mark.parametrize('case', _cases('modulated_conv2d'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_modulated_conv2d_golden(case):
    V = _ops()
    G = golden('modulated_conv2d')
    k = case['key']
    tol = {'float32': 1e-5, 'float64': 1e-6}[case['dtype']]
    x = G.t(k + '_x', DEV).requires_grad_(True)
    w = G.t(k + '_weight', DEV).requires_grad_(True)
    s = G.t(k + '_styles', DEV).requires_grad_(True)
    noise = G.t(k + '_noise', DEV).requires_grad_(True) if case['noise'] else None
    f = G.t(k + '_f', DEV) if case['use_f'] else None
    y = V.modulated_conv2d(x, w, s, noise=noise, up=case['up'], padding=case['k'] // 2, resample_filter=f,
                           demodulate=case['demodulate'], flip_weight=case['flip_weight'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert rel_err(y, ref) <= tol
    leaves = [x, w, s] + ([noise] if noise is not None else [])
    grads = torch.autograd.grad(y, leaves, G.t(k + '_dy', DEV))
    for gr, name in zip(grads, ['dx', 'dweight', 'dstyles', 'dnoise']):
        assert rel_err(gr, G.t(k + '_' + name)) <= tol, name


def _modconv_vs_oracle(N, I, O_, H, W, k, up, demod, dtype, noise_kind, generic, seed=20, oracle_dtype=torch.float32):
    """``oracle_dtype=torch.float64`` evaluates the CPU oracle in double on the same (dtype-rounded) inputs: used at the benchmarked widths, where
    K = I*9 up to 6912 accumulations would eat most of a 1e-5 gate in the fp32 oracle's own rounding."""
    V = _ops()
    from vfm_vae_b200.torch_utils.ops import modulated_conv2d as M
    g = torch.Generator().manual_seed(seed)
    xq = torch.randn(N, I, H, W, generator=g).to(dtype).float()
    w = torch.randn(O_, I, k, k, generator=g)
    s = torch.randn(N, I, generator=g) + 1
    f = O.setup_filter([1, 3, 3, 1]) if up == 2 else None
    noise = None
    if noise_kind == 'const':
        noise = torch.randn(H * up, W * up, generator=g) * 0.3
    elif noise_kind == 'random':
        noise = torch.randn(N, 1, H * up, W * up, generator=g) * 0.3
    leaves_r = [t.clone().to(oracle_dtype).requires_grad_(True) for t in ([xq, w, s] + ([noise] if noise is not None else []))]
    yr = O.modulated_conv2d(leaves_r[0], leaves_r[1], leaves_r[2], noise=(leaves_r[3] if noise is not None else None), up=up,
                            padding=k // 2, resample_filter=(f.to(oracle_dtype) if f is not None else None), demodulate=demod, flip_weight=(up == 1))
    dyq = torch.randn(yr.shape, generator=g).to(dtype).float()
    gr = torch.autograd.grad(yr, leaves_r, dyq.to(oracle_dtype))
    leaves = [xq.to(DEV, dtype).requires_grad_(True), w.to(DEV).requires_grad_(True), s.to(DEV).requires_grad_(True)]
    if noise is not None:
        leaves.append(noise.to(DEV).requires_grad_(True))
    M.force_generic = generic
    try:
        y = V.modulated_conv2d(leaves[0], leaves[1], leaves[2], noise=(leaves[3] if noise is not None else None), up=up,
                               padding=k // 2, resample_filter=(f.to(DEV) if f is not None else None), demodulate=demod,
                               flip_weight=(up == 1))
        gg = torch.autograd.grad(y, leaves, dyq.to(DEV, dtype))
    finally:
        M.force_generic = False
    tol = TOL[str(dtype).split('.')[-1]]
    assert y.dtype == dtype and y.shape == yr.shape
    assert rel_err(y, yr) <= tol, 'y'
    for a, b, name in zip(gg, gr, ['dx', 'dweight', 'dstyles', 'dnoise']):
        assert rel_err(a, b) <= tol, name


@pytest.mark.parametrize('generic', [True, False], ids=['generic', 'auto'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [
    dict(N=2, I=64, O_=64, H=16, W=16, k=3, up=1, demod=True, noise_kind='const'),
    dict(N=2, I=128, O_=64, H=8, W=8, k=3, up=2, demod=True, noise_kind='const'),
    dict(N=3, I=64, O_=128, H=24, W=24, k=3, up=1, demod=True, noise_kind='random'),
    dict(N=2, I=64, O_=3, H=32, W=32, k=1, up=1, demod=False, noise_kind=None),          # ToRGB
    dict(N=1, I=192, O_=64, H=32, W=32, k=3, up=1, demod=True, noise_kind=None),
    dict(N=2, I=40, O_=24, H=9, W=13, k=3, up=1, demod=True, noise_kind='const'),         # ragged channels / sizes
    dict(N=2, I=40, O_=24, H=9, W=13, k=3, up=2, demod=True, noise_kind='const'),
    # shapes the tcgen05 implicit-GEMM path accepts (channels % 128 == 0, power-of-two images)
    dict(N=2, I=128, O_=128, H=16, W=16, k=3, up=1, demod=True, noise_kind='const'),
    dict(N=2, I=256, O_=128, H=32, W=32, k=3, up=1, demod=True, noise_kind='random'),
    dict(N=4, I=128, O_=256, H=8, W=8, k=3, up=1, demod=True, noise_kind='const'),
    dict(N=1, I=128, O_=128, H=64, W=64, k=3, up=1, demod=True, noise_kind=None),
    dict(N=2, I=128, O_=128, H=16, W=16, k=1, up=1, demod=False, noise_kind=None),
    dict(N=2, I=128, O_=128, H=16, W=16, k=3, up=2, demod=True, noise_kind='const'),     # 4-phase transposed conv + strided dgrad
    dict(N=3, I=256, O_=128, H=8, W=8, k=3, up=2, demod=True, noise_kind='random'),
    dict(N=2, I=128, O_=128, H=4, W=4, k=3, up=2, demod=True, noise_kind='const'),       # block-0 sized: multi-sample tiles
    dict(N=1, I=128, O_=256, H=33, W=20, k=3, up=1, demod=True, noise_kind='const'),      # ragged image, partial tiles
    # fp16 rows of a multiple of 64 pixels: the forward reads x straight from NCHW (MN-major TMA boxes, per-sample weights)
    dict(N=2, I=256, O_=128, H=12, W=64, k=3, up=1, demod=True, noise_kind='random'),
    dict(N=1, I=128, O_=128, H=6, W=128, k=3, up=1, demod=True, noise_kind='const'),      # partial tile rows
    dict(N=2, I=128, O_=256, H=8, W=64, k=1, up=1, demod=False, noise_kind=None),
    # >= 2048 output pixels with noise: the backward's one-pass gsum + dnoise kernel, both noise layouts
    dict(N=2, I=128, O_=128, H=64, W=64, k=3, up=1, demod=True, noise_kind='const'),
    dict(N=2, I=128, O_=128, H=32, W=64, k=3, up=1, demod=True, noise_kind='random'),
    dict(N=1, I=128, O_=128, H=32, W=32, k=3, up=2, demod=True, noise_kind='const'),      # 64x64 output through the pitched blur backward
], ids=lambda c: f"N{c['N']}I{c['I']}O{c['O_']}H{c['H']}k{c['k']}up{c['up']}")
def test_modulated_conv2d_vs_oracle(cfg, dtype, generic):
    _modconv_vs_oracle(dtype=dtype, generic=generic, **cfg)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [dict(N=2, I=512, H=8, W=8), dict(N=3, I=512, H=16, W=16), dict(N=2, I=256, H=32, W=32), dict(N=2, I=200, H=8, W=8),
                                 dict(N=1, I=128, H=12, W=8), dict(N=2, I=128, H=64, W=64), dict(N=70, I=64, H=32, W=32)],
                         ids=lambda c: f"N{c['N']}I{c['I']}H{c['H']}W{c['W']}")
def test_torgb_pointwise_shapes(cfg, dtype):
    """ToRGB (1x1, 3 outputs, no demodulation, networks/generator.py:284-312) at the small-image / many-channel shapes of the decoder's first
    blocks: the forward splits the channels over slices of a CTA there (2 .. 16 slices, partial last pixel group, channel counts that are
    not a multiple of the slice count), the backward correlation runs one warp per channel."""
    _modconv_vs_oracle(O_=3, k=1, up=1, demod=False, dtype=dtype, noise_kind=None, generic=False, **cfg)


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['fp16', 'fp32'])
@pytest.mark.parametrize('up', [1, 2])
def test_kept_forward_operand_gives_the_same_weight_gradient(dtype, up):
    """Training keeps the forward's NHWC operand for the weight gradient (keep_forward_operand): same gradients as re-laying x in the backward
    (up to the fp32 atomics of the split-K weight gradient), with styles > 1 so that the power-of-two operand scale is not 1."""
    V = _ops()
    from vfm_vae_b200.torch_utils.ops import modulated_conv2d as M
    g = torch.Generator().manual_seed(40)
    N, I, O_, H = 3, 256, 128, 16
    x = torch.randn(N, I, H, H, generator=g).to(dtype).to(DEV)
    w = torch.randn(O_, I, 3, 3, generator=g).to(DEV)
    st = (torch.randn(N, I, generator=g) * 3 + 1).to(DEV)
    f = O.setup_filter([1, 3, 3, 1]).to(DEV) if up == 2 else None
    dy = None
    res = {}
    for keep in (True, False):
        M.keep_forward_operand = keep
        try:
            leaves = [x.clone().requires_grad_(True), w.clone().requires_grad_(True), st.clone().requires_grad_(True)]
            y = V.modulated_conv2d(leaves[0], leaves[1], leaves[2], up=up, padding=1, resample_filter=f, demodulate=(dtype == torch.float16), flip_weight=(up == 1))
            if dy is None:
                dy = torch.randn(y.shape, generator=g).to(dtype).to(DEV)
            res[keep] = (y, torch.autograd.grad(y, leaves, dy))
        finally:
            M.keep_forward_operand = True
    assert torch.equal(res[True][0], res[False][0])
    for a, b, name in zip(res[True][1], res[False][1], ['dx', 'dw', 'ds']):
        assert rel_err(a, b) <= 2e-6, name


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['fp16', 'fp32'])
@pytest.mark.parametrize('cfg', [
    dict(N=2, I=128, O_=128, H=64, up=1, noise='const', clamp=256.0, gain=1.0),
    dict(N=2, I=256, O_=128, H=48, up=1, noise='random', clamp=3.0, gain=math.sqrt(2)),       # a clamp that bites: the gradient mask
    dict(N=2, I=256, O_=128, H=32, up=2, noise='const', clamp=2.0, gain=math.sqrt(2)),        # epilogue in the blur, backward through it
    dict(N=1, I=128, O_=256, H=64, up=1, noise=None, clamp=None, gain=1.0),
], ids=lambda c: f"I{c['I']}O{c['O_']}H{c['H']}up{c['up']}{c['noise']}")
def test_fused_training_layer_vs_oracle(cfg, dtype):
    """modulated conv + noise + bias_act as ONE autograd node (training): forward and every gradient (x, weight, styles, bias, noise) against the
    oracle's composition of the two reference ops, and against this package's own unfused composition."""
    V = _ops()
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_synthesis_layer_train
    g = torch.Generator().manual_seed(50)
    N, I, O_, H, up = cfg['N'], cfg['I'], cfg['O_'], cfg['H'], cfg['up']
    xq = torch.randn(N, I, H, H, generator=g).to(dtype).float()
    w = torch.randn(O_, I, 3, 3, generator=g)
    st = torch.randn(N, I, generator=g) + 1
    b = (torch.randn(O_, generator=g) * 0.3).to(dtype).float()
    f = O.setup_filter([1, 3, 3, 1]) if up == 2 else None
    noise = None
    if cfg['noise'] == 'const':
        noise = torch.randn(H * up, H * up, generator=g) * 0.3
    elif cfg['noise'] == 'random':
        noise = torch.randn(N, 1, H * up, H * up, generator=g) * 0.3
    lv = [xq, w, st, b] + ([noise] if noise is not None else [])
    ref = [t.clone().double().requires_grad_(True) for t in lv]
    yr_pre = O.modulated_conv2d(ref[0], ref[1], ref[2], noise=(ref[4] if noise is not None else None), up=up, padding=1,
                                resample_filter=(f.double() if f is not None else None), flip_weight=(up == 1))
    yr = O.bias_act(yr_pre, ref[3], act='lrelu', gain=cfg['gain'], clamp=cfg['clamp'])
    dyq = torch.randn(yr.shape, generator=g).to(dtype).float()
    dev = [xq.to(DEV, dtype).requires_grad_(True)] + [t.to(DEV).requires_grad_(True) for t in lv[1:]]
    y = fused_synthesis_layer_train(dev[0], dev[1], dev[2], dev[3], noise=(dev[4] if noise is not None else None), up=up, padding=1,
                                    resample_filter=(f.to(DEV) if f is not None else None), flip_weight=(up == 1), act='lrelu', gain=cfg['gain'], clamp=cfg['clamp'])
    assert y is not None, 'expected the fused training kernel for this shape'
    gg = torch.autograd.grad(y, dev, dyq.to(DEV, dtype))
    tol = TOL[str(dtype).split('.')[-1]]
    names = ['dx', 'dweight', 'dstyles', 'dbias', 'dnoise']
    assert rel_err(y, yr) <= tol, 'y'
    # Reference gradients.  The lrelu branch and the clamp mask of an element are decided from the STORED output (as in the reference's
    # bias_act kernels, bias_act.cu:72,141): an element whose value rounds across 0 or +-clamp takes the other branch than an fp64 evaluation,
    # and one such element moves max|dx| by ~3 %.  The branch decisions are therefore taken from the kernel's own y (whose values the
    # assertion above holds to the oracle), everything else from the oracle: dz = dy * gain * slope(y) * [|y| < clamp], then the oracle's
    # modulated-conv backward on dz, db = sum dz.
    yk = y.detach().double().cpu()
    dz = dyq.double() * cfg['gain'] * torch.where(yk > 0, 1.0, 0.2)
    if cfg['clamp'] is not None:
        dz = dz * (yk.abs() < cfg['clamp'])
    dz = dz.to(dtype).double()                      # the kernel rounds dz to the activation dtype before the conv gradients consume it
    g_conv = torch.autograd.grad(yr_pre, [ref[0], ref[1], ref[2]] + ([ref[4]] if noise is not None else []), dz)
    gr = list(g_conv[:3]) + [dz.sum(dim=(0, 2, 3))] + list(g_conv[3:])
    errs = {name: rel_err(a, b_) for a, b_, name in zip(gg, gr, names)}
    assert max(errs.values()) <= tol, errs
    # and the unfused composition of this package (same kernels underneath): bias_act on the modulated conv.  fp32 only: in fp16 the two
    # forwards round y differently, so a few elements sit on different lrelu branches (see above)
    if dtype == torch.float32:
        dev2 = [t.detach().clone().requires_grad_(True) for t in dev]
        y2 = V.modulated_conv2d(dev2[0], dev2[1], dev2[2], noise=(dev2[4] if noise is not None else None), up=up, padding=1,
                                resample_filter=(f.to(DEV) if f is not None else None), flip_weight=(up == 1))
        y2 = V.bias_act.bias_act(y2, dev2[3], act='lrelu', gain=cfg['gain'], clamp=cfg['clamp'])
        g2 = torch.autograd.grad(y2, dev2, dyq.to(DEV, dtype))
        assert rel_err(y, y2) <= 2e-6
        for a_, b_, name in zip(gg, g2, names):
            assert rel_err(a_, b_) <= 2e-5, name


def test_tensor_core_path_is_taken_for_decoder_shapes():
    """The hot decoder layers must be routed to the tcgen05 kernel (and the odd shapes must not)."""
    from vfm_vae_b200.plugins import modconv_plugin as P
    from vfm_vae_b200 import _lib
    x = torch.empty(2, 128, 16, 16, device=DEV, dtype=torch.float16)
    w = torch.empty(128, 128, 3, 3, device=DEV)
    assert P.uses_tensor_cores(x, w, up=1, padding=1)
    assert not P.uses_tensor_cores(torch.empty(2, 40, 9, 13, device=DEV, dtype=torch.float16), torch.empty(24, 40, 3, 3, device=DEV), up=1, padding=1)
    # and the launch really is the tcgen05 kernel: the timing registry names it
    import ctypes as C
    lib = _lib.load()
    lib.vfm_timing_enable.argtypes = [C.c_int]
    lib.vfm_timing_enable.restype = None
    lib.vfm_timing_enable(1)
    s = torch.ones(2, 128, device=DEV)
    P.forward(torch.randn(2, 128, 16, 16, device=DEV, dtype=torch.float16), torch.randn(128, 128, 3, 3, device=DEV), s, None, 1, 1, None, True, True)
    torch.cuda.synchronize()
    lib.vfm_timing_enable(0)

    buf = (_lib.KernelStat * 32)()          # the binding's own struct / argtypes (vfm_vae_b200/_lib.py)
    n = lib.vfm_timing_report(buf, 32)
    names = {buf[i].name.decode() for i in range(min(n, 32))}
    assert 'modconv_tc_fwd' in names and 'modconv_generic_conv' not in names, names


def test_modulated_conv2d_errors():
    V = _ops()
    x = torch.randn(2, 4, 8, 8, device=DEV)
    w = torch.randn(3, 4, 3, 3, device=DEV)
    s = torch.ones(2, 4, device=DEV)
    with pytest.raises(NotImplementedError):
        V.modulated_conv2d(x, w, s, down=2, padding=1)
    with pytest.raises(AssertionError):
        V.modulated_conv2d(x, w, torch.ones(2, 5, device=DEV), padding=1)
    with pytest.raises(RuntimeError):
        V.modulated_conv2d(x.cpu(), w.cpu(), s.cpu(), padding=1)
    with pytest.raises(RuntimeError):
        V.modulated_conv2d(x, w, s, up=2, padding=1)     # up=2 without a resample filter


@pytest.mark.parametrize('case', _cases('conv2d_resample'), ids=lambda c: c['key'])
def test_conv2d_resample_golden(case):
    V = _ops()
    G = golden('conv2d_resample')
    k = case['key']
    y = V.conv2d_resample.conv2d_resample(G.t(k + '_x', DEV), G.t(k + '_w', DEV), f=G.t('f', DEV), up=case['up'], down=case['down'],
                                          padding=case['padding'], flip_weight=case['flip_weight'])
    ref = G.t(k + '_y')
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 1e-6


@pytest.mark.parametrize('case', _cases('conv2d_resample_ext'), ids=lambda c: f"{c['key']}-{c['dtype']}")
def test_conv2d_resample_ext_golden(case):
    """Everything beyond the decoder's subset (reference conv2d_resample.py:95-141): groups, down > 1, up and down together, per-axis /
    asymmetric / negative padding, flip_filter with an asymmetric filter, separable 1-D filters -- forward and input / weight gradients
    against the oracle."""
    V = _ops()
    G = golden('conv2d_resample_ext')
    k = case['key']
    kw = dict(up=case['up'], down=case['down'], padding=case['padding'], groups=case['groups'], flip_weight=case['flip_weight'],
              flip_filter=case['flip_filter'])
    x = G.t(k + '_x', DEV).requires_grad_(True)
    w = G.t(k + '_w', DEV).requires_grad_(True)
    y = V.conv2d_resample.conv2d_resample(x, w, f=G.t('f::' + case['filter'], DEV), **kw)
    ref = G.t(k + '_y')
    tol = {'float32': 1e-5, 'float64': 1e-6}[case['dtype']]       # fp64: the fast path consumes fp32 weights (modulated_conv2d)
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert rel_err(y, ref) <= tol
    xr = G.t(k + '_x').requires_grad_(True)
    wr = G.t(k + '_w').requires_grad_(True)
    yr = O.conv2d_resample(xr, wr, f=G.t('f::' + case['filter']), **kw)
    dy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(3), dtype=yr.dtype)
    gr = torch.autograd.grad(yr, [xr, wr], dy)
    gg = torch.autograd.grad(y, [x, w], dy.to(DEV))
    for a, b, name in zip(gg, gr, ['dx', 'dw']):
        assert rel_err(a, b) <= tol, name


def test_fma_matches_addcmul_and_unbroadcasts_gradients():
    """fma.fma(a, b, c) = a * b + c with gradients summed back over broadcast dimensions (reference torch_utils/ops/fma.py:15-58; the
    unfused modulated-conv branch calls it with a [N,I,1,1] scale and a [N,1,H,W] / scalar addend)."""
    from vfm_vae_b200.torch_utils.ops import fma as Fm
    g = torch.Generator().manual_seed(1)
    for shapes in [((2, 3, 4, 5), (2, 3, 1, 1), (2, 1, 4, 5)), ((4, 5), (5,), (1,)), ((3, 1, 2), (1, 4, 2), (3, 4, 1)), ((2, 2), (2, 2), (2, 2))]:
        for dtype in (torch.float32, torch.float64):
            a, b, c = (torch.randn(s, generator=g, dtype=dtype).to(DEV).requires_grad_(True) for s in shapes)
            y = Fm.fma(a, b, c)
            ref = a * b + c
            assert y.shape == ref.shape and torch.allclose(y, ref, rtol=1e-6, atol=1e-6)
            dy = torch.randn(ref.shape, generator=g, dtype=dtype).to(DEV)
            gg = torch.autograd.grad(y, [a, b, c], dy)
            gr = torch.autograd.grad(ref, [a, b, c], dy)
            for u, v, t in zip(gg, gr, (a, b, c)):
                assert u.shape == t.shape and torch.allclose(u, v, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------- fused inference layer

@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [
    dict(N=2, I=128, O_=128, H=16, up=1, residual=False),
    dict(N=2, I=128, O_=128, H=32, up=1, residual=True),
    dict(N=2, I=256, O_=128, H=64, up=1, residual=False),     # direct-NCHW forward path (fp16)
    dict(N=1, I=128, O_=128, H=64, up=1, residual=True),
    dict(N=2, I=256, O_=128, H=8, up=2, residual=False),
    dict(N=2, I=128, O_=3, H=32, up=1, residual=False, torgb=True),
], ids=lambda c: f"I{c['I']}O{c['O_']}H{c['H']}up{c['up']}{'res' if c['residual'] else ''}")
def test_fused_layer_matches_oracle(cfg, dtype):
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_modconv_bias_act
    g = torch.Generator().manual_seed(30)
    N, I, O_, H, up = cfg['N'], cfg['I'], cfg['O_'], cfg['H'], cfg['up']
    torgb = cfg.get('torgb', False)
    k = 1 if torgb else 3
    xq = torch.randn(N, I, H, H, generator=g).to(dtype).float()
    w = torch.randn(O_, I, k, k, generator=g) * (0.1 if torgb else 1.0)
    s = (torch.randn(N, I, generator=g) + 1) * (1.0 / math.sqrt(I) if torgb else 1.0)
    b = (torch.randn(O_, generator=g) * 0.2).to(dtype).float()
    noise = None if torgb else torch.randn(H * up, H * up, generator=g) * 0.3
    f = O.setup_filter([1, 3, 3, 1])
    gamma = torch.rand(1, O_, 1, 1, generator=g) + 0.5
    act, gain, clamp = ('linear', 1.0, 256.0) if torgb else ('lrelu', math.sqrt(2) * math.sqrt(0.5), 256.0 * math.sqrt(0.5))
    yr = O.modulated_conv2d(xq, w, s, noise=noise, up=up, padding=k // 2, resample_filter=f if up == 2 else None,
                            demodulate=not torgb, flip_weight=(up == 1))
    yr = O.bias_act(yr, b, act=act, gain=gain, clamp=clamp)
    if cfg['residual']:
        yr = (gamma * yr + xq) * math.sqrt(2)
    with torch.no_grad():
        y = fused_modconv_bias_act(xq.to(DEV, dtype), w.to(DEV), s.to(DEV), b.to(DEV, dtype), noise=noise.to(DEV) if noise is not None else None,
                                   up=up, padding=k // 2, resample_filter=f.to(DEV), demodulate=not torgb, flip_weight=(up == 1), act=act,
                                   gain=gain, clamp=clamp, residual=xq.to(DEV, dtype) if cfg['residual'] else None,
                                   gamma=gamma.to(DEV) if cfg['residual'] else None, res_scale=math.sqrt(2))
    assert y is not None, 'fused path refused a shape it should take'
    assert y.dtype == dtype and y.shape == yr.shape
    assert rel_err(y, yr) <= TOL[str(dtype).split('.')[-1]]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [
    dict(N=2, C=128, H=16, groups=32, residual=True),
    dict(N=3, C=256, H=8, groups=32, residual=True),       # multi-sample tiles: the affine map changes inside a tile
    dict(N=1, C=128, H=64, groups=32, residual=True),
    dict(N=2, C=128, H=32, groups=16, residual=False),     # input map only
], ids=lambda c: f"C{c['C']}H{c['H']}g{c['groups']}{'res' if c['residual'] else ''}")
def test_fused_residual_layer_with_group_norm(cfg, dtype):
    """GroupNorm32 -> modconv -> bias_act -> (gamma*y + x_norm)*sqrt2 (networks/generator.py:261-274) with the normalisation folded
    into the conv (one statistics pass): against the same chain composed from the oracle ops and torch's group_norm in fp32."""
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_modconv_bias_act
    from vfm_vae_b200.plugins import modconv_plugin as P
    g = torch.Generator().manual_seed(31)
    N, Cc, H = cfg['N'], cfg['C'], cfg['H']
    xq = (torch.randn(N, Cc, H, H, generator=g) * 3 + torch.randn(N, Cc, 1, 1, generator=g) * 2).to(dtype).float()
    w = torch.randn(Cc, Cc, 3, 3, generator=g)
    s = torch.randn(N, Cc, generator=g) + 1
    b = (torch.randn(Cc, generator=g) * 0.2).to(dtype).float()
    noise = torch.randn(H, H, generator=g) * 0.3
    gn_w, gn_b = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.3
    gamma = torch.rand(1, Cc, 1, 1, generator=g) + 0.5
    gain, clamp = 1.0, 256.0 * math.sqrt(0.5)
    xn = torch.nn.functional.group_norm(xq, cfg['groups'], gn_w, gn_b, eps=1e-5)
    # the affine form of the statistics, on its own
    sc, sh = P.group_norm_affine(xq.to(DEV, dtype), gn_w.to(DEV), gn_b.to(DEV), cfg['groups'], 1e-5)
    assert rel_err(xq * sc.cpu()[:, :, None, None] + sh.cpu()[:, :, None, None], xn) <= 1e-5
    yr = O.modulated_conv2d(xn, w, s, noise=noise, up=1, padding=1, demodulate=True, flip_weight=True)
    yr = O.bias_act(yr, b, act='lrelu', gain=gain, clamp=clamp)
    if cfg['residual']:
        yr = (gamma * yr + xn) * math.sqrt(2)
    xd = xq.to(DEV, dtype)
    with torch.no_grad():
        y = fused_modconv_bias_act(xd, w.to(DEV), s.to(DEV), b.to(DEV, dtype), noise=noise.to(DEV), up=1, padding=1, act='lrelu', gain=gain,
                                   clamp=clamp, residual=xd if cfg['residual'] else None, gamma=gamma.to(DEV) if cfg['residual'] else None,
                                   res_scale=math.sqrt(2), group_norm=dict(weight=gn_w.to(DEV), bias=gn_b.to(DEV), num_groups=cfg['groups'], eps=1e-5))
    assert y is not None, 'fused path refused a shape it should take'
    assert y.dtype == dtype and y.shape == yr.shape
    assert rel_err(y, yr) <= TOL[str(dtype).split('.')[-1]]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [dict(N=2, C=128, H=16, demod=True), dict(N=3, C=32, H=8, demod=True), dict(N=2, C=128, H=8, demod=False),
                                 dict(N=2, C=128, H=64, demod=True)],      # 64-pixel rows: the forward reads x straight from NCHW
                         ids=lambda c: f"C{c['C']}H{c['H']}{'' if c['demod'] else 'nodemod'}")
def test_modulated_pointwise_conv2d_vs_oracle(cfg, dtype):
    """networks/utils/convnext_utils.py:36 drop-in: 1x1 modulated conv C -> 4C + bias, forward and gradients."""
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import modulated_pointwise_conv2d
    g = torch.Generator().manual_seed(40)
    N, Cc, H = cfg['N'], cfg['C'], cfg['H']
    xq = torch.randn(N, Cc, H, H, generator=g).to(dtype).float()
    w = torch.randn(4 * Cc, Cc, 1, 1, generator=g) * 0.2
    s = torch.randn(N, Cc, generator=g) + 1
    b = torch.randn(1, 4 * Cc, 1, 1, generator=g) * 0.1
    ref_in = [t.clone().requires_grad_(True) for t in (xq, w, s, b)]
    yr = O.modulated_pointwise_conv2d(*ref_in, demodulate=cfg['demod'])
    dy = torch.randn(yr.shape, generator=g)
    gr = torch.autograd.grad(yr, ref_in, dy)
    dev_in = [xq.to(DEV, dtype).requires_grad_(True)] + [t.to(DEV).requires_grad_(True) for t in (w, s, b)]
    y = modulated_pointwise_conv2d(*dev_in, demodulate=cfg['demod'])
    tol = TOL[str(dtype).split('.')[-1]]
    assert rel_err(y, yr) <= tol
    gg = torch.autograd.grad(y, dev_in, dy.to(DEV, y.dtype))
    for name, a, r in zip(['dx', 'dw', 'ds', 'db'], gg, gr):
        assert rel_err(a, r) <= 2 * tol, name


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [dict(N=2, C=128, H=16), dict(N=1, C=128, H=64), dict(N=3, C=256, H=8)], ids=lambda c: f"C{c['C']}H{c['H']}")
def test_fused_convnext_mlp(cfg, dtype):
    """GroupNorm32 -> pwconv1 (modulated, +bias) -> GELU -> pwconv2 (+bias) -> gamma*y + x_in (convnext_utils.py:138-147) as one
    statistics pass + two tensor-core 1x1 convs, against the chain composed from torch / oracle ops in fp32."""
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_convnext_mlp
    g = torch.Generator().manual_seed(41)
    N, Cc, H = cfg['N'], cfg['C'], cfg['H']
    xq = (torch.randn(N, Cc, H, H, generator=g) * 2 + torch.randn(N, Cc, 1, 1, generator=g)).to(dtype).float()
    x_in = torch.randn(N, Cc, H, H, generator=g).to(dtype).float()
    w1 = torch.randn(4 * Cc, Cc, 1, 1, generator=g) * 0.2
    b1 = torch.randn(1, 4 * Cc, 1, 1, generator=g) * 0.1
    w2 = torch.randn(Cc, 4 * Cc, 1, 1, generator=g) * 0.05
    b2 = torch.randn(Cc, generator=g) * 0.1
    s = torch.randn(N, Cc, generator=g) + 1
    gn_w, gn_b = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.3
    gamma = torch.rand(1, Cc, 1, 1, generator=g) + 0.5
    xn = torch.nn.functional.group_norm(xq, 32, gn_w, gn_b, eps=1e-5)
    h = torch.nn.functional.gelu(O.modulated_pointwise_conv2d(xn, w1, s, b1))
    yr = gamma * torch.nn.functional.conv2d(h, w2, b2) + x_in
    with torch.no_grad():
        y = fused_convnext_mlp(xq.to(DEV, dtype), x_in.to(DEV, dtype), gn_w.to(DEV), gn_b.to(DEV), 32, 1e-5, w1.to(DEV), b1.to(DEV), s.to(DEV),
                               w2.to(DEV), b2.to(DEV), gamma.to(DEV))
    assert y is not None and y.dtype == dtype
    assert rel_err(y, yr) <= TOL[str(dtype).split('.')[-1]]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('cfg', [dict(shape=(2, 128, 16, 16), groups=32), dict(shape=(3, 32, 7, 5), groups=8), dict(shape=(1, 512, 8, 8), groups=32),
                                 dict(shape=(2, 64, 64, 64), groups=16),
                                 # backward reduce in warp teams: more channels per group than warps (24, 20), a team count that does not divide 16 (3)
                                 dict(shape=(1, 768, 8, 8), groups=32), dict(shape=(1, 640, 16, 16), groups=32), dict(shape=(2, 96, 12, 12), groups=32),
                                 # N*C = 66560 planes: more than gridDim.y allows (the row kernels index planes through a 1-D grid)
                                 dict(shape=(130, 512, 4, 4), groups=32)],
                         ids=lambda c: 'x'.join(map(str, c['shape'])))
def test_group_norm32_forward_backward(cfg, dtype):
    """GroupNorm32 (fp32 statistics, output in x.dtype) forward and gradients against torch.nn.functional.group_norm in fp64."""
    from vfm_vae_b200.torch_utils.ops.group_norm import group_norm32
    g = torch.Generator().manual_seed(50)
    Cc = cfg['shape'][1]
    xq = (torch.randn(cfg['shape'], generator=g) * 2 + torch.randn(1, Cc, 1, 1, generator=g) * 3).to(dtype).double()
    w, b = torch.rand(Cc, generator=g).double() + 0.5, torch.randn(Cc, generator=g).double()
    dyq = torch.randn(cfg['shape'], generator=g).to(dtype).double()
    ref_in = [t.clone().requires_grad_(True) for t in (xq, w, b)]
    yr = torch.nn.functional.group_norm(ref_in[0], cfg['groups'], ref_in[1], ref_in[2], eps=1e-5)
    gr = torch.autograd.grad(yr, ref_in, dyq)
    x = xq.to(DEV, dtype).requires_grad_(True)
    wd, bd = w.float().to(DEV).requires_grad_(True), b.float().to(DEV).requires_grad_(True)
    y = group_norm32(x, cfg['groups'], wd, bd, 1e-5)
    tol = TOL[str(dtype).split('.')[-1]]
    assert y.dtype == dtype and rel_err(y, yr) <= tol
    gg = torch.autograd.grad(y, [x, wd, bd], dyq.to(DEV, dtype))
    for name, a, r in zip(['dx', 'dweight', 'dbias'], gg, gr):
        assert rel_err(a, r) <= 2 * tol, name


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('shape', [(2, 128, 16, 16), (3, 6, 7, 5), (1, 32, 64, 64), (130, 512, 4, 4)], ids=lambda s: 'x'.join(map(str, s)))
def test_layer_scale_residual(shape, dtype):
    """(gamma*y + x)*sqrt2 of the residual layers (networks/generator.py:272-274), forward and gradients."""
    from vfm_vae_b200.torch_utils.ops.layer_scale import layer_scale_residual
    g = torch.Generator().manual_seed(60)
    yq, xq = (torch.randn(shape, generator=g).to(dtype).double() for _ in range(2))
    gam = (torch.rand(1, shape[1], 1, 1, generator=g) + 0.5).double()
    dq = torch.randn(shape, generator=g).to(dtype).double()
    ref_in = [t.clone().requires_grad_(True) for t in (yq, xq, gam)]
    outr = (ref_in[2] * ref_in[0] + ref_in[1]) * math.sqrt(2)
    gr = torch.autograd.grad(outr, ref_in, dq)
    dev_in = [yq.to(DEV, dtype).requires_grad_(True), xq.to(DEV, dtype).requires_grad_(True), gam.float().to(DEV).requires_grad_(True)]
    out = layer_scale_residual(*dev_in, math.sqrt(2))
    tol = TOL[str(dtype).split('.')[-1]]
    assert out.dtype == dtype and rel_err(out, outr) <= tol
    gg = torch.autograd.grad(out, dev_in, dq.to(DEV, dtype))
    for name, a, r in zip(['dy', 'dx', 'dgamma'], gg, gr):
        assert rel_err(a, r) <= 2 * tol, name


@pytest.mark.parametrize('bias', [True, False])
def test_glue_conv1x1_strict_fp32(bias):
    """decoder.Conv1x1: with TF32 off (the reference's training setting) the glue layers' 1x1 convs run on the tcgen05 split-fp16
    kernels, forward and gradients, at fp32 accuracy; with TF32 allowed they stay on cuDNN."""
    from vfm_vae_b200.decoder import Conv1x1
    from vfm_vae_b200 import _lib
    torch.manual_seed(5)
    conv = Conv1x1(256, 512, 1, bias=bias).to(DEV)
    x = torch.randn(3, 256, 16, 16, device=DEV, requires_grad=True)
    dy = torch.randn(3, 512, 16, 16, device=DEV)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        params = [x, conv.weight] + ([conv.bias] if bias else [])
        xr = x.detach().double().requires_grad_(True)
        wr = conv.weight.detach().double().requires_grad_(True)
        br = conv.bias.detach().double().requires_grad_(True) if bias else None
        yr = torch.nn.functional.conv2d(xr, wr, br)
        gr = torch.autograd.grad(yr, [xr, wr] + ([br] if bias else []), dy.double())
        n0 = _lib.launch_count()
        y = conv(x)
        assert _lib.launch_count() > n0, 'strict-fp32 mode must use the library kernels'
        assert rel_err(y, yr) <= 1e-5
        gg = torch.autograd.grad(y, params, dy)
        for name, a, r in zip(['dx', 'dw', 'db'], gg, gr):
            assert rel_err(a, r) <= 2e-5, name
        torch.backends.cudnn.allow_tf32 = True
        n0 = _lib.launch_count()
        conv(x)
        assert _lib.launch_count() == n0, 'with TF32 allowed the stock cuDNN path is kept'
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_fused_layer_declines_what_it_cannot_do():
    from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_modconv_bias_act
    x = torch.randn(2, 40, 9, 13, device=DEV)            # ragged channels -> generic SIMT path, which does not fuse
    w = torch.randn(24, 40, 3, 3, device=DEV)
    s = torch.ones(2, 40, device=DEV)
    with torch.no_grad():
        assert fused_modconv_bias_act(x, w, s, torch.zeros(24, device=DEV), padding=1) is None
    assert fused_modconv_bias_act(x.requires_grad_(True), w, s, torch.zeros(24, device=DEV), padding=1) is None   # autograd needed -> unfused
