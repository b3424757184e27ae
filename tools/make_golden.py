#!/usr/bin/env python
"""Generate golden input/output vectors from the UNMODIFIED reference (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tools/make_golden.py

Writes small .npz fixtures into tests/golden/.  Every fixture holds the seeded inputs, the call
arguments (as a JSON string) and the reference outputs + gradients produced by the reference's own
``impl='ref'`` CPU path (torch_utils/ops/*.py ``_*_ref`` functions, networks/generator.py).
The reference ships no tests or golden vectors of its own, so these are the pins.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get('VFM_REFERENCE', '/root/reference')
OUT = os.path.join(REPO, 'tests', 'golden')

warnings.filterwarnings('ignore')
sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
sys.path.insert(0, REF)

from torch_utils.ops import bias_act as ref_bias_act  # noqa: E402
from torch_utils.ops import upfirdn2d as ref_upfirdn2d  # noqa: E402
from torch_utils.ops import filtered_lrelu as ref_flrelu  # noqa: E402
from torch_utils.ops import conv2d_resample as ref_c2r  # noqa: E402
from networks import generator as ref_gen  # noqa: E402


def rng(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def randn(g, *shape, dtype=torch.float64):
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)


def save(name, arrays, meta):
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    arrays['meta'] = np.asarray(json.dumps(meta))
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


# ---------------------------------------------------------------------------

def gen_bias_act():
    cases = []
    arrays = {}
    g = rng(1)
    idx = 0
    for dtype in (torch.float32, torch.float64):
        for act in ref_bias_act.activation_funcs.keys():
            for (shape, dim, use_b, gain, clamp, alpha) in [
                ((2, 5, 6, 7), 1, True, None, None, None),
                ((3, 4, 5, 6), 1, True, 0.7, 0.9, 0.3),
                ((4, 6), 1, True, None, 1.5, None),
                ((2, 3, 4, 5), 3, True, 2.0, None, None),
                ((2, 3, 8), 0, False, None, 0.5, None),
            ]:
                x = (randn(g, *shape, dtype=dtype) * 2).requires_grad_(True)
                b = randn(g, shape[dim], dtype=dtype).requires_grad_(True) if use_b else None
                y = ref_bias_act.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref')
                dy = randn(g, *shape, dtype=dtype)
                grads = torch.autograd.grad(y, [x] + ([b] if use_b else []), dy, create_graph=True)
                dx = grads[0]
                # second order: d/d(dy) and d/dx of <dx, ddx>
                ddx = randn(g, *shape, dtype=dtype)
                dy_leaf = dy.clone().requires_grad_(True)
                dx2 = torch.autograd.grad(y, x, dy_leaf, create_graph=True)[0]
                g2 = torch.autograd.grad(dx2, [dy_leaf, x], ddx, allow_unused=True)
                k = f'c{idx}'
                arrays[k + '_x'] = x
                if use_b:
                    arrays[k + '_b'] = b
                    arrays[k + '_db'] = grads[1]
                arrays[k + '_y'] = y
                arrays[k + '_dy'] = dy
                arrays[k + '_dx'] = dx
                arrays[k + '_ddx'] = ddx
                arrays[k + '_g2_dy'] = g2[0]
                arrays[k + '_g2_x'] = g2[1] if g2[1] is not None else torch.zeros_like(x)
                cases.append(dict(key=k, act=act, dim=dim, use_b=use_b, gain=gain, clamp=clamp, alpha=alpha,
                                  dtype=str(dtype).split('.')[-1]))
                idx += 1
    save('bias_act', arrays, dict(cases=cases))


UPFIRDN_CASES = [
    # (shape, filter spec, up, down, padding, flip, gain)
    ((2, 3, 9, 9), ('2d', [1, 3, 3, 1]), 1, 1, [1, 1, 1, 1], False, 4.0),       # post-convT blur of the decoder
    ((2, 3, 8, 8), ('2d', [1, 3, 3, 1]), 2, 1, [2, 1, 2, 1], False, 4.0),       # upsample2d
    ((2, 3, 16, 16), ('2d', [1, 3, 3, 1]), 1, 2, [1, 1, 1, 1], False, 1.0),     # downsample2d
    ((1, 2, 7, 10), ('2d', [1, 2, 3, 4, 5]), [2, 3], [3, 2], [3, 1, 0, 4], True, 1.5),   # asymmetric everything
    ((1, 2, 12, 11), ('2d', [1, -2, 3]), 1, 1, [-2, -1, -1, 0], False, 1.0),    # negative padding = crop
    ((2, 2, 10, 10), ('sep', [1, 4, 6, 4, 1, 2, 3, 1, 2, 1, 1, 3]), 2, 1, [10, 1, 10, 1], False, 4.0),  # separable 12 tap
    ((2, 2, 20, 20), ('sep', [1, 4, 6, 4, 1, 2, 3, 1, 2, 1, 1, 3]), 1, 2, [5, 5, 5, 5], True, 1.0),
    ((1, 1, 5, 5), ('none', None), 2, 1, 0, False, 1.0),                         # f=None
    ((1, 3, 6, 6), ('rect', [[1, 2, 3], [4, 5, 6]]), 1, 1, [1, 1, 0, 1], False, 1.0),    # non-square filter
    ((2, 4, 33, 33), ('2d', [1, 3, 3, 1]), 1, 1, [1, 1, 1, 1], False, 4.0),
    ((1, 2, 6, 6), ('2d', [1, 3, 3, 1]), 4, 1, [3, 3, 3, 3], False, 16.0),       # up=4
]


def _make_filter(spec, dtype=torch.float32):
    kind, taps = spec
    if kind == 'none':
        return None
    if kind == '2d':
        return ref_upfirdn2d.setup_filter(taps, separable=False)
    if kind == 'sep':
        return ref_upfirdn2d.setup_filter(taps, separable=True)
    if kind == 'rect':
        f = torch.tensor(taps, dtype=torch.float32)
        return f / f.sum()
    raise KeyError(kind)


def gen_upfirdn2d():
    arrays, cases = {}, []
    g = rng(2)
    idx = 0
    for dtype in (torch.float32, torch.float64):
        for (shape, fspec, up, down, padding, flip, gain) in UPFIRDN_CASES:
            x = randn(g, *shape, dtype=dtype).requires_grad_(True)
            f = _make_filter(fspec)
            y = ref_upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip, gain=gain, impl='ref')
            dy = randn(g, *y.shape, dtype=dtype)
            dx, = torch.autograd.grad(y, x, dy)
            k = f'c{idx}'
            arrays[k + '_x'] = x
            if f is not None:
                arrays[k + '_f'] = f
            arrays[k + '_y'] = y
            arrays[k + '_dy'] = dy
            arrays[k + '_dx'] = dx
            cases.append(dict(key=k, up=up, down=down, padding=padding, flip=flip, gain=gain, has_f=f is not None,
                              dtype=str(dtype).split('.')[-1]))
            idx += 1
    # helper wrappers
    x = randn(g, 2, 3, 8, 8, dtype=torch.float32)
    f = ref_upfirdn2d.setup_filter([1, 3, 3, 1])
    arrays['h_x'] = x
    arrays['h_f'] = f
    arrays['h_filter2d'] = ref_upfirdn2d.filter2d(x, f, impl='ref')
    arrays['h_upsample2d'] = ref_upfirdn2d.upsample2d(x, f, impl='ref')
    arrays['h_downsample2d'] = ref_upfirdn2d.downsample2d(x, f, impl='ref')
    for spec_name, (taps, kw) in dict(a=([1, 3, 3, 1], {}), b=([1, 2, 1], dict(gain=4)),
                                     c=(list(range(1, 13)), {}), d=([1, 3, 3, 1], dict(flip_filter=True, normalize=False))).items():
        arrays['sf_' + spec_name] = ref_upfirdn2d.setup_filter(taps, **kw)
    save('upfirdn2d', arrays, dict(cases=cases))


FLRELU_CASES = [
    # (shape, fu spec, fd spec, use_b, up, down, padding, gain, slope, clamp, flip)
    ((2, 3, 8, 8), ('sep', list(range(1, 13))), ('sep', [3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5, 8]), True, 2, 2, [10, 11, 10, 11], 2 ** 0.5, 0.2, None, False),
    ((2, 3, 8, 8), ('sep', list(range(1, 13))), ('sep', [3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5, 8]), True, 2, 2, [10, 11, 10, 11], 2 ** 0.5, 0.2, 0.3, False),
    ((1, 2, 9, 7), ('2d', [1, 3, 3, 1]), ('2d', [1, 3, 3, 1]), True, 2, 2, [3, 2, 3, 2], 1.3, 0.1, 0.5, True),
    ((2, 2, 6, 6), ('none', None), ('none', None), True, 1, 1, 0, 2 ** 0.5, 0.2, 0.4, False),
    ((1, 2, 6, 6), ('sep', list(range(1, 9))), ('none', None), False, 2, 1, [4, 3, 4, 3], 1.0, 0.2, None, False),
    ((1, 2, 16, 16), ('none', None), ('sep', list(range(1, 9))), True, 1, 2, [3, 3, 3, 3], 1.0, 0.3, 1.0, False),
    ((1, 2, 5, 5), ('sep', list(range(1, 17))), ('sep', list(range(1, 17))), True, 4, 4, [15, 16, 15, 16], 2 ** 0.5, 0.2, 0.25, False),
    ((1, 2, 6, 6), ('sep', list(range(1, 9))), ('2d', [1, 3, 3, 1]), True, 2, 2, [5, 5, 4, 6], 1.0, 0.2, 0.7, False),
]


def gen_filtered_lrelu():
    arrays, cases = {}, []
    g = rng(3)
    idx = 0
    for dtype in (torch.float32, torch.float64):
        for (shape, fus, fds, use_b, up, down, padding, gain, slope, clamp, flip) in FLRELU_CASES:
            x = randn(g, *shape, dtype=dtype).requires_grad_(True)
            b = (randn(g, shape[1], dtype=dtype) * 0.5).requires_grad_(True) if use_b else None
            fu, fd = _make_filter(fus), _make_filter(fds)
            y = ref_flrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=down, padding=padding, gain=gain,
                                          slope=slope, clamp=clamp, flip_filter=flip, impl='ref')
            dy = randn(g, *y.shape, dtype=dtype)
            grads = torch.autograd.grad(y, [x] + ([b] if use_b else []), dy)
            k = f'c{idx}'
            arrays[k + '_x'] = x
            if use_b:
                arrays[k + '_b'] = b
                arrays[k + '_db'] = grads[1]
            if fu is not None:
                arrays[k + '_fu'] = fu
            if fd is not None:
                arrays[k + '_fd'] = fd
            arrays[k + '_y'] = y
            arrays[k + '_dy'] = dy
            arrays[k + '_dx'] = grads[0]
            cases.append(dict(key=k, use_b=use_b, has_fu=fu is not None, has_fd=fd is not None, up=up, down=down,
                              padding=padding, gain=gain, slope=slope, clamp=clamp, flip=flip,
                              dtype=str(dtype).split('.')[-1]))
            idx += 1
    save('filtered_lrelu', arrays, dict(cases=cases))


MODCONV_CASES = [
    # (N, I, O, H, W, k, up, demod, flip_weight, noise kind, use filter)
    (2, 8, 6, 8, 8, 3, 1, True, True, 'const', False),
    (2, 8, 6, 8, 8, 3, 2, True, False, 'const', True),
    (3, 5, 7, 6, 9, 3, 1, True, True, 'random', False),
    (2, 16, 3, 8, 8, 1, 1, False, True, None, False),      # ToRGB
    (2, 6, 4, 5, 5, 3, 2, True, False, None, True),
    (1, 4, 4, 7, 7, 3, 1, False, False, 'const', False),   # true convolution, no demod
    (2, 4, 5, 6, 6, 1, 2, True, True, None, True),         # 1x1 + up=2 fast path
]


def gen_modconv():
    arrays, cases = {}, []
    g = rng(4)
    idx = 0
    for dtype in (torch.float32, torch.float64):
        for (n, i, o, h, w, k, up, demod, flipw, noise_kind, use_f) in MODCONV_CASES:
            x = randn(g, n, i, h, w, dtype=dtype).requires_grad_(True)
            weight = randn(g, o, i, k, k, dtype=dtype).requires_grad_(True)
            styles = (randn(g, n, i, dtype=dtype) + 1).requires_grad_(True)
            noise = None
            if noise_kind == 'const':
                noise = (randn(g, h * up, w * up, dtype=dtype) * 0.3).requires_grad_(True)
            elif noise_kind == 'random':
                noise = (randn(g, n, 1, h * up, w * up, dtype=dtype) * 0.3).requires_grad_(True)
            f = ref_upfirdn2d.setup_filter([1, 3, 3, 1]) if use_f else None
            y = ref_gen.modulated_conv2d(x=x, weight=weight, styles=styles, noise=noise, up=up, padding=k // 2,
                                         resample_filter=f, demodulate=demod, flip_weight=flipw, fused_modconv=True)
            y2 = ref_gen.modulated_conv2d(x=x, weight=weight, styles=styles, noise=noise, up=up, padding=k // 2,
                                          resample_filter=f, demodulate=demod, flip_weight=flipw, fused_modconv=False)
            assert (y - y2).abs().max() <= 1e-4 * y.abs().max(), 'reference fused/unfused branches disagree'
            dy = randn(g, *y.shape, dtype=dtype)
            leaves = [x, weight, styles] + ([noise] if noise is not None else [])
            grads = torch.autograd.grad(y, leaves, dy)
            kk = f'c{idx}'
            arrays[kk + '_x'], arrays[kk + '_weight'], arrays[kk + '_styles'] = x, weight, styles
            if noise is not None:
                arrays[kk + '_noise'] = noise
                arrays[kk + '_dnoise'] = grads[3]
            if f is not None:
                arrays[kk + '_f'] = f
            arrays[kk + '_y'], arrays[kk + '_dy'] = y, dy
            arrays[kk + '_dx'], arrays[kk + '_dweight'], arrays[kk + '_dstyles'] = grads[0], grads[1], grads[2]
            cases.append(dict(key=kk, up=up, k=k, demodulate=demod, flip_weight=flipw, noise=noise_kind, use_f=use_f,
                              dtype=str(dtype).split('.')[-1]))
            idx += 1
    save('modulated_conv2d', arrays, dict(cases=cases))


def gen_conv2d_resample():
    arrays, cases = {}, []
    g = rng(5)
    f = ref_upfirdn2d.setup_filter([1, 3, 3, 1])
    for idx, (n, i, o, h, k, up, down, pad, flipw) in enumerate([
        (2, 4, 5, 8, 3, 1, 1, 1, True), (2, 4, 5, 8, 3, 2, 1, 1, False), (2, 4, 5, 8, 3, 2, 1, 1, True),
        (1, 3, 2, 8, 3, 1, 2, 1, True), (1, 3, 2, 8, 1, 1, 2, 0, True), (1, 3, 2, 6, 1, 2, 1, 0, True),
        (1, 3, 2, 7, 3, 1, 1, [2, 0, 1, 1], True),
    ]):
        x = randn(g, n, i, h, h, dtype=torch.float64)
        w = randn(g, o, i, k, k, dtype=torch.float64)
        y = ref_c2r.conv2d_resample(x, w, f=f, up=up, down=down, padding=pad, flip_weight=flipw)
        kk = f'c{idx}'
        arrays[kk + '_x'], arrays[kk + '_w'], arrays[kk + '_y'] = x, w, y
        cases.append(dict(key=kk, up=up, down=down, padding=pad, flip_weight=flipw))
    arrays['f'] = f
    save('conv2d_resample', arrays, dict(cases=cases))


def gen_conv2d_resample_ext():
    """conv2d_resample beyond the decoder's subset: groups, down-sampling, up + down, per-axis / asymmetric padding, flip_filter with an
    asymmetric filter, separable 1-D filters (torch_utils/ops/conv2d_resample.py:95-141), fp32 and fp64."""
    arrays, cases = {}, []
    g = rng(15)
    filters = {'f4': ref_upfirdn2d.setup_filter([1, 3, 3, 1]), 'fasym': ref_upfirdn2d.setup_filter([1, 2, 4, 3]),
               'fsep8': ref_upfirdn2d.setup_filter([1, 2, 3, 5, 5, 3, 2, 1])}
    assert filters['fsep8'].ndim == 1
    idx = 0
    for dtype in (torch.float32, torch.float64):
        for (n, i, o, h, k, up, down, pad, flipw, groups, flipf, fname) in [
            (2, 4, 6, 9, 3, 1, 2, 1, True, 2, False, 'f4'), (2, 4, 6, 8, 3, 2, 1, 1, False, 2, False, 'f4'),
            (1, 3, 2, 8, 3, 2, 2, 1, True, 1, False, 'f4'), (1, 3, 2, 8, 3, 1, 1, [1, 2], True, 1, False, 'f4'),
            (1, 3, 2, 8, 3, 2, 1, [1, 0, 2, 1], True, 1, True, 'fasym'), (1, 4, 4, 8, 1, 1, 2, 0, True, 4, False, 'f4'),
            (2, 3, 5, 10, 3, 1, 2, 1, True, 1, True, 'fsep8'), (2, 3, 5, 6, 3, 2, 1, 1, True, 1, False, 'fsep8'),
            (1, 4, 2, 7, 1, 2, 1, 0, False, 2, False, 'fasym'), (1, 2, 3, 9, 3, 1, 1, [0, -1, 2, 0], False, 1, False, 'f4'),
        ]:
            x = randn(g, n, i, h, h, dtype=dtype)
            w = randn(g, o, i // groups, k, k, dtype=dtype)
            y = ref_c2r.conv2d_resample(x, w, f=filters[fname], up=up, down=down, padding=pad, groups=groups, flip_weight=flipw, flip_filter=flipf)
            kk = f'e{idx}'
            idx += 1
            arrays[kk + '_x'], arrays[kk + '_w'], arrays[kk + '_y'] = x, w, y
            cases.append(dict(key=kk, up=up, down=down, padding=pad, flip_weight=flipw, groups=groups, flip_filter=flipf, filter=fname,
                              dtype=str(dtype).split('.')[-1]))
    for k_, v in filters.items():
        arrays['f::' + k_] = v
    save('conv2d_resample_ext', arrays, dict(cases=cases))


DECODER_KW = dict(c_dim=0, w_dim=32, img_resolution=64, img_channels=3, z_resolution=8, z_dim=16,
                  concat_z_block_indices=[0, 1], concat_z_mapped_dims=[32, 32], how_to_process_concat_z='unshuffle',
                  activation_for_concat_z='lrelu', attn_block_indices=[0], attn_depths=[1], use_self_attn=True,
                  use_cross_attn=False, use_convnext=False, use_multiscale_output=True, num_blocks=4, num_fp16_res=2,
                  conv_clamp=256, channel_base=8192, channel_max=32, num_res_blocks=2, architecture='skip')


def gen_decoder():
    """Tiny D-legacy SynthesisNetwork (use_convnext=False): weights, inputs, outputs, a few gradients."""
    torch.manual_seed(6)
    net = ref_gen.SynthesisNetwork(**DECODER_KW)
    # make noise and biases non-trivial so every term is exercised
    g = rng(7)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('noise_strength'):
                p.fill_(0.1)
            elif name.endswith('.bias') and p.ndim == 1 and 'affine' not in name and 'norm' not in name:
                p.copy_(randn(g, *p.shape, dtype=torch.float32) * 0.1)
            elif name.endswith('gamma') and p.ndim == 4:
                p.fill_(0.3)
            elif name.endswith('to_out.weight') or (name.endswith('.3.weight') and '.ff.' in name):
                p.copy_(randn(g, *p.shape, dtype=torch.float32) * 0.05)
    z = randn(g, 2, 16, 8, 8, dtype=torch.float32)
    ws = randn(g, 2, net.num_ws, 32, dtype=torch.float32)
    img, multi = net(z, ws, None, None)
    loss = img.square().mean() + sum(m.square().mean() for m in multi)
    names = ['blocks.3.convs1.3.weight', 'blocks.3.convs1.3.bias', 'blocks.2.conv0.weight', 'blocks.0.convs1.1.gamma',
             'blocks.3.conv0.noise_strength', 'blocks.1.torgb.weight', 'blocks.3.conv0.affine.proj.weight']
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    arrays = {'sd::' + k: v for k, v in net.state_dict().items()}
    arrays['z'], arrays['ws'], arrays['img'] = z, ws, img
    for i, m in enumerate(multi):
        arrays[f'multi{i}'] = m
    for n, gr in zip(names, grads):
        arrays['grad::' + n] = gr
    save('decoder_legacy', arrays, dict(kwargs=DECODER_KW, num_ws=net.num_ws, grad_names=names, loss=float(loss)))


def gen_decoder_convnext():
    """Tiny ConvNeXt-variant SynthesisNetwork (use_convnext=True, what the shipped configs run)."""
    kw = dict(DECODER_KW, use_convnext=True, add_additional_convnext=True, legacy=True, use_gaussian_blur=True)
    torch.manual_seed(8)
    net = ref_gen.SynthesisNetwork(**kw)
    g = rng(9)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('noise_strength'):
                p.fill_(0.1)
            elif name.endswith('gamma') and p.ndim == 4:
                p.fill_(0.3)
            elif (name.endswith('.bias') and 'affine' not in name and 'norm' not in name) or name.endswith('pwconv1.bias'):
                p.copy_(randn(g, *p.shape, dtype=torch.float32) * 0.1)
            elif name.endswith('pwconv1.weight') or name.endswith('pwconv2.weight') or name.endswith('dwconv.weight'):
                p.copy_(randn(g, *p.shape, dtype=torch.float32) * 0.2)
            elif name.endswith('to_out.weight') or (name.endswith('.3.weight') and '.ff.' in name):
                p.copy_(randn(g, *p.shape, dtype=torch.float32) * 0.05)
    z = randn(g, 2, 16, 8, 8, dtype=torch.float32)
    ws = randn(g, 2, net.num_ws, 32, dtype=torch.float32)
    img, multi = net(z, ws, None, None)
    loss = img.square().mean() + sum(m.square().mean() for m in multi)
    names = ['blocks.3.convs1.3.pwconv1.weight', 'blocks.3.convs1.3.pwconv1.bias', 'blocks.2.conv0.dwconv.weight', 'blocks.0.convs1.1.gamma',
             'blocks.3.conv0.noise_strength', 'blocks.1.torgb.weight', 'blocks.3.conv0.affine_pw1.proj.weight', 'blocks.2.seperate_upsample_conv.pointwise.weight']
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    arrays = {'sd::' + k: v for k, v in net.state_dict().items()}
    arrays['z'], arrays['ws'], arrays['img'] = z, ws, img
    for i, m in enumerate(multi):
        arrays[f'multi{i}'] = m
    for n, gr in zip(names, grads):
        arrays['grad::' + n] = gr
    save('decoder_convnext', arrays, dict(kwargs=kw, num_ws=net.num_ws, grad_names=names, loss=float(loss)))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:            # python tools/make_golden.py gen_conv2d_resample_ext  -> only that fixture
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    gen_bias_act()
    gen_upfirdn2d()
    gen_filtered_lrelu()
    gen_modconv()
    gen_conv2d_resample()
    gen_conv2d_resample_ext()
    gen_decoder()
    gen_decoder_convnext()
