#!/usr/bin/env python
"""Runs one modulated-conv layer shape a few times (for ncu captures and quick timing of a single launch).

    python tools/conv_probe.py --cin 128 --cout 128 --res 256 [--up 2] [--mode fwd|fused|fused_res|bwd] [--dtype f16|f32] [--iters 3]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_vae_b200 as V  # noqa: E402
from vfm_vae_b200.torch_utils.ops import upfirdn2d as U  # noqa: E402
from vfm_vae_b200.torch_utils.ops.modulated_conv2d import fused_modconv_bias_act  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--cin', type=int, default=128)
ap.add_argument('--cout', type=int, default=128)
ap.add_argument('--res', type=int, default=256, help='input resolution')
ap.add_argument('--up', type=int, default=1)
ap.add_argument('--k', type=int, default=3)
ap.add_argument('--act', default='lrelu')
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--mode', default='fwd')
ap.add_argument('--dtype', default='f16')
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--warm', type=int, default=2, help='untimed warm-up runs (0 for ncu captures)')
a = ap.parse_args()
dev = 'cuda'
dt = torch.float16 if a.dtype == 'f16' else torch.float32
torch.manual_seed(0)
x = torch.randn(a.batch, a.cin, a.res, a.res, device=dev, dtype=dt)
w = torch.randn(a.cout, a.cin, a.k, a.k, device=dev)
s = torch.randn(a.batch, a.cin, device=dev) + 1
b = torch.randn(a.cout, device=dev, dtype=dt)
ro = a.res * a.up
noise = torch.randn(ro, ro, device=dev) if a.k == 3 else None
gamma = torch.full([1, a.cout, 1, 1], 1e-5, device=dev)
f4 = U.setup_filter([1, 3, 3, 1]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
kw = dict(noise=noise, up=a.up, padding=a.k // 2, resample_filter=f4, flip_weight=(a.up == 1))


def run():
    if a.mode == 'fwd':
        with torch.no_grad():
            return V.modulated_conv2d(x, w, s, **kw)
    if a.mode in ('fused', 'fused_res'):
        with torch.no_grad():
            res = x if (a.mode == 'fused_res' and a.cin == a.cout and a.up == 1) else None
            return fused_modconv_bias_act(x, w, s, b, act=a.act, gain=1.0, clamp=181.0 if a.act == 'lrelu' else None, residual=res, gamma=gamma if res is not None else None,
                                          res_scale=math.sqrt(2), **kw)
    xg = x.detach().requires_grad_(True)
    wg = w.detach().requires_grad_(True)
    sg = s.detach().requires_grad_(True)
    y = V.modulated_conv2d(xg, wg, sg, **kw)
    return torch.autograd.grad(y, [xg, wg, sg], torch.ones_like(y))


for _ in range(a.warm):
    run()
ts = []
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f'{a.mode} {a.cin}->{a.cout} @{a.res} up{a.up} {a.dtype}: {sorted(ts)[len(ts) // 2]:.4f} ms (whole op incl. pre-passes)')
