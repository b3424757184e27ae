#!/usr/bin/env python
"""Op-level roofline numbers at the decoder's layer shapes: achieved GB/s for the HBM-bound ops and
TFLOP/s for the modulated conv, timed with CUDA events on the launching stream, L2 flushed between iterations.

    python tools/op_bench.py [--batch 64] [--iters 10] [--json out.json]

Algorithmic bytes/FLOPs (the numerators) are the formulas of SURVEY.md 8(d); peaks come from MEASURED_PEAKS.json."""
import argparse
import json
import math
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import vfm_vae_b200 as V  # noqa: E402
from vfm_vae_b200.torch_utils.ops import upfirdn2d as U  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--json', default=None)
ap.add_argument('--only', default='')
args = ap.parse_args()
dev = 'cuda'
N = args.batch
peaks = {}
try:
    peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = peaks.get('hbm_gbs', 6650.0)
TC = peaks.get('bf16_tflops', 1590.0)          # burst figure: kernels are timed alone here
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []


def timeit(fn, iters=args.iters):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes=None, flops=None):
    r = {'op': name, 'ms': round(ms, 4)}
    if nbytes is not None:
        r['GB/s'] = round(nbytes / ms / 1e6, 1)
        r['frac_hbm'] = round(nbytes / ms / 1e6 / HBM, 3)
    if flops is not None:
        r['TFLOP/s'] = round(flops / ms / 1e9, 1)
        r['frac_tc'] = round(flops / ms / 1e9 / TC, 3)
    rows.append(r)
    print(r, flush=True)


def want(name):
    return not args.only or args.only in name


# ---- bias_act: the three fp16 block tensors + an fp32 one -------------------------------------------------------------
for (C, H, dt) in [(128, 256, torch.float16), (256, 128, torch.float16), (512, 64, torch.float16), (512, 32, torch.float32)]:
    if not want('bias_act'):
        break
    es = 2 if dt == torch.float16 else 4
    x = torch.randn(N, C, H, H, device=dev, dtype=dt)
    b = torch.randn(C, device=dev, dtype=dt)
    numel = x.numel()
    ms = timeit(lambda: V.bias_act.bias_act(x, b, act='lrelu', gain=math.sqrt(2), clamp=256.0))
    report(f'bias_act fwd [{N},{C},{H},{H}] {str(dt)[6:]}', ms, nbytes=(2 * numel + C) * es)
    xg = x.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    y = V.bias_act.bias_act(xg, bg, act='lrelu', gain=math.sqrt(2), clamp=256.0)
    dy = torch.randn_like(y)
    ms = timeit(lambda: torch.autograd.grad(y, [xg, bg], dy, retain_graph=True))
    report(f'bias_act bwd(dx+db fused) [{N},{C},{H},{H}] {str(dt)[6:]}', ms, nbytes=3 * numel * es + C * 4)
    del x, xg, y, dy

# ---- upfirdn2d ---------------------------------------------------------------------------------------------------------
f = U.setup_filter([1, 3, 3, 1]).to(dev)
for (C, H, dt) in [(128, 257, torch.float16), (256, 129, torch.float16), (512, 65, torch.float16), (512, 33, torch.float32)]:
    if not want('upfirdn2d'):
        break
    es = 2 if dt == torch.float16 else 4
    x = torch.randn(1, N * C, H, H, device=dev, dtype=dt)
    ms = timeit(lambda: U.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4))
    report(f'upfirdn2d blur 4x4 [1,{N * C},{H},{H}]->{H - 1} {str(dt)[6:]} (dense rows)', ms, nbytes=(x.numel() + N * C * (H - 1) ** 2) * es + 64)
    del x
    Hp = (H + 7) & ~7      # what the modulated conv's up=2 path feeds the blur: rows padded to 16 bytes
    xp = torch.randn(1, N * C, H, Hp, device=dev, dtype=dt)[..., :H]
    ms = timeit(lambda: U.upfirdn2d(xp, f, padding=[1, 1, 1, 1], gain=4))
    report(f'upfirdn2d blur 4x4 [1,{N * C},{H},{H}]->{H - 1} {str(dt)[6:]} (16B-pitched rows)', ms, nbytes=(xp.numel() + N * C * (H - 1) ** 2) * es + 64)
    del xp
if want('upfirdn2d'):
    x = torch.randn(N, 128, 128, 128, device=dev, dtype=torch.float16)
    ms = timeit(lambda: U.upsample2d(x, f))
    report(f'upsample2d [{N},128,128,128]->256 f16', ms, nbytes=x.numel() * 5 * 2)
    x = torch.randn(N, 128, 256, 256, device=dev, dtype=torch.float16)
    ms = timeit(lambda: U.downsample2d(x, f))
    report(f'downsample2d [{N},128,256,256]->128 f16', ms, nbytes=x.numel() * 1.25 * 2)
    del x

# ---- filtered_lrelu (StyleGAN3 shape: separable 12-tap up2/down2, out == in) -------------------------------------------
if want('filtered_lrelu'):
    fu = U.setup_filter([1, 4, 8, 12, 14, 16, 16, 14, 12, 8, 4, 1]).to(dev)
    for (C, H, dt) in [(128, 128, torch.float16), (512, 32, torch.float16), (128, 128, torch.float32)]:
        es = 2 if dt == torch.float16 else 4
        x = torch.randn(N, C, H, H, device=dev, dtype=dt)
        b = torch.randn(C, device=dev, dtype=dt)
        ms = timeit(lambda: V.filtered_lrelu.filtered_lrelu(x, fu, fu, b, up=2, down=2, padding=[10, 11, 10, 11], clamp=256.0), iters=5)
        report(f'filtered_lrelu up2/down2 12-tap [{N},{C},{H},{H}] {str(dt)[6:]} (inference)', ms, nbytes=2 * x.numel() * es + C * es)
        del x

# ---- modulated conv -----------------------------------------------------------------------------------------------------
f4 = U.setup_filter([1, 3, 3, 1]).to(dev)
for (I, O, H, up, dt) in [(128, 128, 256, 1, torch.float16), (256, 256, 128, 1, torch.float16), (512, 512, 64, 1, torch.float16),
                          (256, 128, 128, 2, torch.float16), (640, 512, 32, 2, torch.float16), (512, 512, 32, 1, torch.float32)]:
    if not want('modconv'):
        break
    x = torch.randn(N, I, H, H, device=dev, dtype=dt).requires_grad_(True)
    w = torch.randn(O, I, 3, 3, device=dev, requires_grad=True)
    s = (torch.randn(N, I, device=dev) + 1).requires_grad_(True)
    noise = torch.randn(H * up, H * up, device=dev)
    flops = 2.0 * N * H * H * O * I * 9
    fn = lambda: V.modulated_conv2d(x, w, s, noise=noise, up=up, padding=1, resample_filter=f4, flip_weight=(up == 1))  # noqa: E731
    with torch.no_grad():
        ms = timeit(fn)
    report(f'modulated_conv2d fwd {I}->{O} @{H} up{up} {str(dt)[6:]} (incl. pre-passes)', ms, flops=flops)
    y = fn()
    dy = torch.randn_like(y)
    ms = timeit(lambda: torch.autograd.grad(y, [x, w, s], dy, retain_graph=True), iters=5)
    report(f'modulated_conv2d bwd {I}->{O} @{H} up{up} {str(dt)[6:]} (dx+dw+ds, incl. pre-passes)', ms, flops=2 * flops)
    del x, y, dy

if args.json:
    json.dump({'batch': N, 'hbm_peak_gbs': HBM, 'tc_peak_tflops': TC, 'rows': rows}, open(args.json, 'w'), indent=1)
