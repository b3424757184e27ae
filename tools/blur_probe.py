#!/usr/bin/env python
"""One decoder-sized blur (upfirdn2d up=down=1, 4x4) on 32-byte-pitched rows, optionally with the fused noise/bias/lrelu
epilogue -- for ncu captures of upfirdn2d_blur."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfm_vae_b200.torch_utils.ops import upfirdn2d as U  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--planes', type=int, default=8192)
ap.add_argument('--res', type=int, default=257)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--dtype', default='f16')
a = ap.parse_args()
dt = torch.float16 if a.dtype == 'f16' else torch.float32
H = a.res
Hp = (H + 15) & ~15
x = torch.randn(1, a.planes, H, Hp, device='cuda', dtype=dt)[..., :H]
f = U.setup_filter([1, 3, 3, 1]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(2):
    U.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)
ts = []
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = U.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
nbytes = (a.planes * H * H + y.numel()) * x.element_size()
print(f'blur [{a.planes},{H},{H}] {a.dtype}: {ms:.4f} ms  {nbytes / ms / 1e6:.1f} GB/s')
