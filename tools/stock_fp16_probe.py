#!/usr/bin/env python
"""Is PyTorch's stock fp16 depthwise 3x3 conv (nn.Conv2d(C, C, 3, padding=1, groups=C), the first op after GroupNorm in the reference's
SeparableUpsampleWithFixedBlur, convnext_utils.py:219, and in its z-convs, generator.py:729-783) reliable on this GPU / torch build?
Compares fp16 (autocast and explicit half) against fp32 over batch sizes, image sizes, cudnn.benchmark and allocator states."""
import torch, torch.nn as nn, torch.nn.functional as F
dev = 'cuda'
def rel(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
def poison():
    bufs = [torch.full([1 << 26], float('nan'), device=dev) for _ in range(16)]
    del bufs
torch.manual_seed(0)
bad = 0
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for C in (512, 128):
        dw = nn.Conv2d(C, C, 3, padding=1, groups=C, bias=False).to(dev)
        for N in (1, 2, 4):
            for H in (16, 32, 64, 128):
                x = torch.randn(N, C, H, H, device=dev)
                with torch.no_grad():
                    y32 = dw(x)
                    for rep in range(3):
                        poison()
                        with torch.autocast('cuda', dtype=torch.float16):
                            y16 = dw(x)
                        yh = F.conv2d(x.half(), dw.weight.half(), padding=1, groups=C)
                        ycl = F.conv2d(x.half().contiguous(memory_format=torch.channels_last), dw.weight.half(), padding=1, groups=C)
                        for tag, y in (('autocast', y16), ('half', yh), ('half channels_last', ycl)):
                            r = rel(y, y32)
                            if not (r < 5e-3):
                                bad += 1
                                print(f'BAD bench={bench} C={C} N={N} H={H} rep={rep} {tag}: finite={bool(torch.isfinite(y).all())} rel={r:.3g}')
print('bad cases:', bad, '| torch', torch.__version__, 'cudnn', torch.backends.cudnn.version())
