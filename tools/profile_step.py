"""Per-kernel GPU time of one decoder step (torch.profiler, CUPTI).  Diagnostic only -- numbers taken under a profiler
are never reported as bench values."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfm_vae_b200.decoder import SynthesisNetwork, F16D32_LEGACY_KWARGS, F16D32_CONVNEXT_KWARGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mode', default='decode')
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--rows', type=int, default=40)
ap.add_argument('--variant', default='legacy')
ap.add_argument('--shapes', action='store_true', help='list aten ops by input shape instead of kernels')
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'
torch.manual_seed(0)
net = SynthesisNetwork(**(F16D32_CONVNEXT_KWARGS if args.variant == 'convnext' else F16D32_LEGACY_KWARGS)).to(dev)
z = torch.randn(args.batch, 512, 16, 16, device=dev)
ws = torch.randn(args.batch, net.num_ws, 512, device=dev)
if args.mode == 'decode':
    net.eval().requires_grad_(False)

    def step():
        with torch.no_grad():
            return net(z, ws)[0]
else:
    def step():
        img, multi = net(z, ws)
        loss = img.square().mean() + sum(m.square().mean() for m in multi)
        loss.backward()
        return img
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=args.shapes) as prof:
    step()
    torch.cuda.synchronize()
if args.shapes:      # aten ops by input shape: which glue op (add / copy_ / mul ...) on which tensor costs the time
    for e in sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)[:args.rows]:
        if e.self_device_time_total > 100:
            print(f'{e.self_device_time_total / 1e3:9.3f} ms  n={e.count:4d}  {e.key[:44]:44s} {str(e.input_shapes)[:150]}')
    sys.exit(0)
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=args.rows, max_name_column_width=70))
