#!/usr/bin/env python
"""CONTEXT number, not a bench line: the UNMODIFIED reference decoder (oracle/_ref) on its own stock GPU path on this B200 -- its
JIT-compiled StyleGAN plugins (torch_utils/ops/*.cu, built by its own torch_utils/custom_ops.py) plus cuDNN grouped convs for the
modulated conv (networks/generator.py:46-103) -- on the workload bench.py times (f16d32 D-legacy, 256x256, batch 64/GPU).

    python tools/ref_gpu_bench.py [--batch 64] [--steps 5] [--warmup 3] [--mode decode|train|both] [--json out.json]

Prints one JSON object per mode.  If the reference's plugins do not build on this box (its sources target torch 2.4), the tool says
so (`"unavailable": ...`) and exits 0: SURVEY.md 2.2 names this path as the bar on a GPU box, the tier's reference arm is the CPU path.
Nothing of this repo's kernels is on the path: vfm_vae_b200 is not imported.
"""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import reference  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--mode', default='both')
ap.add_argument('--json', default=None)
args = ap.parse_args()
dev = 'cuda'
# as the reference's training loop sets them (training/training_loop.py:503-506)
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

out = []
loader = "the reference's own torch_utils/custom_ops.py"
try:
    gen = reference.load()
    t0 = time.time()
    from torch_utils.ops import bias_act, upfirdn2d      # the staged reference's modules
    from torch_utils import custom_ops
    try:
        bias_act._init()
        upfirdn2d._init()
    except Exception as e1:     # noqa: BLE001
        # The reference pins torch 2.4; under this image's torch 2.11 its loader compiles the plugins and then fails at
        # `importlib.import_module(module_name)` (custom_ops.py:141: cpp_extension.load no longer leaves the module importable by name).
        # The SOURCES are fine: tools/build_ref_plugins.py compiles them unmodified (same flags) into oracle/_ref_plugins/ in the build
        # container, and this harness hands the modules to the reference through its own plugin cache (custom_ops._cached_plugins) --
        # no reference file is edited, its kernels and Python wrappers run as they are.
        import importlib.machinery
        import importlib.util
        for name in ('bias_act_plugin', 'upfirdn2d_plugin'):
            so = os.path.join(REPO, 'oracle', '_ref_plugins', name, name + '.so')
            spec = importlib.util.spec_from_loader(name, importlib.machinery.ExtensionFileLoader(name, so))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            custom_ops._cached_plugins[name] = mod
        bias_act._plugin = None
        upfirdn2d._plugin = None
        bias_act._init()
        upfirdn2d._init()
        loader = f"pre-built from the unmodified sources, injected through custom_ops._cached_plugins (its own loader fails under torch {torch.__version__}: {type(e1).__name__}: {str(e1)[:120]})"
    build_s = time.time() - t0
except Exception as e:     # noqa: BLE001
    line = {'impl': 'reference-stock-gpu', 'unavailable': f'{type(e).__name__}: {str(e)[:300]}'}
    print(json.dumps(line))
    if args.json:
        json.dump([line], open(args.json, 'w'))
    sys.exit(0)

torch.manual_seed(0)
net = gen.SynthesisNetwork(**reference.F16D32_LEGACY_KWARGS).to(dev)
z = torch.randn(args.batch, 512, 16, 16, device=dev)
ws = torch.randn(args.batch, net.num_ws, 512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(step):
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def decode():
    with torch.no_grad():
        r = net(z, ws, None, None)
    return r


opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.0, 0.99))


def train():
    opt.zero_grad(set_to_none=True)
    r = net(z, ws, None, None)
    img = r[0] if isinstance(r, (tuple, list)) else r
    multi = r[1] if isinstance(r, (tuple, list)) and len(r) > 1 and isinstance(r[1], (tuple, list)) else []
    loss = img.float().square().mean() + sum(m.float().square().mean() for m in multi)
    loss.backward()
    opt.step()


for mode, fn in (('decode', decode), ('train', train)):
    if args.mode not in (mode, 'both'):
        continue
    try:
        ms = timed(fn)
        line = {'impl': 'reference-stock-gpu', 'mode': mode, 'value': args.batch / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms, 'batch': args.batch,
                'steps': args.steps, 'warmup': args.warmup, 'plugin_setup_s': round(build_s, 1), 'plugin_loader': loader,
                'what': 'unmodified reference SynthesisNetwork(use_convnext=False, f16d32) on CUDA: its own JIT plugins (bias_act, upfirdn2d) + '
                        'cuDNN grouped conv for the modulated conv, cudnn.benchmark on, TF32 off as in its training loop; CUDA events, L2 flushed'}
    except Exception as e:     # noqa: BLE001
        line = {'impl': 'reference-stock-gpu', 'mode': mode, 'unavailable': f'{type(e).__name__}: {str(e)[:300]}'}
        torch.cuda.empty_cache()
    print(json.dumps(line))
    out.append(line)
if args.json:
    json.dump(out, open(args.json, 'w'), indent=1)
