#!/usr/bin/env python
"""Debug aid: fill the caching allocator's free blocks with NaN bit patterns, then run the decoder: any kernel that reads memory it (or its
producer) never wrote shows up as a non-finite / different output at the first affected module."""
import contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfm_vae_b200.decoder import SynthesisNetwork, F16D32_LEGACY_KWARGS

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nograd = '--nograd' in sys.argv
torch.manual_seed(3)
net = SynthesisNetwork(**dict(F16D32_LEGACY_KWARGS, img_resolution=res, z_resolution=res // 16)).cuda()
with torch.no_grad():
    for n, p in net.named_parameters():
        if n.endswith('noise_strength'): p.fill_(0.1)
        elif n.endswith('gamma') and p.ndim == 4: p.fill_(0.3)
g = torch.Generator().manual_seed(7)
z = torch.randn(B, 512, res // 16, res // 16, generator=g).cuda(); ws = torch.randn(B, net.num_ws, 512, generator=g).cuda()

def poison(gb=24):
    bufs = [torch.full([1 << 28], float('nan'), dtype=torch.float32, device='cuda') for _ in range(gb)]   # 1 GiB each
    small = [torch.full([n], float('nan'), device='cuda') for n in (256, 4096, 65536, 1 << 20, 1 << 22) for _ in range(64)]
    del bufs, small            # back to the caching allocator (NOT to the driver): later torch.empty() returns these bytes

def run():
    log = []
    hooks = []
    for name, m in net.named_modules():
        if name.startswith('blocks.') and name.count('.') in (1, 2) and not name.endswith(('affine', 'norm')):
            def hook(mod, inp, out, name=name):
                o = out[0] if isinstance(out, tuple) else out
                if torch.is_tensor(o): log.append((name, o.float().clone()))
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad() if nograd else contextlib.nullcontext():
        img, multi = net(z, ws)
        grads = None
        if not nograd:
            names = ['blocks.5.convs1.3.weight', 'blocks.4.conv0.weight', 'blocks.3.convs1.0.weight', 'blocks.5.convs1.2.bias', 'blocks.3.convs1.0.affine.proj.weight', 'blocks.5.torgb.weight',
                     'blocks.4.conv0.noise_strength', 'blocks.2.convs1.1.weight']
            ps = dict(net.named_parameters())
            grads = dict(zip(names, torch.autograd.grad((img.square().mean() + sum(m.square().mean() for m in multi)) * 4096, [ps[n] for n in names])))
    for h in hooks: h.remove()
    return img, log, grads

img0, log0, g0 = run()
torch.cuda.synchronize()
for rep in range(2):
    poison()
    img1, log1, g1 = run()
    torch.cuda.synchronize()
    bad = 0
    for (n, a), (_, b) in zip(log0, log1):
        same = torch.equal(a, b)
        if not same:
            bad += 1
            print(f'  rep {rep} {n:32s} finite {bool(torch.isfinite(b).all())} maxdiff {(a - b).abs().max().item():.4g}')
    if g0 is not None:
        for n in g0:
            d = ((g0[n] - g1[n]).abs().max() / g0[n].abs().max()).item()
            if not (d < 1e-4): print(f'  rep {rep} grad {n}: finite {bool(torch.isfinite(g1[n]).all())} rel {d:.3g}')
    print(f'rep {rep}: res {res} batch {B} nograd {nograd}: {bad} modules differ after poisoning; img equal {bool(torch.equal(img0, img1))}')
