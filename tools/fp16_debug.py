#!/usr/bin/env python
"""Debug aid: per-module output magnitude / first non-finite value of the fp16-blocks decoder (ours and the staged reference)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import reference as R
gen = R.load()
from vfm_vae_b200.decoder import SynthesisNetwork, F16D32_LEGACY_KWARGS
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_benchmark_config_gpu import _perturb

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else 'ours'
perturb = (sys.argv[3] != 'noperturb') if len(sys.argv) > 3 else True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
with contextlib.redirect_stdout(io.StringIO()):
    ref = gen.SynthesisNetwork(**dict(R.F16D32_LEGACY_KWARGS, img_resolution=res, z_resolution=res // 16))
if perturb:
    _perturb(ref, 4)
g = torch.Generator().manual_seed(5)
z = torch.randn(1, 512, res // 16, res // 16, generator=g); ws = torch.randn(1, ref.num_ws, 512, generator=g)

def run(net, fp16):
    log = []
    hooks = []
    for name, m in net.named_modules():
        if name.count('.') in (1, 2) and name.startswith('blocks.') and not name.endswith(('affine', 'norm')):
            def hook(mod, inp, out, name=name):
                o = out[0] if isinstance(out, tuple) else out
                if torch.is_tensor(o):
                    log.append((name, o.float().clone()))
            hooks.append(m.register_forward_hook(hook))
    for b in net.blocks.values():
        b._saved = b.use_fp16
        if not fp16: b.use_fp16 = False
    with torch.no_grad() if '--nograd' in sys.argv else contextlib.nullcontext():
        img, _ = net(z.cuda(), ws.cuda(), None, None)
    for b in net.blocks.values(): b.use_fp16 = b._saved
    for h in hooks: h.remove()
    return img, log

if which == 'ours':
    net = SynthesisNetwork(**dict(F16D32_LEGACY_KWARGS, img_resolution=res, z_resolution=res // 16))
    net.load_state_dict(ref.state_dict()); net = net.cuda()
elif which == 'ref_stock':
    ops = R.ops()
    for mod, names in ((ops.bias_act, ['bias_act']), (ops.upfirdn2d, ['upfirdn2d', 'filter2d', 'upsample2d', 'downsample2d'])):
        for n in names:
            fn = getattr(mod, n); fn.__defaults__ = tuple('ref' if d == 'cuda' else d for d in fn.__defaults__)
    net = ref.cuda()
else:   # ref_ours: reference wrappers on our kernels
    import vfm_vae_b200.integration as integ
    integ.install()
    net = ref.cuda()
img32, log32 = run(net, False)
img16, log16 = run(net, True)
for (n, a), (_, b) in zip(log32, log16):
    d = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-30)
    print(f'{n:32s} max32 {a.abs().max().item():10.4g} max16 {b.abs().max().item():10.4g} finite16 {bool(torch.isfinite(b).all())} rel {d:.3g}')
print('img rel', ((img16 - img32).abs().max() / img32.abs().max()).item())

if which != 'ours':
    # where does the reference's own x_sum branch diverge?  capture the inputs of blocks.3.last_upsample_conv in both modes and re-run it standalone
    cap = {}
    m = net.blocks['3'].last_upsample_conv
    h = m.register_forward_pre_hook(lambda mod, inp: cap.setdefault('in', []).append(inp[0].detach().clone()))
    run(net, False); run(net, True)
    h.remove()
    a, b = cap['in']
    print('input of blocks.3.last_upsample_conv: dtype', a.dtype, b.dtype, 'rel diff between modes', ((a.float() - b.float()).abs().max() / a.float().abs().max()).item())
    with torch.no_grad():
        y32 = m(a)
        with torch.autocast('cuda', dtype=torch.float16):
            y16 = m(a)
            steps = [('norm', m.norm), ('depthwise', m.depthwise), ('pointwise', m.pointwise), ('shuffle', m.shuffle)]
        print('standalone module fp16-autocast vs fp32 on the same input:', ((y16.float() - y32).abs().max() / y32.abs().max()).item(), y16.dtype)
        t32, t16 = a, a
        for nm, f in steps:
            t32 = f(t32)
            with torch.autocast('cuda', dtype=torch.float16):
                t16 = f(t16)
            print('   after', nm, ((t16.float() - t32).abs().max() / t32.abs().max()).item(), t16.dtype, tuple(t16.shape), t16.stride())
        import torch.nn.functional as F
        p32 = F.pad(t32, m.pad, mode=m.pad_mode); p16 = F.pad(t16, m.pad, mode=m.pad_mode)
        print('   after pad', ((p16.float() - p32).abs().max() / p32.abs().max()).item())
        c32 = F.conv2d(p32, m.blur_weight, groups=m.out_channels)
        with torch.autocast('cuda', dtype=torch.float16):
            c16 = F.conv2d(p16, m.blur_weight, groups=m.out_channels)
        print('   after blur', ((c16.float() - c32).abs().max() / c32.abs().max()).item(), 'pad', m.pad, 'blur_weight', tuple(m.blur_weight.shape), m.blur_weight.dtype)
        c16b = F.conv2d(p16.float(), m.blur_weight, groups=m.out_channels)
        print('   blur in fp32 on the fp16 chain input', ((c16b - c32).abs().max() / c32.abs().max()).item())
