#!/usr/bin/env python
"""bench.py -- images/sec of the VFM-VAE f16d32 pixel-decoder hot path on B200, and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|decode] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input:
  train  (default, the line's `value`): forward + backward through the D-legacy SynthesisNetwork + the gradient mean over ranks
          (vfm_vae_b200.sync.GradExchange: bucketed all_reduce overlapped with backward, reference semantics of
          training/training_loop.py:272-289) + Adam step, 64 images per GPU  (BASELINE configs[2] restricted to the hot path)
  decode (secondary block `decode`, and `decode512` = BASELINE configs[4]: 512x512, 32 images per GPU): the forward alone
          (z [B,512,16,16], ws [B,36,512] -> img [B,3,256,256] + 5 multi-scale images)
Workload: f16d32, num_fp16_res=3 (fp16 blocks 3-5, fp32 blocks 0-2), random-init weights, synthetic latents.  The frozen
SigLIP2 encoder / losses of the configs are outside the hot-path scope.  D-legacy = use_convnext=False, the
variant north_star describes (the shipped YAMLs run the ConvNeXt variant, which calls none of these ops: SURVEY.md 0.2).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM and per-launch timing events
OFF; `e2e` = the same metric through the public Python API with pinned HOST buffers, H2D of the inputs and D2H of the result
inside the timed region; `roofline` / `kernels` come from a separate pass with per-launch CUDA events; `cpu_baseline` and
`--impl reference` time the UNMODIFIED reference decoder (oracle/_ref, staged by tools/stage_reference.py) on the host cores;
`parity` compares the CUDA path with that reference on the same weights and inputs before anything is timed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--mode', choices=['train', 'decode'], default='train',
                    help="what `value` measures; 'train' (default) also reports the decode / decode512 blocks unless --no-secondary")
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per step (default 64; 32 at --res 512)')
    ap.add_argument('--res', type=int, default=256, choices=[256, 512])
    ap.add_argument('--fp16-res', type=int, default=3, help='num_fp16_res (3 = training configs, 0 = the inference tools)')
    ap.add_argument('--cpu-batch', type=int, default=None, help='sample size (images per step) of the CPU legs; default: bounded by the step count')
    ap.add_argument('--variant', choices=['legacy', 'convnext'], default='legacy',
                    help="decoder variant: 'legacy' = use_convnext=False (the path north_star names, default), 'convnext' = the shipped YAMLs' layers")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the decode / decode512 blocks of the default (train) line')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--cuda-graph', choices=['auto', 'on', 'off'], default='auto',
                    help='decode: replay the step as ONE captured CUDA graph (SynthesisNetwork.decode_graph); auto = fall back to eager if capture fails')
    ap.add_argument('--sync', choices=['overlap', 'posthoc'], default='overlap',
                    help="train: 'overlap' = GradExchange (buckets leave during backward); 'posthoc' = the reference-style sync_grads after backward")
    ap.add_argument('--no-cudnn-benchmark', action='store_true', help='leave torch.backends.cudnn.benchmark off for the glue layers')
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 32 if a.res == 512 else 64
    return a


def decoder_kwargs(variant, res, fp16_res):
    from vfm_vae_b200.decoder import F16D32_LEGACY_KWARGS, F16D32_CONVNEXT_KWARGS
    kw = dict(F16D32_CONVNEXT_KWARGS if variant == 'convnext' else F16D32_LEGACY_KWARGS)
    kw['img_resolution'] = res
    kw['z_resolution'] = res // 16
    kw['num_fp16_res'] = fp16_res
    return kw


def make_inputs(kw, batch, num_ws, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(batch, kw['z_dim'], kw['z_resolution'], kw['z_resolution'], generator=g)
    ws = torch.randn(batch, num_ws, kw['w_dim'], generator=g)
    return z, ws


def workload_name(variant, mode, res, batch, fp16_res):
    v = 'D-legacy pixel decoder (SynthesisNetwork use_convnext=False)' if variant == 'legacy' else 'D-convnext pixel decoder (SynthesisNetwork use_convnext=True)'
    what = {'train': 'train step (fwd + bwd + gradient mean over ranks + Adam)', 'decode': 'decode (forward)'}[mode]
    cfg = 'BASELINE configs[2]' if mode == 'train' else ('BASELINE configs[4]' if res == 512 else 'BASELINE configs[1]')
    return (f'f16d32 {v} {what}, {res}x{res}, batch {batch}/GPU, num_fp16_res={fp16_res}, random-init weights, synthetic latents '
            f'({cfg} restricted to the hot path; SigLIP2 encoder and losses out of scope)')


def loss_fn(img, multi):
    # synthetic stand-in for the recon losses; x 4096 keeps d(loss)/d(img) (~5e-6 for a plain mean) out of the fp16 subnormal range so that
    # the fp16 backward is numerically meaningful, as it is under the reference's real loss weights
    return (img.square().mean() + sum(m.square().mean() for m in multi)) * 4096.0


# ----------------------------------------------------------------------------------------------------------- CPU legs
# The reference's own CPU path: oracle/_ref holds the UNMODIFIED reference packages (tools/stage_reference.py); CPU tensors take
# its impl='ref' branch automatically (torch_utils/ops/bias_act.py:84, upfirdn2d.py:160).  Only this section touches oracle/.

def reference_net(variant, res, state_dict=None):
    """-> (net, kind): the reference SynthesisNetwork on CPU (kind 'reference'), or the oracle port if nothing is staged ('port')."""
    from oracle import reference as R
    torch.manual_seed(0)
    if R.available():
        gen = R.load()
        kw = dict(R.F16D32_CONVNEXT_KWARGS if variant == 'convnext' else R.F16D32_LEGACY_KWARGS)
        kw.update(img_resolution=res, z_resolution=res // 16)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its layer plan while constructing
            net = gen.SynthesisNetwork(**kw)
        kind = 'reference'
    else:
        from types import SimpleNamespace
        from oracle import ref_ops as O
        from vfm_vae_b200.decoder import SynthesisNetwork

        def modconv(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True, flip_weight=True, fused_modconv=True):
            return O.modulated_conv2d(x, weight, styles, noise=noise, up=up, down=down, padding=padding, resample_filter=resample_filter,
                                      demodulate=demodulate, flip_weight=flip_weight)
        ops = SimpleNamespace(bias_act=O.bias_act, def_gain=lambda a: O.ACTIVATIONS[a][1], setup_filter=O.setup_filter, upsample2d=O.upsample2d,
                              modulated_conv2d=modconv, modulated_pointwise_conv2d=O.modulated_pointwise_conv2d)
        net = SynthesisNetwork(ops=ops, **decoder_kwargs(variant, res, 0))
        kind = 'port'
    if state_dict is not None:
        net.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()}, strict=True)
    return net, kind


def cpu_steps(net, mode, z, ws):
    """-> step() running the reference decoder's `mode` step on CPU; for train it returns (img, {name: grad})."""
    if mode == 'decode':
        net.eval().requires_grad_(False)

        def step():
            with torch.no_grad():
                return net(z, ws, None, None)[0], None
        return step
    net.train().requires_grad_(True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.0, 0.99))

    def step(apply=True):
        img, multi = net(z, ws, None, None)
        loss = loss_fn(img, multi)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        # the reference's sync_grads at world size 1 (training/training_loop.py:281-289): concat -> nan_to_num -> split back
        params = [p for p in net.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.flatten() for p in params])
        torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
        for p, g in zip(params, flat.split([p.numel() for p in params])):
            p.grad = g.reshape(p.size()).to(p.dtype)
        grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        if apply:
            opt.step()
        return img.detach(), grads
    return step


def default_cpu_batch(args):
    if args.cpu_batch is not None:
        return args.cpu_batch
    n = args.steps + args.warmup
    return 4 if n <= 8 else (2 if n <= 16 else 1)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = default_cpu_batch(args)
    net, kind = reference_net(args.variant, args.res)
    z, ws = make_inputs(dict(z_dim=net.z_dim, w_dim=net.w_dim, z_resolution=args.res // 16), batch, net.num_ws, seed=1)   # no vfm_vae_b200 import on this arm
    step = cpu_steps(net, args.mode, z, ws)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    what = ("the UNMODIFIED reference SynthesisNetwork (oracle/_ref, impl='ref' CPU path, MKLDNN convs)" if kind == 'reference'
            else 'oracle port (oracle/ref_ops.py) of the decoder: the reference is not staged')
    sample = f'{what}, torch CPU fp32, {cores} threads, {batch} images per step, {args.steps} timed steps after {args.warmup} warm-ups'
    _emit(json.dumps({
        'impl': 'reference', 'metric': f'images/sec ({args.mode})', 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.variant, args.mode, args.res, args.batch, args.fp16_res), 'cpu_sample_batch': batch},
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = b.abs().max().item()
    return (a - b).abs().max().item() / d if d > 0 else (a - b).abs().max().item()


PARITY_GRADS = ('blocks.5.convs1.3.weight', 'blocks.5.conv0.weight', 'blocks.4.convs1.0.weight', 'blocks.3.conv0.weight', 'blocks.2.convs1.1.weight',
                'blocks.0.conv0.weight', 'blocks.5.convs1.2.bias', 'blocks.3.convs1.0.affine.proj.weight', 'blocks.5.torgb.weight')


def cpu_baseline_and_parity(args, net, dev):
    """Rank 0, N=1: time the reference's CPU step on a bounded sample AND use its outputs / gradients as the parity check of the
    CUDA path on the same weights and inputs (before anything is timed).  -> (cpu_baseline dict, parity dict)"""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # TF32 off, as the reference's training loop sets it (training/training_loop.py:504-505): the fp32 parity numbers below would
    # otherwise measure cuDNN's TF32 glue convs (z-convs, attention projections), not the kernels
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    batch = args.cpu_batch if args.cpu_batch is not None else 2
    ref, kind = reference_net(args.variant, args.res, net.state_dict())
    z, ws = make_inputs(decoder_kwargs(args.variant, args.res, args.fp16_res), batch, net.num_ws, seed=7)
    out = {}
    # decode sample (also the image parity reference)
    dstep = cpu_steps(ref, 'decode', z, ws)
    t0 = time.perf_counter()
    img_ref, _ = dstep()
    t_dec = time.perf_counter() - t0
    out['decode'] = {'value': batch / t_dec, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                     'sample': f"reference SynthesisNetwork forward (impl='ref', torch CPU fp32), one pass over {batch} images ({t_dec:.1f} s)"}
    parity = {'metric': 'max|a-b|/max|b| vs the reference CPU fp32 path, same weights and inputs', 'sample_images': batch, 'reference_kind': kind}
    zd, wd = z.to(dev), ws.to(dev)
    fp16_flags = [b.use_fp16 for b in net.blocks.values()]

    def set_fp16(flags):          # all False = the tools' num_fp16_res=0 configuration (force_fp32 alone keeps block 3's z-convs under fp16 autocast)
        for b, f in zip(net.blocks.values(), flags):
            b.use_fp16 = f
    with torch.no_grad():
        net.eval()
        img = net(zd, wd)[0]
        set_fp16([False] * len(fp16_flags))
        img32 = net(zd, wd)[0]
        set_fp16(fp16_flags)
    parity['image_fp16_blocks'] = rel_err(img, img_ref)
    parity['image_fp32'] = rel_err(img32, img_ref)
    if args.mode == 'train':
        tstep = cpu_steps(ref, 'train', z, ws)
        t0 = time.perf_counter()
        _, g_ref = tstep(apply=False)
        t_tr = time.perf_counter() - t0
        out['train'] = {'value': batch / t_tr, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                        'sample': f"reference SynthesisNetwork forward + backward + sync_grads(world 1) (impl='ref', torch CPU fp32), one pass over {batch} images ({t_tr:.1f} s; Adam update not included)"}
        net.train().requires_grad_(True)
        for name, flags in (('grads_fp16_blocks', fp16_flags), ('grads_fp32', [False] * len(fp16_flags))):
            net.zero_grad(set_to_none=True)
            set_fp16(flags)
            i2, m2 = net(zd, wd)
            loss_fn(i2, m2).backward()
            gp = dict(net.named_parameters())
            parity[name] = {n: rel_err(gp[n].grad, g_ref[n]) for n in PARITY_GRADS if n in g_ref and gp[n].grad is not None}
        set_fp16(fp16_flags)
        net.zero_grad(set_to_none=True)
    # whole-network bounds (tests/test_benchmark_config_gpu.py holds the per-op gates 1e-5 / 2e-3 for every layer shape of this network, and
    # explains the whole-network bounds): images 1e-4 all-fp32, 5e-3 with the fp16 blocks; gradients are reported
    tol16, tol32 = 5e-3, 1e-4
    parity['tol'] = {'image_fp16_blocks': tol16, 'image_fp32': tol32,
                     'note': 'whole-network bounds; per-op gates (1e-5 fp32 / 2e-3 fp16, north_star) are held per layer shape in tests/test_benchmark_config_gpu.py. '
                             'Random-init weights here (layer scales 1e-5, noise strength 0), so gradients of the residual layers are tiny and their fp16 relative error is large by construction.'}
    parity['max_rel'] = parity['image_fp16_blocks']
    parity['ok'] = bool(parity['image_fp16_blocks'] <= tol16 and parity['image_fp32'] <= tol32)
    del ref
    return out, parity


# ------------------------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------------------- ours

def timing_report():
    import ctypes as C
    from vfm_vae_b200 import _lib
    lib = _lib.load()
    buf = (_lib.KernelStat * 512)()
    n = min(lib.vfm_timing_report(buf, 512), 512)
    return [dict(name=buf[i].name.decode(), launches=int(buf[i].launches), total_ms=buf[i].total_ms, flops=buf[i].flops, bytes=buf[i].bytes)
            for i in range(n)]


_NCU_KEYS = {'modconv_tc_fwd': 'conv_tc_kernel<__half, 0, 0', 'modconv_tc_fwd_split': 'conv_tc_kernel<float, 0, 1', 'modconv_tc_dgrad': 'conv_tc_kernel<__half, 1, 0',
             'modconv_tc_wgrad': 'wgrad_tc_kernel<0', 'modconv_nhwc_prepass': 'nhwc_prepass_kernel<__half', 'upfirdn2d_blur': 'upfirdn2d_blur<__half'}


def ncu_traffic(name, mode):
    """DRAM read+write bytes per launch of `name` from the committed `ncu --set full` capture of this command (profiles/r02_traffic_<mode>.json,
    written by tools/ncu_summarize.py); None when no capture covers the kernel."""
    key = _NCU_KEYS.get(name.split(':')[0])
    for fn in (f'r02_traffic_{mode}.json', 'r01c_traffic.json' if mode == 'decode' else None):
        if fn is None:
            continue
        try:
            tr = json.load(open(os.path.join(REPO, 'profiles', fn)))
        except Exception:
            continue
        sel = [v for k, v in tr.items() if key and key in k]
        n = sum(v['launches'] for v in sel)
        if n:
            return sum(v['launches'] * v['dram_bytes_per_launch'] for v in sel) / n, fn
    return None, None


def kernel_table(stats, ms_total, mode, traffic_applies):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks.get('hbm_gbs'), 'MEASURED_PEAKS.json hbm_gbs') if peaks.get('hbm_gbs') else (6650.0, 'fallback (B200_PROFILING.md)')
    tc_peak, tc_src = ((peaks.get('bf16_tflops_sustained'), 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)')
                       if peaks.get('bf16_tflops_sustained') else (1400.0, 'fallback (B200_PROFILING.md)'))
    stats = sorted(stats, key=lambda s: -s['total_ms'])
    kernels = []
    for s in stats:
        per = s['total_ms'] / max(s['launches'], 1)
        entry = {'name': s['name'], 'launches': s['launches'], 'total_ms': round(s['total_ms'], 3), 'avg_ms': round(per, 4)}
        if s['flops'] > 0:
            entry['tflops'] = round(s['flops'] / (s['total_ms'] * 1e-3) / 1e12, 3)
        if s['bytes'] > 0:
            entry['gbs'] = round(s['bytes'] / (s['total_ms'] * 1e-3) / 1e9, 1)
        kernels.append(entry)
    roofline = None
    if stats:
        top = stats[0]
        tr, src = ncu_traffic(top['name'], mode) if traffic_applies else (None, None)
        common = {'kernel': top['name'], 'traffic': tr, 'traffic_unit': f'bytes/launch (dram read+write, ncu --set full, profiles/{src})' if src else None,
                  'share_of_step': top['total_ms'] / ms_total if ms_total else None, 'launches': top['launches']}
        if top['flops'] > 0:
            ach = top['flops'] / (top['total_ms'] * 1e-3) / 1e12
            roofline = dict(common, bound='tensor', achieved=ach, peak=tc_peak, unit='TFLOP/s', frac=ach / tc_peak,
                            algorithmic_per_launch=top['flops'] / max(top['launches'], 1), peak_source=tc_src)
        else:
            ach = top['bytes'] / (top['total_ms'] * 1e-3) / 1e9
            roofline = dict(common, bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s', frac=ach / hbm_peak,
                            algorithmic_per_launch=top['bytes'] / max(top['launches'], 1), peak_source=hbm_src)
    return roofline, kernels


class Ctx:
    pass


def measure(cx, net, mode, res, batch, steps, warmup, table_steps, graph_mode, sync_mode):
    """One workload: warm up, device-timed pass (timing events off), end-to-end pass (host buffers), kernel-table pass.  -> dict"""
    import torch.distributed as dist
    from vfm_vae_b200 import _lib, sync
    dev, world, rank, lib = cx.dev, cx.world, cx.rank, cx.lib
    kw = cx.kw(res)
    z_h, ws_h = make_inputs(kw, batch, net.num_ws, seed=1 + rank)
    z_h, ws_h = z_h.pin_memory(), ws_h.pin_memory()
    z, ws = z_h.to(dev), ws_h.to(dev)
    flush = cx.flush
    sync_events = []
    exchange = None
    graph_used = False

    if mode == 'decode':
        torch.backends.cudnn.allow_tf32 = True          # the reference's decode / reconstruct tools leave PyTorch's defaults untouched
        torch.backends.cuda.matmul.allow_tf32 = False
        net.eval().requires_grad_(False)

        def eager(z_, ws_):
            with torch.no_grad():
                return net(z_, ws_)[0]
        step = eager
        if graph_mode in ('on', 'auto') and not os.environ.get('VFM_CUDA_PROFILER_RANGE'):
            try:
                ref_img = eager(z, ws).clone()
                l0 = _lib.launch_count()
                got = net.decode_graph(z, ws)[0]
                torch.cuda.synchronize()
                if not torch.equal(got, ref_img):       # same kernels, same inputs: the replay must reproduce the eager result
                    raise RuntimeError('graph replay differs from the eager step')
                graph_used = True
                del ref_img

                def step(z_, ws_):
                    return net.decode_graph(z_, ws_)[0]
            except Exception as e:                        # noqa: BLE001
                if graph_mode == 'on':
                    raise
                print(f'[bench] CUDA graph capture unavailable ({e}); timing the eager step', file=sys.stderr)
                net.__dict__.pop('_decode_graphs', None)
                step = eager
        result_bytes = batch * 3 * res * res * 4
        out_h = torch.empty(batch, 3, res, res, dtype=torch.float32).pin_memory()
    else:
        # same numerics switches as the reference's training loop: TF32 off for the fp32 glue layers (training/training_loop.py:504-505)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        net.train().requires_grad_(True)
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4, betas=(0.0, 0.99))       # the configs' G_opt_kwargs (stage_1 YAML:147-151)
        if sync_mode == 'overlap':
            exchange = sync.GradExchange(params)

        def step(z_, ws_):
            img, multi = net(z_, ws_)
            loss = loss_fn(img, multi)
            if exchange is not None:
                exchange.zero_grad()
                loss.backward()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                exchange.finish()
                e1.record()
            else:
                opt.zero_grad(set_to_none=True)
                loss.backward()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sync.sync_grads(params)
                e1.record()
            sync_events.append((e0, e1))
            opt.step()
            return loss.detach()
        result_bytes = 4
        out_h = torch.empty([], dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib.vfm_timing_enable(0)
    for _ in range(max(warmup, 3)):
        step(z, ws)
        flush.zero_()
    barrier()

    # ---- device-timed region: inputs resident in HBM, no per-launch events ----
    sampler = ClockSampler(cx.local)
    if rank == 0:
        sampler.start()
    sync_events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = bool(os.environ.get('VFM_CUDA_PROFILER_RANGE')) and mode == os.environ.get('VFM_CUDA_PROFILER_RANGE')
    launches0 = _lib.launch_count()
    barrier()
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(steps):
        step(z, ws)
        flush.zero_()              # L2 flush between iterations (256 MiB write, ~0.05 ms)
    e1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    if graph_used:
        launches = cx.graph_launches(net, z, ws) * steps
    clocks = sampler.stop() if rank == 0 else None
    sync_ms = None
    if sync_events:
        sync_ms = sum(a.elapsed_time(b) for a, b in sync_events) / len(sync_events)
    t = torch.tensor([ms, sync_ms or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, sync_ms_max = t.tolist()

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H of the result, every step ----
    barrier()
    e0.record()
    for _ in range(steps):
        zd = z_h.to(dev, non_blocking=True)
        wd = ws_h.to(dev, non_blocking=True)
        r = step(zd, wd)
        out_h.copy_(r.detach(), non_blocking=True)
        flush.zero_()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()

    # ---- kernel table / roofline: the same step with per-launch CUDA events, OUTSIDE `value` (always the eager step) ----
    if graph_used:
        step = eager
    lib.vfm_timing_enable(1)
    sync_events.clear()
    e0.record()
    for _ in range(table_steps):
        step(z, ws)
        flush.zero_()
    e1.record()
    barrier()
    ms_table = e0.elapsed_time(e1)
    lib.vfm_timing_enable(0)
    stats = timing_report()
    if exchange is not None:
        xstats = dict(exchange.stats)
        exchange.remove()
    else:
        xstats = None
    if mode == 'train':
        net.zero_grad(set_to_none=True)
        del opt
    total = batch * world * steps
    res_d = {'value': total / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms / steps,
             'e2e': {'value': total / (ms_e2e * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': (z_h.numel() + ws_h.numel()) * 4, 'd2h_bytes_per_step': result_bytes},
             'gpu_launches': int(launches), 'clocks': clocks, 'stats': stats, 'ms_table': ms_table, 'table_steps': table_steps,
             'timed_region': ('one CUDA graph replay per step (SynthesisNetwork.decode_graph; bit-identical to the eager step, checked); ' if graph_used else 'eager step; ')
                             + 'per-launch timing events off; kernel table / roofline from a separate eager pass with events on'}
    if mode == 'train':
        res_d['sync'] = {'sync_ms': sync_ms_max, 'what': 'exposed gradient-exchange time per step (CUDA events around the wait + finalize after backward; max over ranks)',
                         'mode': sync_mode, 'gradient_bytes': (xstats or {}).get('bytes', sum(p.numel() for p in net.parameters()) * 4),
                         'buckets': (xstats or {}).get('buckets'), 'buckets_sent_during_backward': (xstats or {}).get('launched_in_backward')}
    return res_d


def run_ours(args):
    import ctypes as C
    import torch.distributed as dist
    from vfm_vae_b200 import _lib
    from vfm_vae_b200.decoder import SynthesisNetwork
    from vfm_vae_b200 import sync

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # the reference turns cuDNN autotuning on (training/training_loop.py:490,503); it only affects the stock-PyTorch glue layers
    # around the hot-path ops (z-convs, attention 1x1 convs, pixel-shuffle upsampler), never the kernels measured here
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cx = Ctx()
    cx.dev, cx.world, cx.rank, cx.local = dev, world, rank, local
    cx.lib = _lib.load()
    cx.kw = lambda res: decoder_kwargs(args.variant, res, args.fp16_res)
    cx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def graph_launches(net, z, ws):
        with torch.no_grad():
            l0 = _lib.launch_count()
            net(z, ws)
            return _lib.launch_count() - l0
    cx.graph_launches = graph_launches

    torch.manual_seed(0)
    net = SynthesisNetwork(**cx.kw(args.res)).to(dev)
    sync.broadcast_module(net)

    cpu_base = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, parity = cpu_baseline_and_parity(args, net, dev)
        if args.no_parity:
            parity = None

    table_steps = min(args.steps, 5)
    main = measure(cx, net, args.mode, args.res, args.batch, args.steps, args.warmup, table_steps, args.cuda_graph, args.sync)
    blocks = {}
    if args.mode == 'train' and not args.no_secondary:
        sec_steps = max(3, min(args.steps, 10))
        blocks['decode'] = (measure(cx, net, 'decode', args.res, args.batch, sec_steps, 3, min(sec_steps, 5), args.cuda_graph, args.sync), args.res, args.batch)
        net.__dict__.pop('_decode_graphs', None)
        if args.res == 256 and args.variant == 'legacy':
            del net
            torch.cuda.empty_cache()
            torch.manual_seed(0)
            net512 = SynthesisNetwork(**cx.kw(512)).to(dev)
            sync.broadcast_module(net512)
            blocks['decode512'] = (measure(cx, net512, 'decode', 512, 32, sec_steps, 3, min(sec_steps, 5), args.cuda_graph, args.sync), 512, 32)
            del net512

    if rank == 0:
        def finish(m, mode, res, batch):
            traffic_applies = args.variant == 'legacy' and args.fp16_res == 3 and ((mode == 'train' and batch == 64 and res == 256) or (mode == 'decode' and batch == 64 and res == 256))
            roofline, kernels = kernel_table(m.pop('stats'), m['ms_table'], mode, traffic_applies)
            m['roofline'], m['kernels'] = roofline, kernels
            m['workload'] = workload_name(args.variant, mode, res, batch, args.fp16_res)
            return m
        main = finish(main, args.mode, args.res, args.batch)
        line = {
            'metric': f'images/sec ({args.mode})', 'value': main['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16' if args.fp16_res > 0 else 'f32', 'data': 'synthetic',
            'config': {'workload': main['workload'], 'global_batch': args.batch * world,
                       'parallelism': f'batch-sharded x{world}' + (' (replicas, no collective)' if args.mode == 'decode' else
                                                                   ' + gradient mean over ranks (NCCL all_reduce of the full fp32 gradient, bucketed' +
                                                                   (', overlapped with backward)' if args.sync == 'overlap' else ', post hoc)')),
                       'l2': 'explicit 256 MiB flush write between timed iterations; per-step activations (GBs) exceed the 126 MB L2 anyway',
                       'decoder_variant': 'D-legacy (use_convnext=False)' if args.variant == 'legacy' else 'D-convnext (use_convnext=True)',
                       'cudnn_benchmark': bool(torch.backends.cudnn.benchmark), 'timed_region': main['timed_region']},
            'e2e': main['e2e'], 'gpu_launches': main['gpu_launches'], 'clocks': main['clocks'], 'roofline': main['roofline'],
        }
        if 'sync' in main:
            line['sync'] = main['sync']
            line['sync_ms'] = main['sync']['sync_ms']
        if cpu_base is not None:
            line['cpu_baseline'] = cpu_base[args.mode]
        if parity is not None:
            line['parity'] = parity
        for name, (m, res, batch) in blocks.items():
            m = finish(m, 'decode', res, batch)
            blk = {k: m[k] for k in ('workload', 'value', 'unit', 'ms_per_step', 'e2e', 'gpu_launches', 'roofline', 'timed_region', 'clocks')}
            blk['n_gpus'], blk['global_batch'] = world, batch * world
            if cpu_base is not None and name == 'decode':
                blk['cpu_baseline'] = cpu_base['decode']
            blk['kernels'] = m['kernels'][:12]
            line[name] = blk
        line['kernels'] = main['kernels']
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints to fd 1 while the bench runs (NCCL's version
    banner, the reference's constructor prints) was diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (line + '\n').encode())


_REAL_STDOUT = 1

if __name__ == '__main__':
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == 'reference':
        run_reference_arm(a)
    else:
        run_ours(a)
