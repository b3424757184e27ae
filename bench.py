#!/usr/bin/env python
"""bench.py -- images/sec of the VFM-VAE f16d32 pixel-decoder hot path on B200 (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode decode|train] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input:
  decode: D-legacy SynthesisNetwork forward (z [B,512,16,16], ws [B,36,512] -> img [B,3,256,256] + 5 multi-scale images)
  train : forward + backward through the same decoder + reference-style gradient all-mean (vfm_vae_b200/sync.py) + Adam step
Workload = BASELINE.json configs[1] restricted to the hot path: f16d32, 256x256, batch 64 per GPU, num_fp16_res=3 (fp16
blocks 3-5, fp32 blocks 0-2), random-init weights, synthetic latents.  The frozen SigLIP2 encoder of configs[1] is outside
the hot-path scope and is not part of the timed region.  D-legacy = use_convnext=False, the variant
north_star describes (the shipped YAMLs run the ConvNeXt variant, which calls none of these ops: SURVEY.md 0.2).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` = same metric
through the public Python API with pinned HOST buffers, H2D of the inputs and D2H of the images inside the timed region.
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
    ap.add_argument('--mode', choices=['decode', 'train'], default='decode')
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--res', type=int, default=256, choices=[256, 512])
    ap.add_argument('--fp16-res', type=int, default=3, help='num_fp16_res (3 = training configs, 0 = the inference tools)')
    ap.add_argument('--cpu-batch', type=int, default=4, help='sample size of the CPU legs')
    ap.add_argument('--variant', choices=['legacy', 'convnext'], default='legacy',
                    help="decoder variant: 'legacy' = use_convnext=False (the path north_star names, default), 'convnext' = the shipped YAMLs' layers")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cuda-graph', choices=['auto', 'on', 'off'], default='off',
                    help='decode only: replay the whole step as ONE captured CUDA graph in the timed regions (the step is ~900 launches, '
                         'partly launch-bound in the 8x8..32x32 blocks); the per-kernel table then comes from a separate eager pass')
    ap.add_argument('--no-cudnn-benchmark', action='store_true', help='leave torch.backends.cudnn.benchmark off for the glue layers')
    return ap.parse_args()


def decoder_kwargs(args):
    from vfm_vae_b200.decoder import F16D32_LEGACY_KWARGS, F16D32_CONVNEXT_KWARGS
    kw = dict(F16D32_CONVNEXT_KWARGS if args.variant == 'convnext' else F16D32_LEGACY_KWARGS)
    kw['img_resolution'] = args.res
    kw['z_resolution'] = args.res // 16
    kw['num_fp16_res'] = args.fp16_res
    return kw


def make_inputs(kw, batch, num_ws, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(batch, kw['z_dim'], kw['z_resolution'], kw['z_resolution'], generator=g)
    ws = torch.randn(batch, num_ws, kw['w_dim'], generator=g)
    return z, ws


def workload_name(args):
    variant = 'D-legacy pixel decoder (SynthesisNetwork use_convnext=False)' if args.variant == 'legacy' else 'D-convnext pixel decoder (SynthesisNetwork use_convnext=True)'
    return (f'f16d32 {variant} {args.mode}, {args.res}x{args.res}, '
            f'batch {args.batch}/GPU, num_fp16_res={args.fp16_res}, random-init weights, synthetic latents '
            f'(BASELINE configs[1] restricted to the hot path; SigLIP2 encoder out of scope)')


# ----------------------------------------------------------------------------------------------------------- CPU legs

def oracle_ops():
    """CPU oracle ops injected into the decoder mirror -- the cpu_baseline / reference arm, the only place bench.py may
    execute oracle/."""
    from types import SimpleNamespace
    from oracle import ref_ops as O

    def modconv(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True, flip_weight=True, fused_modconv=True):
        return O.modulated_conv2d(x, weight, styles, noise=noise, up=up, down=down, padding=padding, resample_filter=resample_filter,
                                  demodulate=demodulate, flip_weight=flip_weight)
    return SimpleNamespace(bias_act=O.bias_act, def_gain=lambda a: O.ACTIVATIONS[a][1], setup_filter=O.setup_filter,
                           upsample2d=O.upsample2d, modulated_conv2d=modconv, modulated_pointwise_conv2d=O.modulated_pointwise_conv2d)


def cpu_step_fn(args, batch):
    from vfm_vae_b200.decoder import SynthesisNetwork
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = decoder_kwargs(args)
    torch.manual_seed(0)
    net = SynthesisNetwork(ops=oracle_ops(), **kw)
    z, ws = make_inputs(kw, batch, net.num_ws, seed=1)
    if args.mode == 'decode':
        net.eval().requires_grad_(False)

        def step():
            with torch.no_grad():
                return net(z, ws)[0]
    else:
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)

        def step():
            img, multi = net(z, ws)
            loss = img.square().mean() + sum(m.square().mean() for m in multi)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return img
    return step, cores


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = args.cpu_batch if (args.steps + args.warmup) <= 16 else max(1, args.cpu_batch // 2)
    step, cores = cpu_step_fn(args, batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    sample = (f'oracle port (torch CPU fp32, oracle/ref_ops.py) of the same decoder, {batch} images per step, '
              f'{args.steps} timed steps after {args.warmup} warm-ups')
    _emit(json.dumps({
        'impl': 'reference', 'metric': f'images/sec ({args.mode})', 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'cpu_sample_batch': batch},
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def cpu_baseline(args):
    batch = args.cpu_batch
    step, cores = cpu_step_fn(args, batch)
    step() if args.mode == 'decode' else None     # one warm-up for the cheap mode only (bounded CPU time)
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return {'value': batch / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
            'sample': f'oracle port (torch CPU fp32) of the same decoder {args.mode} step, one pass over {batch} images ({dt:.1f} s)'}


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

    class Stat(C.Structure):
        _fields_ = [('name', C.c_char * 64), ('launches', C.c_int64), ('total_ms', C.c_double), ('flops', C.c_double), ('bytes', C.c_double)]
    lib = _lib.load()
    lib.vfm_timing_report.restype = C.c_int
    lib.vfm_timing_report.argtypes = [C.POINTER(Stat), C.c_int]
    buf = (Stat * 256)()
    n = min(lib.vfm_timing_report(buf, 256), 256)
    return [dict(name=buf[i].name.decode(), launches=int(buf[i].launches), total_ms=buf[i].total_ms, flops=buf[i].flops, bytes=buf[i].bytes)
            for i in range(n)]


def run_ours(args):
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
    # same numerics switches as the reference: its training loop turns TF32 off for the fp32 glue layers
    # (training/training_loop.py:504-505); its decode/reconstruct tools leave PyTorch's defaults untouched
    if args.mode == 'train':
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    # the reference turns cuDNN autotuning on (training/training_loop.py:490,503); it only affects the stock-PyTorch glue layers
    # around the hot-path ops (z-convs, attention 1x1 convs, pixel-shuffle upsampler), never the kernels measured here
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    lib.vfm_timing_enable.argtypes = [__import__('ctypes').c_int]
    lib.vfm_timing_enable.restype = None

    kw = decoder_kwargs(args)
    torch.manual_seed(0)
    net = SynthesisNetwork(**kw).to(dev)
    sync.broadcast_module(net)
    z_h, ws_h = make_inputs(kw, args.batch, net.num_ws, seed=1 + rank)
    z_h, ws_h = z_h.pin_memory(), ws_h.pin_memory()
    z, ws = z_h.to(dev), ws_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    if args.mode == 'decode':
        net.eval().requires_grad_(False)

        def step(z_, ws_):
            with torch.no_grad():
                return net(z_, ws_)[0]
    else:
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4)

        def step(z_, ws_):
            img, multi = net(z_, ws_)
            loss = img.square().mean() + sum(m.square().mean() for m in multi)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            sync.sync_grads(params)
            opt.step()
            return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(z, ws)
        flush.zero_()
    barrier()

    # ---- optional: the decode step as one CUDA graph (captured after the eager warm-up, so cuDNN autotuning is done) ----
    use_graph = args.mode == 'decode' and args.cuda_graph in ('on', 'auto') and not os.environ.get('VFM_CUDA_PROFILER_RANGE')
    graph = out_s = None
    graph_launches = 0
    if use_graph:
        try:
            z_s, ws_s = z.clone(), ws.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(z_s, ws_s)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                out_s = step(z_s, ws_s)
            graph_launches = _lib.launch_count() - l0
            graph.replay()
            torch.cuda.synchronize()
            ref_img = step(z, ws)
            if not torch.equal(out_s, ref_img):                 # same kernels, same inputs: the replay must reproduce the eager result
                raise RuntimeError('graph replay differs from the eager step')
        except Exception as e:                                    # noqa: BLE001
            if args.cuda_graph == 'on':
                raise
            print(f'[bench] CUDA graph capture unavailable ({e}); timing the eager step', file=sys.stderr)
            graph = None
    barrier()

    # ---- device-timed region: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = bool(os.environ.get('VFM_CUDA_PROFILER_RANGE'))     # ncu --profile-from-start off: capture the timed region only
    if graph is not None:
        barrier()
        e0.record()
        for _ in range(args.steps):
            graph.replay()
            flush.zero_()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = graph_launches * args.steps
        # per-kernel durations (roofline, kernel table): the same kernels in an eager pass with per-launch events, outside `value`
        lib.vfm_timing_enable(1)
        for _ in range(args.steps):
            step(z, ws)
            flush.zero_()
        barrier()
        lib.vfm_timing_enable(0)
        stats = timing_report()
    else:
        lib.vfm_timing_enable(1)
        launches0 = _lib.launch_count()
        barrier()
        if prof_range:
            torch.cuda.profiler.start()
        e0.record()
        for _ in range(args.steps):
            step(z, ws)
            flush.zero_()          # L2 flush between iterations (256 MiB write, ~0.05 ms)
        e1.record()
        barrier()
        if prof_range:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        lib.vfm_timing_enable(0)
        stats = timing_report()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()

    # ---- end-to-end: pinned host inputs -> H2D -> decoder -> D2H of the images, every step ----
    out_h = torch.empty(args.batch, 3, args.res, args.res, dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            z_s.copy_(z_h, non_blocking=True)
            ws_s.copy_(ws_h, non_blocking=True)
            graph.replay()
            out_h.copy_(out_s, non_blocking=True)
        else:
            zd = z_h.to(dev, non_blocking=True)
            wd = ws_h.to(dev, non_blocking=True)
            img = step(zd, wd)
            out_h.copy_(img.detach(), non_blocking=True)
        flush.zero_()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()

    if rank == 0:
        total_imgs = args.batch * world * args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks.get('hbm_gbs'), 'measured') if peaks.get('hbm_gbs') else (6650.0, 'fallback')
        tc_peak, tc_src = (peaks.get('bf16_tflops_sustained'), 'measured sustained') if peaks.get('bf16_tflops_sustained') else (1400.0, 'fallback')
        stats.sort(key=lambda s: -s['total_ms'])
        roofline = None
        kernels = []
        for s in stats:
            per = s['total_ms'] / max(s['launches'], 1)
            entry = {'name': s['name'], 'launches': s['launches'], 'total_ms': round(s['total_ms'], 3), 'avg_ms': round(per, 4)}
            if s['flops'] > 0:
                entry['tflops'] = round(s['flops'] / (s['total_ms'] * 1e-3) / 1e12, 3)
            if s['bytes'] > 0:
                entry['gbs'] = round(s['bytes'] / (s['total_ms'] * 1e-3) / 1e9, 1)
            kernels.append(entry)
        # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
        # (profiles/r01c_traffic.json, written by tools/ncu_summarize.py); None when no capture covers the kernel
        def ncu_traffic(name):
            key = {'modconv_tc_fwd': 'conv_tc_kernel<__half, 0, 0', 'modconv_tc_fwd_split': 'conv_tc_kernel<float, 0, 1',
                   'modconv_nhwc_prepass': 'nhwc_prepass_kernel<__half', 'upfirdn2d_blur': 'upfirdn2d_blur<__half'}.get(name.split(':')[0])
            try:
                tr = json.load(open(os.path.join(REPO, 'profiles', 'r01c_traffic.json')))
            except Exception:
                return None
            sel = [v for k, v in tr.items() if key and key in k]
            n = sum(v['launches'] for v in sel)
            return sum(v['launches'] * v['dram_bytes_per_launch'] for v in sel) / n if n else None
        # the committed ncu capture is of the default workload (legacy decoder, decode, 256x256, batch 64): other runs report null
        traffic_applies = args.mode == 'decode' and args.variant == 'legacy' and args.res == 256 and args.batch == 64 and args.fp16_res == 3
        if stats:
            top = stats[0]
            if top['flops'] > 0:
                ach = top['flops'] / (top['total_ms'] * 1e-3) / 1e12
                roofline = {'kernel': top['name'], 'bound': 'tensor', 'achieved': ach, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': ach / tc_peak,
                            'traffic': ncu_traffic(top['name']) if traffic_applies else None,
                            'traffic_unit': 'bytes/launch (dram read+write, ncu --set full, profiles/r01c_ncu_full_decode.md)',
                            'algorithmic_per_launch': top['flops'] / max(top['launches'], 1), 'peak_source': tc_src, 'share_of_step': top['total_ms'] / ms}
            else:
                ach = top['bytes'] / (top['total_ms'] * 1e-3) / 1e9
                roofline = {'kernel': top['name'], 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                            'traffic': ncu_traffic(top['name']) if traffic_applies else None,
                            'traffic_unit': 'bytes/launch (dram read+write, ncu --set full, profiles/r01c_ncu_full_decode.md)',
                            'algorithmic_per_launch': top['bytes'] / max(top['launches'], 1), 'peak_source': hbm_src, 'share_of_step': top['total_ms'] / ms}
        line = {
            'metric': f'images/sec ({args.mode})', 'value': total_imgs / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16' if args.fp16_res > 0 else 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'global_batch': args.batch * world, 'parallelism': f'batch-sharded x{world}' + (' (replicas, no collective)' if args.mode == 'decode' else ' + gradient all-mean (NCCL)'),
                       'l2': 'explicit 256 MiB flush write between timed iterations; per-step activations (GBs) exceed the 126 MB L2 anyway',
                       'decoder_variant': 'D-legacy (use_convnext=False)' if args.variant == 'legacy' else 'D-convnext (use_convnext=True)', 'cudnn_benchmark': bool(torch.backends.cudnn.benchmark),
                       'timed_region': ('one CUDA graph replay per step (captured eager step, bit-identical output checked); kernel table / roofline from a '
                                        'separate eager pass with per-launch events') if graph is not None else 'eager step, per-launch events inside the timed region'},
            'e2e': {'value': total_imgs / (ms_e2e * 1e-3), 'unit': 'images/s',
                    'h2d_bytes_per_step': (z_h.numel() + ws_h.numel()) * 4, 'd2h_bytes_per_step': out_h.numel() * 4},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'kernels': kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(args)
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints to fd 1 while the bench runs (NCCL's version
    banner, for instance) was diverted to stderr by main()."""
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
