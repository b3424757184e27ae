#!/usr/bin/env python
"""Markdown rows for DESIGN.md from a `VFM_TIMING_DETAIL=1 python bench.py ...` JSON line: per layer shape ms / TFLOP/s / GB/s and the
fraction of the measured peak (MEASURED_PEAKS.json).    python tools/design_tables.py gpurun_out/<detail>.json [min_ms_per_step]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
tc, hbm = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('hbm_gbs', 6650.0)
j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
steps = j['steps']
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
print(f"step: {j['value']:.1f} {j['unit']}, {j['ms_per_step']:.2f} ms; kernels listed from {floor} ms per step\n")
print('| kernel : layer | launches / step | ms / launch | ms / step | achieved | of measured peak |\n|---|---|---|---|---|---|')
for k in j['kernels']:
    per_step = k['total_ms'] / steps
    if per_step < floor:
        continue
    if 'tflops' in k and k['tflops'] > 20:
        ach, frac = f"{k['tflops']:.0f} TFLOP/s", k['tflops'] / tc
    elif 'gbs' in k:
        ach, frac = f"{k['gbs']:.0f} GB/s", k['gbs'] / hbm
    else:
        ach, frac = '', 0
    print(f"| `{k['name']}` | {k['launches'] / steps:g} | {k['avg_ms']:.3f} | {per_step:.2f} | {ach} | {frac:.2f} |")
