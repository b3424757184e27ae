#!/usr/bin/env python
"""Prints the per-kernel table of a bench.py JSON line (stdin or file): name, launches, total ms, avg ms, TFLOP/s, GB/s."""
import json
import sys

line = [l for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin) if l.startswith('{')][-1]
j = json.loads(line)
print(f"{j['metric']}: {j['value']:.1f} {j['unit']}  {j['ms_per_step']:.2f} ms/step  e2e {j.get('e2e', {}).get('value', 0):.1f}  launches {j.get('gpu_launches')}")
for k in j.get('kernels', []):
    print(f"{k['name'][:48]:48s} n={k['launches']:5d} total={k['total_ms']:9.3f} avg={k['avg_ms']:8.4f} "
          f"{('%7.1f TF' % k['tflops']) if 'tflops' in k else '          '} {('%7.1f GB/s' % k['gbs']) if 'gbs' in k else ''}")
