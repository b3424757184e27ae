#!/usr/bin/env python
"""Transposes `ncu -i REP --page raw --csv` (stdin) and prints, per captured launch, the warp-stall breakdown and the
issue / pipe utilisation metrics (the numbers quoted in DESIGN.md for the epilogue-bound 1x1 convs)."""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]
pat = re.compile(r'issue_stalled.*_per_warp_active\.pct|gpu__time_duration\.sum|smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|'
                 r'sm__inst_executed_pipe_(alu|fma|fmaheavy|xu|uniform|tmem|lsu)[a-z_]*\.sum$|sm__pipe_tensor.*pct_of_peak_sustained_active|'
                 r'dram__bytes_(read|write)\.sum$|sm__warps_active\.avg\.pct|launch__registers_per_thread|smsp__cycles_active\.avg$|'
                 r'lts__t_sectors_srcunit_tex_op_read\.sum$|sm__cycles_elapsed\.max$')
for r in data:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '')[:90], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    out = []
    for k, v in d.items():
        if pat.search(k):
            try:
                out.append((k, float(v.replace(',', ''))))
            except ValueError:
                pass
    for k, v in sorted(out, key=lambda kv: (('stalled' in kv[0]), -kv[1] if 'stalled' in kv[0] else 0, kv[0])):
        if 'stalled' in k and v < 1.0:
            continue
        print(f'   {k:95s} {v:16.2f}')
