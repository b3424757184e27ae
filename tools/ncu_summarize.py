#!/usr/bin/env python
"""Turns ncu output into the markdown / json summaries committed under profiles/.

    python tools/ncu_summarize.py launches <launches.csv> <out.md> [title]
        per-kernel launch counts, total time and share of a `--metrics gpu__time_duration.sum --csv` launch list
    python tools/ncu_summarize.py merge <out.json> <traffic1.json> <traffic2.json> ...
        launch-weighted mean DRAM bytes per kernel name over several `full` jsons (the per-layer probes of one step)
    python tools/ncu_summarize.py full <report.ncu-rep> <out.md> [out.json]
        one row per captured launch of a `--set full` report: duration, DRAM bytes, tensor-pipe / DRAM / issue utilisation;
        out.json gets the per-kernel-name average DRAM traffic per launch (bench.py reads it for roofline.traffic)
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|vfm::|at::native::', '', name)
    name = re.sub(r'\(.*', '', name)
    return name[:90]


def launches(path, out, title):
    rows = list(csv.reader(l for l in open(path, errors='replace') if l.startswith('"')))
    hdr = rows[0]
    iN, iM, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rows[1:]:
        if len(r) <= iV or r[iM] != 'gpu__time_duration.sum':
            continue
        v = float(r[iV].replace(',', ''))
        ms = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(r[iU], 1e-6) * v
        k = short(r[iN])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
        n += 1
    with open(out, 'w') as f:
        f.write(f'# {title}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n')
        f.write(f'# total profiled GPU time: {total:.1f} ms over {n} launches\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n')
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
            f.write(f'| `{k}` | {c} | {ms:.2f} | {100 * ms / total:.1f}% |\n')


def merge(out_json, paths):
    """paths: `file.json` or `file.json:weight` -- weight = how many layers of the benchmarked step the probe stands for"""
    acc = {}
    for spec in paths:
        path, _, wt = spec.partition(':')
        wt = float(wt) if wt else 1.0
        for k, v in json.load(open(path)).items():
            k = re.sub(r'void |unnamed>::|modconv::|vfm::', '', k)
            a = acc.setdefault(k, {'launches': 0.0, 'bytes': 0.0, 'probes': []})
            a['launches'] += wt * v['launches']
            a['bytes'] += wt * v['launches'] * v['dram_bytes_per_launch']
            a['probes'].append(path.split('/')[-1].replace('r02_traffic_', '').replace('.json', '') + (f' x{wt:g}' if wt != 1 else ''))
    res = {k: {'launches': v['launches'], 'dram_bytes_per_launch': v['bytes'] / v['launches'], 'probes': v['probes']} for k, v in acc.items()}
    json.dump(res, open(out_json, 'w'), indent=1)


def full(rep, out, out_json):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = OrderedDict([
        ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'), ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid')])
    idx = {k: hdr.index(k) for k in cols if k in hdr}
    iN = hdr.index('Kernel Name')

    def to_bytes(v, u):
        return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)

    traffic = {}
    with open(out, 'w') as f:
        f.write(f'# ncu --set full, one row per captured launch ({rep.split("/")[-1]})\n\n| kernel | ' + ' | '.join(cols[k] for k in idx) + ' |\n|---|' + '---|' * len(idx) + '\n')
        for r in rows[2:]:
            if len(r) <= iN:
                continue
            cells = []
            for k, i in idx.items():
                v, u = r[i], units[i]
                try:
                    cells.append(f'{float(v):.4g} {u}'.strip())
                except ValueError:
                    cells.append(v)
            f.write(f'| `{short(r[iN])}` | ' + ' | '.join(cells) + ' |\n')
            if 'dram__bytes_read.sum' in idx:
                b = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
                t = traffic.setdefault(short(r[iN]), [0, 0.0])
                t[0] += 1
                t[1] += b
    if out_json:
        json.dump({k: {'launches': c, 'dram_bytes_per_launch': b / c} for k, (c, b) in traffic.items()}, open(out_json, 'w'), indent=1)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'merge':
    merge(sys.argv[2], sys.argv[3:])
    sys.exit(0)
if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else 'ncu launch list')
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
