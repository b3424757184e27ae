#!/bin/bash
# Quick GPU check of a kernel change (one B200, a few minutes): GPU test suite, the default bench line, optional op-level table.
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
O=gpurun_out; TAG=${1:-chk}
stamp pytest; timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
stamp bench; timeout 500 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench.err
python tools/bench_kernels.py $O/${TAG}_bench.json | head -40
python - <<P
import json
d=[json.loads(l) for l in open('$O/${TAG}_bench.json') if l.startswith('{')][-1]
for b in ('decode','decode512'):
    if b in d: print(b, round(d[b]['value'],1), 'img/s', round(d[b]['ms_per_step'],2), 'ms  e2e', round(d[b]['e2e']['value'],1))
print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k in ('ok','max_rel','image_fp16_blocks','image_fp32')})
print('cpu', d.get('cpu_baseline',{}).get('value'), 'e2e', d['e2e']['value'])
P
stamp done
