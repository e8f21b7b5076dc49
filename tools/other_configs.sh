#!/bin/bash
# BASELINE configs 2 and 3 (1 GPU): small AV B=16 greedy, medium AV B=64 beam 5
mkdir -p gpurun_out
timeout 200 python bench.py --workload small --no-profile --no-cpu-baseline > gpurun_out/bench_small.json 2>/dev/null
timeout 300 python bench.py --workload medium --beam 5 --no-profile --no-cpu-baseline > gpurun_out/bench_medium_beam5.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/bench_small.json','gpurun_out/bench_medium_beam5.json'):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            j=json.loads(l); print(f, {k:j[k] for k in ('value','ms_per_step') if k in j}, j['config'].get('workload'))
PY
