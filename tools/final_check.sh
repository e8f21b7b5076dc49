#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_default.json'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print({k:j[k] for k in ('value','steps','warmup','ms_per_step','e2e','gpu_launches','clocks','roofline') if k in j})
PY
