#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -k "latent or large_v2" 2>&1 | tail -2
timeout 600 python bench.py --no-profile > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_pdl.json'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print({k:j[k] for k in ('value','ms_per_step','e2e') if k in j})
PY
