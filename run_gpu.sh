#!/bin/bash
LOG=gpurun_out/run34.log; : > $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k linear 2>&1 | tail -3 >> $LOG
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_large7.json 2>> $LOG
python - >> $LOG <<'P'
import json
for f in ('bench_large7',):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']))
P
