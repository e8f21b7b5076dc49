#!/bin/bash
LOG=gpurun_out/run27.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> $LOG
WF_ATTN_STAGES=3 timeout 300 python tools/microbench.py attn2 2>&1 | grep -v Warning >> $LOG
timeout 300 python tools/microbench.py attn2 2>&1 | grep -v Warning >> $LOG
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_large3.json 2>> $LOG
python - >> $LOG <<'P'
import json
d=json.loads(open('gpurun_out/bench_large3.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']), d['clocks'])
print(d.get('roofline'))
for k in d.get('kernels',[]): print(k)
P
