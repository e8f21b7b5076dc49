#!/bin/bash
LOG=gpurun_out/run15.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $LOG
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_large2.json 2>> $LOG
python - >> $LOG <<'P'
import json
d=json.loads(open('gpurun_out/bench_large2.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']), d['clocks'])
print(d.get('roofline'))
for k in d.get('kernels',[]): print(k)
print(d.get('cpu_baseline'))
P
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3200 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --single-step > gpurun_out/ncu_launch2.log 2>&1
echo "ncu rc $?" >> $LOG
