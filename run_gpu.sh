#!/bin/bash
LOG=gpurun_out/run59.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> $LOG
for i in 1 2; do
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/tmp.json 2>> $LOG
python - >> $LOG <<'P'
import json
d=json.loads(open('gpurun_out/tmp.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']))
P
done
