#!/bin/bash
LOG=gpurun_out/run44.log; : > $LOG
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k linear 2>&1 | tail -3 >> $LOG
for a in "--workload medium --beam 5 --steps 2" ""; do
echo "== $a" >> $LOG
WF_TIMING=1 timeout 900 python bench.py $a --no-cpu-baseline --no-profile > gpurun_out/tmp.json 2>> $LOG
python - >> $LOG <<'P'
import json
try:
    d=json.loads(open('gpurun_out/tmp.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']), d['config']['workload'])
except Exception as e:
    print('FAILED', e)
P
done
