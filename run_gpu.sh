#!/bin/bash
LOG=gpurun_out/run38.log; : > $LOG
for v in "WF_NO_ENC_LN_FUSION=1" "WF_NO_ENC_LN_FUSION=0" "WF_GEMM_PAIR=0"; do
echo "== $v" >> $LOG
env $v WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/tmp.json 2>> $LOG
python - >> $LOG <<'P'
import json
d=json.loads(open('gpurun_out/tmp.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']), d.get('phases_ms'), round(d['e2e']['value']))
P
done
