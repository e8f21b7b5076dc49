#!/bin/bash
LOG=gpurun_out/run12.log; : > $LOG
export WF_TIMING=1
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -3 >> $LOG
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_large.json
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_large.json').read()); print(round(d['value']), round(d['ms_per_step']), d['phases_ms'], round(d['e2e']['value']), d['clocks']); print(d['roofline']); [print(k) for k in d['kernels']]" >> $LOG
