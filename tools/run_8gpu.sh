#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "bench8 rc $?"; tail -2 gpurun_out/bench_8gpu.err; python - <<'PY'
import json
for l in open('gpurun_out/bench_8gpu.json'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print({k:j[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks') if k in j})
PY
