#!/bin/bash
# final check of a tree: full GPU suite, smoke, default bench, one ncu --set full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
WF_TIMING=1 timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_final.json'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print({k:j[k] for k in ('value','steps','warmup','ms_per_step','e2e','gpu_launches','phases_ms','roofline') if k in j})
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:latent_attn -s 2 -c 2 -o gpurun_out/prof_latent_attn_r01 -f python bench.py --single-step > gpurun_out/ncu_latent_attn.log 2>&1
echo "ncu rc $?"
python tools/ncu_summary.py gpurun_out/prof_latent_attn_r01.ncu-rep > gpurun_out/r01_ncu_full_latent_attn.txt
head -12 gpurun_out/r01_ncu_full_latent_attn.txt
timeout 120 python tools/microbench.py latent 2>&1 | tail -4
