#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --single-step > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc $?"
python tools/ncu_launches.py gpurun_out/launches_r01f.csv > gpurun_out/r01_launches_large-v2_B128.txt
tail -12 gpurun_out/r01_launches_large-v2_B128.txt
