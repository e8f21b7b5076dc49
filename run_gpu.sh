#!/bin/bash
LOG=gpurun_out/run46.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $LOG
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2>> $LOG
tail -1 gpurun_out/bench_final.json >> $LOG
timeout 600 python bench.py --single-step >> $LOG 2>&1 || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3300 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --single-step > gpurun_out/ncu_launch5.log 2>&1
echo "ncu launches rc $?" >> $LOG
prof() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_r01 -f python bench.py --single-step > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc $?" >> $LOG
}
prof gemm_tc2 gemm_tc2 6 4
timeout 600 python tools/microbench.py gemm attn2 skinny3 mel > gpurun_out/micro_final.log 2>&1
