#!/bin/bash
LOG=gpurun_out/run35.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 >> $LOG
timeout 600 python bench.py --single-step >> $LOG 2>&1 || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3300 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --single-step > gpurun_out/ncu_launch4.log 2>&1
echo "ncu launches rc $?" >> $LOG
prof() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_r01 -f python bench.py --single-step > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc $?" >> $LOG
}
prof attn_decode_hm attn_decode_hm 2 2
prof gemm_tc gemm_tc_kernel 6 2
prof gemm_skinny gemm_skinny 10 3
prof fa_tc fa_tc 1 1
prof logmel logmel_fft 0 1
