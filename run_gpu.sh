#!/bin/bash
# One gpurun call that regenerates the round's evidence (what profiles/r02_* were made from):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash run_gpu.sh [tag]'
# full GPU suite, bench, ncu launch list, ncu --set full of the hot kernels, isolated microbenchmarks.
TAG=${1:-r02}
mkdir -p gpurun_out
LOG=gpurun_out/run_${TAG}.log; : > $LOG
(cd whisper-flamingo_b200 && make > /dev/null 2>&1)
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $LOG
WF_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2>> $LOG
tail -1 gpurun_out/bench_${TAG}.json | cut -c1-400 >> $LOG
timeout 600 python bench.py --single-step >> $LOG 2>&1 || { cat $LOG; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --single-step > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc $?" >> $LOG
python tools/ncu_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/${TAG}_launches_large-v2_B128.txt 2>> $LOG
prof() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_${TAG} -f python bench.py --single-step > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc $?" >> $LOG
  python tools/ncu_summary.py gpurun_out/prof_$1_${TAG}.ncu-rep > gpurun_out/${TAG}_ncu_full_$1.txt 2>> $LOG
}
prof latent_pair latent_pair 4 4
prof latent_value latent_value 4 2
prof latent_query latent_query 4 2
prof gemm_skinny gemm_skinny 40 4
prof fa_tc fa_tc_kernel 2 1
prof gemm_tc2 gemm_tc2 6 4
prof logmel logmel 0 2
timeout 600 python tools/mel_sweep.py > gpurun_out/${TAG}_mel_sweep.txt 2>> $LOG
timeout 600 python tools/microbench.py gemm attn2 skinny3 mel fa latent > gpurun_out/micro_${TAG}.log 2>&1
echo "micro rc $?" >> $LOG
cat $LOG
