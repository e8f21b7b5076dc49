#!/bin/bash
mkdir -p gpurun_out
(cd whisper-flamingo_b200 && touch csrc/latent.cu && make > /dev/null 2>&1)
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -k "latent or large_v2" > gpurun_out/engine_latent_test.log 2>&1
echo "rc $?" >> gpurun_out/engine_latent_test.log
tail -15 gpurun_out/engine_latent_test.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_latent.json 2> gpurun_out/bench_latent.err
echo "rc $?"; tail -3 gpurun_out/bench_latent.err; cat gpurun_out/bench_latent.json
WF_LATENT=0 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_nolatent.json 2> gpurun_out/bench_nolatent.err
echo "rc $?"; cat gpurun_out/bench_nolatent.json
