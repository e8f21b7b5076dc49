#!/bin/bash
# scratch driver for one gpurun call (kernel bring-up)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k latent > gpurun_out/latent_test.log 2>&1
echo "rc $?" >> gpurun_out/latent_test.log
timeout 120 python tools/microbench.py latent > gpurun_out/latent_bench.log 2>&1
echo "rc $?" >> gpurun_out/latent_bench.log
tail -5 gpurun_out/latent_test.log
cat gpurun_out/latent_bench.log
cd whisper-flamingo_b200 && touch csrc/latent.cu && make EXTRA=-DLA_TIMING > /dev/null 2>&1 && cd .. && timeout 60 python tools/latent_once.py 2>&1 | tail -5 | tee gpurun_out/latent_timing.log
