#!/bin/bash
# scratch: A/B of compile-time variants of csrc/latent.cu (EXTRA flags), parity tests + microbenchmark per variant
mkdir -p gpurun_out
: > gpurun_out/latent_exp.log
for e in "-DLA_KEYS_PER_TILE=64" "-DLA_KEYS_PER_TILE=64 -DLA_AHEAD_PCT=100" "-DLA_KEYS_PER_TILE=64 -DLA_AHEAD_PCT=200"; do
  (cd whisper-flamingo_b200 && touch csrc/latent.cu && make EXTRA="$e" > /dev/null 2>&1)
  echo "variant [$e]" >> gpurun_out/latent_exp.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k latent 2>&1 | tail -2 >> gpurun_out/latent_exp.log
  timeout 120 python tools/microbench.py latent 2>&1 | tail -4 >> gpurun_out/latent_exp.log
done
cat gpurun_out/latent_exp.log
