#!/bin/bash
# scratch: A/B of compile-time variants of csrc/latent.cu (EXTRA flags), microbenchmark per variant
mkdir -p gpurun_out
: > gpurun_out/latent_exp.log
for e in "-DLA_STREAM_Q=1 -DLA_NB_FORCE=4" "-DLA_STREAM_Q=1 -DLA_NB_FORCE=3 -DLA_AHEAD_PCT=75" "-DLA_STREAM_Q=1 -DLA_NB_FORCE=3 -DLA_AHEAD_PCT=25"; do
  (cd whisper-flamingo_b200 && touch csrc/latent.cu && make EXTRA="$e" > /dev/null 2>&1)
  echo "variant [$e]" >> gpurun_out/latent_exp.log
  LATENT_SHAPES=2 timeout 120 python tools/microbench.py latent 2>&1 | tail -2 >> gpurun_out/latent_exp.log
done
cat gpurun_out/latent_exp.log
