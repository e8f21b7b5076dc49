#!/bin/bash
# scratch driver for one gpurun call (kernel bring-up)
mkdir -p gpurun_out
: > gpurun_out/latent_exp.log
for e in 1 2 3 4; do
  (cd whisper-flamingo_b200 && touch csrc/latent.cu && make EXTRA=-DLA_EXP=$e > /dev/null 2>&1)
  echo "LA_EXP=$e" >> gpurun_out/latent_exp.log
  LATENT_SHAPES=1 timeout 120 python tools/microbench.py latent 2>&1 | tail -1 >> gpurun_out/latent_exp.log
done
cat gpurun_out/latent_exp.log
