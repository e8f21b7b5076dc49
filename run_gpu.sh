#!/bin/bash
mkdir -p gpurun_out
cd whisper-flamingo_b200 && touch csrc/latent.cu && make EXTRA=-DLA_TIMING > /dev/null 2>&1 && cd ..
B=16 H=12 timeout 60 python tools/latent_once.py 2>&1 | tail -19 > gpurun_out/latent_timing.log
B=128 H=20 timeout 60 python tools/latent_once.py 2>&1 | tail -19 >> gpurun_out/latent_timing.log
cat gpurun_out/latent_timing.log
