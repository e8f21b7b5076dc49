#!/bin/bash
LOG=gpurun_out/run50.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> $LOG
