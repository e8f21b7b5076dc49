#!/bin/bash
LOG=gpurun_out/run5.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -5 >> $LOG
echo "=== bench large-v2 B=128" >> $LOG
timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 >> $LOG
echo "=== bench small B=16" >> $LOG
timeout 1200 python bench.py --workload small --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 >> $LOG
