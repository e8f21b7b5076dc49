#!/bin/bash
LOG=gpurun_out/run23.log; : > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $LOG
WF_DECODE_SPLIT=1 timeout 900 python tools/split_probe.py large-v2 1 2>&1 | grep -v Warning >> $LOG
WF_NO_LN_FUSION=1 WF_DECODE_SPLIT=1 timeout 900 python tools/split_probe.py large-v2 1 2>&1 | grep -v Warning >> $LOG
WF_SKINNY=0 WF_NO_LN_FUSION=1 WF_DECODE_SPLIT=1 timeout 900 python tools/split_probe.py large-v2 1 2>&1 | grep -v Warning >> $LOG
