#!/bin/bash
timeout 60 ./tools/probe/probe_m64 > gpurun_out/probe_m64.log 2>&1
echo "rc $?" >> gpurun_out/probe_m64.log
