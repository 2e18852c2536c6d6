#!/bin/bash
# GPU box: ncu --set full of the K2 kernel on the ppa-like micro list (tools/k2_bench.py), source counters included
TAG=${1:-r2s}
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linkpred_tc3_kernel --launch-skip 2 --launch-count 1 \
  -o /tmp/ncu/${TAG}_k2 -f python tools/k2_bench.py 24 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i /tmp/ncu/${TAG}_k2.ncu-rep --page raw --csv > gpurun_out/${TAG}_k2_raw.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}_k2.ncu-rep --page source --csv > gpurun_out/${TAG}_k2_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
