#!/bin/bash
# GPU box: ncu --set full + source counters of one kernel (regex $2) in a short bench run
TAG=${1:-r3l}; K=${2:-twohop_score_kernel}
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip ${SKIP:-4} --launch-count 1 \
  -o /tmp/ncu/${TAG} -f python bench.py --steps 1 --warmup 1 --owners-frac 0.05 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i /tmp/ncu/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
