#!/bin/bash
# GPU box: ncu --set full of the final fused one-pass kernel only (CSV export).  usage: tools/gpu_round_q.sh TAG
TAG=${1:-rq}
mkdir -p gpurun_out /tmp/ncu
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"twohop_score_kernel" --launch-skip 3 --launch-count 1 \
  -o /tmp/ncu/${TAG} -f python bench.py --steps 1 --warmup 3 --slabs 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/ncu/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -3
