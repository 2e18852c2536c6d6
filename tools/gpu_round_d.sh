#!/bin/bash
# GPU box: ncu --set full of ONE kernel (regex $2) of a bench step, per env variant; CSV exports only.
# usage: tools/gpu_round_d.sh TAG REGEX SKIP "ENV1" "ENV2" ...
TAG=$1; RX=$2; SKIP=$3; shift 3
mkdir -p gpurun_out /tmp/ncu
i=0
for envs in "$@"; do
  i=$((i+1)); if [ "$envs" = "-" ]; then envs=""; fi
  env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" --launch-skip $SKIP --launch-count 1 \
    -o /tmp/ncu/${TAG}_v${i} -f python bench.py --steps 1 --warmup 3 --slabs 1 --no-cpu-baseline > gpurun_out/${TAG}_v${i}.log 2>&1
  ncu -i /tmp/ncu/${TAG}_v${i}.ncu-rep --page raw --csv > gpurun_out/${TAG}_v${i}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/${TAG}_v${i}.ncu-rep --page source --csv > gpurun_out/${TAG}_v${i}_source.csv 2>/dev/null
done
du -sh gpurun_out; echo done > gpurun_out/${TAG}_done
