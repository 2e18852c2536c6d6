#!/bin/bash
# GPU box: final evidence — launch list + ncu --set full of the round's new kernels (CSV exports only), then the
# default bench (with cpu baseline) and the reference arm.  usage: tools/gpu_round_n.sh TAG
TAG=${1:-rn}
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
tail -c 200 gpurun_out/${TAG}_launches.log
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"twohop_score_kernel|twohop_compact|topk_hist|topk_count|topk_write|owner_scan" --launch-skip 69 --launch-count 23 \
  -o /tmp/ncu/${TAG}_full -f python bench.py --steps 1 --warmup 3 --slabs 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page source --csv -k regex:"twohop_score_kernel" > gpurun_out/${TAG}_twohop_source.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out | tail -12
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.json
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 300 gpurun_out/${TAG}_ref.json
echo done > gpurun_out/${TAG}_done
