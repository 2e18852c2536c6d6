#!/bin/bash
# usage (on the GPU box): tools/prof_round.sh TAG  -> gpurun_out/TAG_launches.csv, gpurun_out/TAG_full.ncu-rep
TAG=${1:-prof}
mkdir -p gpurun_out
# launch list of one bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
# full-set capture of the scoring kernels of the third step
ncu --set full --clock-control none --import-source on \
  -k regex:"linkpred_tc3|cn_grouped|spmm_csr|twohop|fused" --launch-skip ${SKIP:-14} --launch-count ${COUNT:-7} \
  -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out
