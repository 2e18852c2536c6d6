#!/bin/bash
# usage (on the GPU box): tools/prof_round2.sh TAG -> gpurun_out/TAG_launches.csv, gpurun_out/TAG_full_raw.csv (+ .ncu-rep)
TAG=${1:-r2p}
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --owners-frac 0.06 --no-cpu-baseline --no-extras"
# launch list of one bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# full-set capture of the hot kernels in the timed step (skip the warm-up launches of each kernel by name)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"linkpred_tc3_kernel|twohop_score_kernel|twohop_compact_kernel|threshold_count_kernel|threshold_write_kernel|spmm_csr_kernel|linkpred_fp32_kernel|topk_hist_kernel" \
  --launch-skip ${SKIP:-60} --launch-count ${COUNT:-24} -o /tmp/ncu/${TAG}_full -f $CMD > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
cp /tmp/ncu/${TAG}_full.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -6
