#!/bin/bash
# GPU box: ncu --set full of the scoring kernels of one step; only CSV exports travel back
# (gpurun_out/ is capped at 64 MiB).  usage: tools/gpu_round_b.sh TAG
TAG=${1:-rb}
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"linkpred_tc3|spmm_csr|twohop_score_kernel" --launch-skip ${SKIP:-10} --launch-count ${COUNT:-5} \
  -o /tmp/ncu/${TAG}_full -f python bench.py --steps 1 --warmup 3 --slabs 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page details --csv > gpurun_out/${TAG}_full_details.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page source --csv -k regex:"linkpred_tc3" > gpurun_out/${TAG}_tc3_source.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}_full.ncu-rep --page source --csv -k regex:"spmm_csr" --launch-count 1 > gpurun_out/${TAG}_spmm_source.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out
timeout 600 ncu --set full --clock-control none \
  -k regex:"topk_|sort_scatter|twohop_kernel|sgemm|gcn_norm" --launch-skip ${SKIP2:-60} --launch-count 30 \
  -o /tmp/ncu/${TAG}_full2 -f python bench.py --steps 1 --warmup 3 --slabs 1 --no-cpu-baseline > gpurun_out/${TAG}_full2.log 2>&1
ncu -i /tmp/ncu/${TAG}_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_full2_raw.csv 2>/dev/null
du -sh gpurun_out
echo done > gpurun_out/${TAG}_done
