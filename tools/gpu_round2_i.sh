#!/bin/bash
# GPU box, round 2 call I: TMEM-read / fence microbenchmarks, K2 producer A/B (cp.async vs register buffers), no-gather ceiling.
mkdir -p gpurun_out
L=gpurun_out/r2i_k2.log; : > $L
timeout 60 tools/micro/bin/tmem_ld_bw >> $L 2>&1
timeout 60 tools/micro/bin/fence_vs_loads >> $L 2>&1
for t in 1 9 13 11 3; do
  EPS_TC3_TUNE=$t timeout 120 python tools/k2_bench.py 25 10 2>&1 | grep -v Warning >> $L
done
EPS_TC3_TUNE=9 EPS_TC3_GROUPS=3 timeout 120 python tools/k2_bench.py 25 10 2>&1 | grep -v Warning >> $L
cat $L
