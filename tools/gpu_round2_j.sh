#!/bin/bash
# GPU box, round 2 call J: K2 with TMA gather4 + 8 KB activation chunks: parity tests, then timing variants.
mkdir -p gpurun_out
L=gpurun_out/r2j_k2.log; : > $L
timeout 300 python -m pytest tests/test_gpu_mlp_tc.py -m gpu -x -q 2>&1 | tail -15 >> $L
if grep -q "passed" $L && ! grep -q "failed\|error" $L; then
  for v in "" "EPS_TC3_EPI=8" "EPS_TC3_GROUPS=3" "EPS_TC3_L2PROMO=0" "EPS_TC3_L2PROMO=128" "EPS_TC3_RING=5"; do
    env $v timeout 120 python tools/k2_bench.py 25 10 2>&1 | grep -v Warning >> $L
  done
fi
cat $L
