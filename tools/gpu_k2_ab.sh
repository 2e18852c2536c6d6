#!/bin/bash
# GPU box: K2 parity tests, then timing variants (tag = $1)
TAG=${1:-r2q}
mkdir -p gpurun_out
L=gpurun_out/${TAG}_k2.log; : > $L
timeout 300 python -m pytest tests/test_gpu_mlp_tc.py -m gpu -x -q 2>&1 | tail -15 >> $L
if grep -q "passed" $L && ! grep -q "failed\|error" $L; then
  for v in "" "EPS_TC3_TUNE=0" "EPS_TC3_TUNE=0 EPS_TC3_SHAPE=42" "EPS_TC3_TUNE=0 EPS_TC3_SHAPE=81" "EPS_TC3_TUNE=0 EPS_TC3_SHAPE=82" "EPS_TC3_SHAPE=82"; do
    env $v timeout 120 python tools/k2_bench.py 25 10 2>&1 | grep -v Warning >> $L
  done
fi
cat $L
