#!/bin/bash
# GPU box: K2 parity tests, then timing variants on the micro list (tag = $1; variants = remaining args, default one plain run)
TAG=${1:-k2ab}; shift
mkdir -p gpurun_out
L=gpurun_out/${TAG}_k2.log; : > $L
timeout 300 python -m pytest tests/test_gpu_mlp_tc.py -m gpu -x -q 2>&1 | tail -15 >> $L
if grep -q "passed" $L && ! grep -q "failed\|error" $L; then
  if [ $# -eq 0 ]; then set -- ""; fi
  for v in "$@"; do
    env $v timeout 120 python tools/k2_bench.py 25 10 2>&1 | grep -v Warning >> $L
  done
fi
cat $L
