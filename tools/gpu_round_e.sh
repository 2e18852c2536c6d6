#!/bin/bash
# GPU box: rebuild with the tc3 timeline trace and dump one trace per env variant.  usage: tools/gpu_round_e.sh TAG "ENV1" ...
TAG=$1; shift
mkdir -p gpurun_out
EPS_EXTRA_NVCC_FLAGS=-DEPS_TC3_TRACE python -m edge_proposal_sets_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
i=0
for envs in "$@"; do
  i=$((i+1)); if [ "$envs" = "-" ]; then envs=""; fi
  env $envs EPS_TC3_TRACE_FILE=gpurun_out/${TAG}_trace${i}.bin PYTHONPATH=. timeout 120 python tools/tc3_trace.py run
done
echo done > gpurun_out/${TAG}_done
