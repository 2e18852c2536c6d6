#!/bin/bash
# GPU box: K2 tests, then the ppa bench over u-block budgets.  usage: tools/gpu_round_m.sh TAG "0 32 48 64 96"
TAG=${1:-rm}; MBS=${2:-"0 32 48 64 96"}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_gnn.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
for mb in $MBS; do
  EPS_TC3_UBLOCK_MB=$mb timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_mb$mb.json 2> gpurun_out/${TAG}_bench_mb$mb.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_mb$mb.json")); print("mb=$mb", round(d["value"]/1e6,1), {k: round(v,2) for k,v in d["detail"]["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as e: print("mb=$mb ERR", e)
PY
done
echo done > gpurun_out/${TAG}_done
