#!/bin/bash
# GPU box: A/B of the fused kernel's CTA size.  usage: tools/gpu_round_o.sh TAG
TAG=${1:-ro}
mkdir -p gpurun_out
for t in 512 1024; do
  for w in ppa collab ddi; do
    EPS_TS_THREADS=$t timeout 200 python bench.py --no-cpu-baseline --workload $w --steps 4 > gpurun_out/${TAG}_${w}_t$t.json 2> gpurun_out/${TAG}_${w}_t$t.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${w}_t$t.json")); print("$w threads=$t", round(d["value"]/1e6,1), {k: round(v,2) for k,v in d["detail"]["phase_ms"].items()})
except Exception as e: print("$w threads=$t ERR", e)
PY
  done
done
echo done > gpurun_out/${TAG}_done
