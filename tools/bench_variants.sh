#!/bin/bash
# usage: tools/bench_variants.sh TAG "ENV1" "ENV2" ...   (each ENV is a space-separated list of VAR=val, "-" for none)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  if [ "$envs" = "-" ]; then envs=""; fi
  env $envs timeout 180 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_v${i}.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_v${i}.json")); print("v${i} [${envs}]", round(d["value"]/1e6,1), {k: round(v,2) for k,v in d["detail"]["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as e: print("v${i} [${envs}] ERR", e)
PY
done
tail -3 gpurun_out/${TAG}_bench.err
