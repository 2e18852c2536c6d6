#!/bin/bash
# GPU box, round 2 call E: score statistics of the trained-like bench model, 10 % bench run, then the full-graph bench.
mkdir -p gpurun_out
timeout 300 python tools/prefilter_stats.py ppa 26 > gpurun_out/r2e_stats.log 2>&1
grep -v Warning gpurun_out/r2e_stats.log | tail -16
timeout 600 python bench.py --steps 2 --warmup 3 --owners-frac 0.1 --no-cpu-baseline > gpurun_out/r2e_bench10.json 2> gpurun_out/r2e_bench10.log
echo "bench10 rc=$?"; tail -3 gpurun_out/r2e_bench10.log | cut -c1-300; head -c 1800 gpurun_out/r2e_bench10.json; echo
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.log
echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.log | cut -c1-300; head -c 1500 gpurun_out/r2e_bench.json
