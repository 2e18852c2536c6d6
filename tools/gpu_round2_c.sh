#!/bin/bash
# GPU box, round 2 call C: prefilter shape tests, K2 micro-benchmark (baseline), full-graph bench (N=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -q -s -k "ddi or ppa" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "prefilter|passed|failed|rc=" gpurun_out/r2c_pytest.log | tail -8
python tools/k2_bench.py 25 5 > gpurun_out/r2c_k2.log 2>&1; tail -2 gpurun_out/r2c_k2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.log
echo "bench rc=$?"
tail -5 gpurun_out/r2c_bench.log
head -c 2500 gpurun_out/r2c_bench.json
