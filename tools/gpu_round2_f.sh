#!/bin/bash
# GPU box, round 2 call F: the fp16 tensor arm — MLP tests, K2 micro-benchmark, shape tests, statistics, 10 % bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp_tc.py tests/test_model_golden.py -q -s > gpurun_out/r2f_pytest_mlp.log 2>&1; grep -E "max|overlap|scale|passed|failed|Error|error" gpurun_out/r2f_pytest_mlp.log | tail -30
python tools/k2_bench.py 25 5 > gpurun_out/r2f_k2.log 2>&1; tail -1 gpurun_out/r2f_k2.log
timeout 300 python tools/prefilter_stats.py ppa 26 > gpurun_out/r2f_stats.log 2>&1; grep -v Warning gpurun_out/r2f_stats.log | tail -9
timeout 900 python -m pytest tests/test_gpu_shapes.py -q -s -k "ddi or ppa" > gpurun_out/r2f_pytest_shapes.log 2>&1; grep -E "prefilter|passed|failed|Error|assert" gpurun_out/r2f_pytest_shapes.log | tail -12
timeout 300 python bench.py --steps 2 --warmup 3 --owners-frac 0.1 --no-cpu-baseline --no-extras > gpurun_out/r2f_bench10.json 2> gpurun_out/r2f_bench10.log
echo "bench10 rc=$?"; tail -3 gpurun_out/r2f_bench10.log | cut -c1-300; python -c "
import json; j=json.load(open('gpurun_out/r2f_bench10.json')); print(j['value'], j['ms_per_step'], j['detail']['phase_ms_rank0'], j['detail']['prefilter'], j['detail'].get('prefilter_fallback'), j['e2e'])"
