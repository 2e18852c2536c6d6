#!/bin/bash
# GPU box, round 2 call D: K2 with two epilogue warps per TMEM quarter (A/B vs 4), MLP tests, score statistics.
mkdir -p gpurun_out
( EPS_TC3_EPI=4 python tools/k2_bench.py 25 5; python tools/k2_bench.py 25 5 ) > gpurun_out/r2d_k2.log 2>&1; tail -2 gpurun_out/r2d_k2.log
timeout 600 python -m pytest tests/test_gpu_mlp_tc.py -q -x > gpurun_out/r2d_pytest.log 2>&1; tail -3 gpurun_out/r2d_pytest.log
timeout 400 python tools/prefilter_stats.py ddi 26 > gpurun_out/r2d_stats.log 2>&1
timeout 400 python tools/prefilter_stats.py ppa 26 >> gpurun_out/r2d_stats.log 2>&1
grep -v Warning gpurun_out/r2d_stats.log | tail -40
