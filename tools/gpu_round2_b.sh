#!/bin/bash
# GPU box, round 2 call B: the tests that failed in call A + the new ones, then the full-graph bench (N=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_cli.py::test_twitch_sage_filter_with_real_features_then_aa_rank_sweep \
  tests/test_gpu_mlp_tc.py::test_tc_context_reuses_table_and_weight_images -q -s --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.log
echo "bench rc=$?"
tail -5 gpurun_out/r2b_bench.log
head -c 3000 gpurun_out/r2b_bench.json
