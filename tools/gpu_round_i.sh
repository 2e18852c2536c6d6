#!/bin/bash
# GPU box: heuristics/candidate tests, then the bench with the one-pass and the two-pass enumeration.  usage: tools/gpu_round_i.sh TAG
TAG=${1:-ri}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_heuristics.py tests/test_gpu_cli.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_onepass.json 2> gpurun_out/${TAG}_bench_onepass.err; tail -c 400 gpurun_out/${TAG}_bench_onepass.json; tail -3 gpurun_out/${TAG}_bench_onepass.err
timeout 300 python bench.py --no-cpu-baseline --twopass > gpurun_out/${TAG}_bench_twopass.json 2> gpurun_out/${TAG}_bench_twopass.err; tail -c 400 gpurun_out/${TAG}_bench_twopass.json
timeout 300 python bench.py --no-cpu-baseline --workload ddi > gpurun_out/${TAG}_bench_ddi.json 2> gpurun_out/${TAG}_bench_ddi.err; tail -c 400 gpurun_out/${TAG}_bench_ddi.json
echo done > gpurun_out/${TAG}_done
