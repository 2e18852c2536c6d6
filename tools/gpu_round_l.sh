#!/bin/bash
# GPU box: candidate / top-k / cli / multi tests + the three shape benches.  usage: tools/gpu_round_l.sh TAG
TAG=${1:-rl}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_heuristics.py tests/test_gpu_topk.py tests/test_gpu_cli.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
for w in ppa ddi collab; do
  timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; tail -c 300 gpurun_out/${TAG}_bench_$w.json; tail -2 gpurun_out/${TAG}_bench_$w.err
done
echo done > gpurun_out/${TAG}_done
