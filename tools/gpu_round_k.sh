#!/bin/bash
# GPU box: heuristics + topk + cli tests, bench collab/ppa, then the ncu launch list of one ppa bench step.  usage: tools/gpu_round_k.sh TAG
TAG=${1:-rk}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_heuristics.py tests/test_gpu_topk.py tests/test_gpu_cli.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
for w in collab ppa; do
  timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; tail -c 300 gpurun_out/${TAG}_bench_$w.json; tail -2 gpurun_out/${TAG}_bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
tail -c 300 gpurun_out/${TAG}_launches.log
echo done > gpurun_out/${TAG}_done
