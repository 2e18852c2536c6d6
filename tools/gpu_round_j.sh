#!/bin/bash
# GPU box: full gpu test suite + smoke + the three shape benches.  usage: tools/gpu_round_j.sh TAG
TAG=${1:-rj}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
for w in ppa ddi collab; do
  timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; tail -c 300 gpurun_out/${TAG}_bench_$w.json; tail -2 gpurun_out/${TAG}_bench_$w.err
done
echo done > gpurun_out/${TAG}_done
