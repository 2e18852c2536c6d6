#!/bin/bash
# GPU box: full validation of HEAD — gpu tests, smoke, bench (both arms) + ddi / collab shape lines.  usage: tools/gpu_round_g.sh TAG
TAG=${1:-rg}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
for w in ddi collab; do
  timeout 300 python bench.py --workload $w --cpu-seconds 6 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; tail -c 300 gpurun_out/${TAG}_bench_$w.json; tail -2 gpurun_out/${TAG}_bench_$w.err
done
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 400 gpurun_out/${TAG}_ref.json
echo done > gpurun_out/${TAG}_done
