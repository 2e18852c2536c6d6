#!/bin/bash
# GPU box: final validation — full gpu suite, smoke, default bench (with cpu baseline), reference arm.  usage: tools/gpu_round_p.sh TAG
TAG=${1:-rp}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --cpu-seconds 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 --cpu-seconds 8 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 200 gpurun_out/${TAG}_ref.json
echo done > gpurun_out/${TAG}_done
