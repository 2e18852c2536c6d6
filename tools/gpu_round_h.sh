#!/bin/bash
# GPU box: selected test files + smoke.  usage: tools/gpu_round_h.sh TAG "tests/a.py tests/b.py"
TAG=${1:-rh}; FILES=${2:-tests}
mkdir -p gpurun_out
( time timeout 900 python -m pytest $FILES -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -40 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -3 gpurun_out/${TAG}_smoke.log
echo done > gpurun_out/${TAG}_done
