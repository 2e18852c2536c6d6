#!/bin/bash
# GPU box: quick parity of the changed kernels, then bench A/B over env variants.  usage: tools/gpu_round_c.sh TAG "ENV1" "ENV2" ...
TAG=$1; shift
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_topk.py -x -q --timeout 60 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
tools/bench_variants.sh ${TAG} "$@"
echo done > gpurun_out/${TAG}_done
