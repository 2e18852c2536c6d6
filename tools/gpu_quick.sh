#!/bin/bash
# GPU box: K2 / fp32-arm / model tests + a short bench (10 % of the owners) for the phase times
TAG=${1:-r3t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_gnn.py tests/test_model_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --owners-frac 0.1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', j['value'], 'ms', j['ms_per_step'])
print('phases', j['detail']['phase_ms_rank0'])
print('identical', j['detail'].get('topk_identical_to_fp32'), 'roofline', j['roofline']['frac'])
"
