#!/bin/bash
# GPU box (2 GPUs), round 2 call H: the 2-GPU test and the strong-scaled full-graph bench with its self-check.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r4k_pytest.log 2>&1; tail -5 gpurun_out/r4k_pytest.log | cut -c1-400
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r4k_bench2.json 2> gpurun_out/r4k_bench2.log
echo "bench2 rc=$?"; grep -v Warning gpurun_out/r4k_bench2.log | tail -4 | cut -c1-400
python -c "
import json; j=json.load(open('gpurun_out/r4k_bench2.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'verified', j['detail']['multi_gpu_verified'])
print('phases', j['detail']['phase_ms_rank0'])
print('prefilter', j['detail']['prefilter'])
print('shapes', {k: (v['value'], v['ms_per_step'], v['multi_gpu_verified'], v['phase_ms_rank0']) for k, v in j['detail']['shapes'].items()})
"
