#!/bin/bash
# GPU box (N GPUs): the strong-scaled full-graph bench with its self-check.   usage: tools/gpu_ngpu_check.sh N TAG
N=${1:-4}; TAG=${2:-r3v}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.log
echo "bench$N rc=$?"; grep -v Warning gpurun_out/${TAG}_bench$N.log | tail -4 | cut -c1-300
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench$N.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'verified', j['detail']['multi_gpu_verified'])
print('phases', j['detail']['phase_ms_rank0'])
"
