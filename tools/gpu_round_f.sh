#!/bin/bash
# multi-GPU box: 2-GPU sharded-filter test + torchrun bench at N GPUs.  usage: tools/gpu_round_f.sh TAG N
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 1500 gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
echo done > gpurun_out/${TAG}_done
