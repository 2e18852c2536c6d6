mkdir -p gpurun_out
for t in 0 1 2 3; do
  for w in ddi ppa; do
    EPS_TC3_TUNE=$t timeout -k 10 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --pairs 33554432 > gpurun_out/r15_${w}_t$t.log 2>&1
  done
done
echo done > gpurun_out/r15_done
