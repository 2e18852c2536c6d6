#!/bin/bash
# GPU box, round 2 call G: full-graph bench, N=1 (driver's flags), with the launch list under ncu of a short run.
mkdir -p gpurun_out
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.log
echo "bench rc=$?"; grep -v Warning gpurun_out/r2g_bench.log | tail -3 | cut -c1-300
python -c "
import json; j=json.load(open('gpurun_out/r2g_bench.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('phases', j['detail']['phase_ms_rank0'])
print('prefilter', j['detail']['prefilter'])
print('roofline', {k: j['roofline'][k] for k in ('kernel','achieved','frac','ms_per_launch','launches_per_step')})
for r in j['roofline_other']: print('  other', {k: r[k] for k in ('kernel','achieved','frac','ms_per_launch','launches_per_step')})
print('shapes', json.dumps(j['detail']['shapes'])[:1500])
print('library', j['detail'].get('library_baseline'))
print('cpu', j.get('cpu_baseline'))
print('clocks', j['clocks'], 'launches', j['gpu_launches'])
"
