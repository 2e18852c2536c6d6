#!/bin/bash
# GPU box: the whole -m gpu suite, smoke(), then the N=1 bench with the driver's flags.
TAG=${1:-r3h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
echo "bench rc=$?"; grep -v Warning gpurun_out/${TAG}_bench.log | tail -3 | cut -c1-300
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('phases', j['detail']['phase_ms_rank0'])
print('prefilter', j['detail']['prefilter'])
print('roofline', {k: j['roofline'][k] for k in ('kernel','achieved','frac','ms_per_launch','launches_per_step')})
for r in j['roofline_other']: print('  other', {k: r[k] for k in ('kernel','achieved','frac','ms_per_launch','launches_per_step')})
print('shapes', json.dumps(j['detail']['shapes'])[:1200])
print('library', j['detail'].get('library_baseline'))
print('clocks', j['clocks'], 'launches', j['gpu_launches'])
"
