#!/bin/bash
# final state, defaults only: smoke + the complete bench line (own + library comparison + CPU arm + e2e)
TAG=${1:-r02zd}
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 420 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'lib', d.get('library_path',{}).get('ms_per_step'), 'cpu', d['cpu_baseline'].get('seconds_per_iteration'), 'e2e', d['e2e'].get('ms_per_step'), 'frac', d['roofline']['frac'], d.get('parity_vs_1gpu'), d['attempts_mean'])
except Exception as e: print('no bench line', e)
PY
tail -c 300 gpurun_out/${TAG}_bench.err
