#!/bin/bash
TAG=${1:-r02ze}
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d.get('parity_vs_1gpu'), d['attempts_mean'])
except Exception as e: print('no bench line', e)
PY
tail -c 200 gpurun_out/${TAG}_bench.err
