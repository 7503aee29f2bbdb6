#!/bin/bash
# full GPU test suite + bench lines of the 1-GPU configurations
TAG=${1:-r02i}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
for c in 1 3 4; do
  python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config$c.json 2> gpurun_out/${TAG}_bench_config$c.err
done
python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_config*.json')):
    try:
        d=json.load(open(f))
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'frac %.2f'%d['roofline']['frac'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['attempts_per_step'])
    except Exception as e: print(f, 'FAILED', e)
PY
