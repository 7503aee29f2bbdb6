#!/bin/bash
TAG=${1:-r02w}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dense" > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('N=1 ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d.get('parity_vs_1gpu'))
PY
bash scripts/r02_mgpu.sh 2 ${TAG} 2>&1 | grep -v "Warning\|return func" | tail -6
