#!/bin/bash
# quick validation + timing pass (1 GPU)
TAG=${1:-r02c}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dense" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python scripts/dense_timing.py 2>&1 | tee -a gpurun_out/${TAG}_dense_timing.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own.json 2> gpurun_out/${TAG}_bench_own.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_own*.json')):
    try:
        d=json.load(open(f))
        print(f, 'ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
    except Exception as e: print(f, 'FAILED', e)
PY
