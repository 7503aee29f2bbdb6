#!/bin/bash
# A/B of the panel step of the blocked Cholesky: B200BA_PANEL=2 (blocked tile kernel, out-of-place solves) vs 1
TAG=${1:-r02z}
export B200BA_PANEL=2
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or trajectory or schur" > gpurun_out/${TAG}_pytest_panel2.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_panel2.log
timeout 120 python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing_panel2.log
B200BA_PANEL=1 timeout 120 python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing_panel1.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potrf_trinv_tile -c 6 --csv --log-file gpurun_out/${TAG}_tile_kernel.csv python scripts/dense_timing.py 2048 > /dev/null 2>&1; tail -4 gpurun_out/${TAG}_tile_kernel.csv
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench_panel2.json 2> gpurun_out/${TAG}_bench_panel2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_panel2.json'))
    print('panel2 N=1 ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d.get('parity_vs_1gpu'), d['attempts_mean'])
except Exception as e: print('no bench line', e)
PY
tail -c 300 gpurun_out/${TAG}_bench_panel2.err
