#!/bin/bash
# second version of the triangular solves (B200BA_TRSV=2) against the first
TAG=${1:-r02zc}
export B200BA_TRSV=2
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or schur" > gpurun_out/${TAG}_pytest_a.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_a.log
timeout 120 python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:potrf_trinv_tile -s 3 -c 1 -o gpurun_out/${TAG}_prof_tile -f python scripts/dense_timing.py 2048 > gpurun_out/${TAG}_prof_tile.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('N=1 ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d.get('parity_vs_1gpu'), d['attempts_mean'])
except Exception as e: print('no bench line', e)
PY
tail -c 300 gpurun_out/${TAG}_bench.err
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "config5_dense_size or (full_size_other_configs and 4) or trajectory" > gpurun_out/${TAG}_pytest_b.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_b.log
