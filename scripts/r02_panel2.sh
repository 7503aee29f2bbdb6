#!/bin/bash
# second A/B of the panel step: revised tile kernel (pipelined column steps), aux stream for U(k, k+2)
TAG=${1:-r02zb}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or trajectory or schur" > gpurun_out/${TAG}_pytest_a.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_a.log
timeout 120 python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing.log
B200BA_AUX=0 timeout 120 python scripts/dense_timing.py 2>&1 | tee gpurun_out/${TAG}_dense_timing_noaux.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potrf_trinv_tile -c 4 --csv --log-file gpurun_out/${TAG}_tile_kernel.csv python scripts/dense_timing.py 2048 > /dev/null 2>&1; tail -2 gpurun_out/${TAG}_tile_kernel.csv | cut -d, -f5,15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('N=1 ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d.get('parity_vs_1gpu'), d['attempts_mean'])
except Exception as e: print('no bench line', e)
PY
tail -c 300 gpurun_out/${TAG}_bench.err
timeout 420 python -m pytest tests -x -q -m gpu -k "not full_size and not config5 and not dense and not trajectory and not schur" > gpurun_out/${TAG}_pytest_b.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_b.log
