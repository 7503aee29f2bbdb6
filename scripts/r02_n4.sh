#!/bin/bash
# 4-GPU box: the look-ahead schedule at R = 4 (narrow panels on the small problems, then config 2 at full size)
TAG=${1:-r02y}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
MGPU_SMALL_PANELS=1 timeout 200 $TR --master-port 29543 tests/mgpu_check.py > gpurun_out/${TAG}_check_panels_n4.log 2>&1; tail -2 gpurun_out/${TAG}_check_panels_n4.log
timeout 240 $TR --master-port 29544 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_bench_c2_n4.json 2> gpurun_out/${TAG}_bench_c2_n4.err
tail -c 300 gpurun_out/${TAG}_bench_c2_n4.err | tr '\n' ' '; echo
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_c2_n4.json'))
    print('N=4 ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'attempts', d['attempts_mean'], d.get('parity_vs_1gpu'))
except Exception as e: print('no bench line', e)
PY
