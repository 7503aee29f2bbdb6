#!/bin/bash
# multi-GPU validation + bench (run with gpurun --gpus N)
N=${1:-2}; TAG=${2:-r02m}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/${TAG}_check.log 2>&1; tail -2 gpurun_out/${TAG}_check.log
MGPU_FORCE_GROUPED=1 timeout 600 $TR --master-port 29542 tests/mgpu_check.py > gpurun_out/${TAG}_check_grouped.log 2>&1; tail -2 gpurun_out/${TAG}_check_grouped.log
MGPU_SMALL_PANELS=1 timeout 600 $TR --master-port 29543 tests/mgpu_check.py > gpurun_out/${TAG}_check_panels.log 2>&1; tail -2 gpurun_out/${TAG}_check_panels.log
timeout 900 $TR --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 400 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_n${N}.json'))
    print('N=${N} ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'attempts', d['attempts_mean'], 'cost', d['final_cost'])
except Exception as e: print('no bench line', e)
PY
