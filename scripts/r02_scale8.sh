#!/bin/bash
# 8-GPU box: config 2 at N = 8 and 4, config 5 at N = 8 and 1
TAG=${1:-r02s}
run() { # N config steps port
  N=$1; C=$2; K=$3; P=$4
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --config $C --steps $K --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_c${C}_n${N}.json 2> gpurun_out/${TAG}_c${C}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --config $C --gpus $N --steps $K --warmup 3 > gpurun_out/${TAG}_c${C}_n${N}.json 2> gpurun_out/${TAG}_c${C}_n${N}.err
  fi
  tail -c 300 gpurun_out/${TAG}_c${C}_n${N}.err | tr '\n' ' '; echo
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_c${C}_n${N}.json'))
    print('config $C N=$N ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'attempts', d['attempts_per_step'], 'cost %.9f'%d['final_cost'], d.get('parity_vs_1gpu'))
except Exception as e: print('config $C N=$N: no line', e)
PY
}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
run 8 2 6 29551
run 4 2 6 29552
run 8 5 3 29553
run 1 5 3 0
