#!/bin/bash
# bench lines of configs 1, 3 (1 GPU) and 4 (1 and 2 GPUs: BASELINE quotes it on 2 x B200)
TAG=${1:-r02t}
for c in 1 3 4; do
  python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c${c}_n1.json 2> gpurun_out/${TAG}_bench_c${c}_n1.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --config 4 --gpus 2 --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_c4_n2.json 2> gpurun_out/${TAG}_bench_c4_n2.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_c*.json')):
    try:
        d=json.load(open(f))
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'frac %.2f'%d['roofline']['frac'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['attempts_per_step'], 'lib', (d.get('library_path') or {}).get('ms_per_step'))
    except Exception as e: print(f, 'FAILED', e)
PY
