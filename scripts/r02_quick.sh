#!/bin/bash
# quick validation + timing pass (1 GPU)
TAG=${1:-r02c}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or structured" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for v in 64 128; do
  B200BA_GEMM=$v python scripts/dense_timing.py 2>&1 | sed "s/^/gemm$v: /" | tee -a gpurun_out/${TAG}_dense_timing.log
done
for v in 64 128; do
B200BA_GEMM=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own_$v.json 2> gpurun_out/${TAG}_bench_own_$v.err
done
B200BA_GROUP_BLOCKS=160 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own_g160.json 2> gpurun_out/${TAG}_bench_own_g160.err
B200BA_DENSE_NB=512 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own_nb512.json 2> gpurun_out/${TAG}_bench_own_nb512.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_own_*.json')):
    try:
        d=json.load(open(f))
        print(f, 'ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
    except Exception as e: print(f, 'FAILED', e)
PY
