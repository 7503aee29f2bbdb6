#!/bin/bash
# quick validation + timing pass (1 GPU)
TAG=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or trajectory or structured" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python scripts/dense_timing.py > gpurun_out/${TAG}_dense_timing.log 2>&1; cat gpurun_out/${TAG}_dense_timing.log
B200BA_PANEL_SMS=0 python scripts/dense_timing.py 2>&1 | sed 's/^/reserve0: /' | tee -a gpurun_out/${TAG}_dense_timing.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own.json 2> gpurun_out/${TAG}_bench_own.err
python scripts/evalhist.py > gpurun_out/${TAG}_evalhist.log 2>&1; tail -12 gpurun_out/${TAG}_evalhist.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_own.json'))
print('ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
PY
