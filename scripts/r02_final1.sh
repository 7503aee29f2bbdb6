#!/bin/bash
TAG=${1:-r02j}
python scripts/cfg4_check.py > gpurun_out/${TAG}_cfg4_check.log 2>&1; tail -4 gpurun_out/${TAG}_cfg4_check.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 120 -c 4 -o gpurun_out/${TAG}_prof_dgemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_full_dgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residual_jacobian_kernel -s 4 -c 2 -o gpurun_out/${TAG}_prof_jac \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_full_jac.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'lib', d.get('library_path'), 'cpu', d.get('cpu_baseline',{}).get('seconds_per_iteration'))
PY
