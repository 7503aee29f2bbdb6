#!/bin/bash
# final 1-GPU pass of the round: full GPU test suite, the driver's bench command (both arms), ncu passes
TAG=${1:-r02x}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 120 -c 3 -o gpurun_out/${TAG}_prof_dgemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_full_dgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residual_jacobian_kernel -s 6 -c 1 -o gpurun_out/${TAG}_prof_jac \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_full_jac.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"potrf_tile_kernel|trinv_tile_kernel|accumulate_cells_kernel|trsv_forward_step_kernel" -s 20 -c 6 -o gpurun_out/${TAG}_prof_misc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-comparison > gpurun_out/${TAG}_ncu_full_misc.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'lib', (d.get('library_path') or {}).get('ms_per_step'), 'cpu', d.get('cpu_baseline',{}).get('seconds_per_iteration'), 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
PY
