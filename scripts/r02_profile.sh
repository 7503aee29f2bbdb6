#!/bin/bash
# Round-2 profiling pass (1 GPU): launch list of a short bench, full captures of the DMMA GEMM and the
# Jacobian kernel, and the library-path bench for comparison. Outputs under gpurun_out/ (summaries are
# copied to profiles/ by profiles/summarize.py).
set -x
TAG=${1:-r02b}
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own.json 2> gpurun_out/${TAG}_bench_own.err
B200BA_DENSE=lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_lib.json 2> gpurun_out/${TAG}_bench_lib.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 60 -c 3 -o gpurun_out/${TAG}_prof_dgemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_dgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residual_jacobian_kernel -s 4 -c 2 -o gpurun_out/${TAG}_prof_jac \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_jac.log 2>&1
ls -la gpurun_out | tail -20
