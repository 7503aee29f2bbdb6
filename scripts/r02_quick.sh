#!/bin/bash
# quick validation + timing pass (1 GPU)
TAG=${1:-r02c}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or structured" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for bk in 32 16; do for sms in 8 0; do
  B200BA_GEMM_BK=$bk B200BA_PANEL_SMS=$sms python scripts/dense_timing.py 2>&1 | sed "s/^/bk$bk sms$sms: /" | tee -a gpurun_out/${TAG}_dense_timing.log
done; done
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own.json 2> gpurun_out/${TAG}_bench_own.err
B200BA_GEMM_BK=16 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_own_bk16.json 2> gpurun_out/${TAG}_bench_own_bk16.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench_own.json','gpurun_out/${TAG}_bench_own_bk16.json'):
    d=json.load(open(f))
    print(f, 'ms/step %.2f'%d['ms_per_step'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
PY
