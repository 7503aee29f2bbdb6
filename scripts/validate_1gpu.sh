#!/bin/bash
# round-end style validation on one B200: GPU tests, smoke, bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
main() {
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_stderr.log > gpurun_out/bench_latest.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_latest.json'))
print('bench', d['value'], d['ms_per_step'], d['phases_ms_per_step'])
print('e2e', d['e2e']); print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms']); print('clocks', d['clocks']); print('cpu', d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
}
main > gpurun_out/validate_1gpu.log 2>&1
tail -40 gpurun_out/validate_1gpu.log
