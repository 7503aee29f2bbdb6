#!/bin/bash
# tuning sweep of the main Jacobian pass (launch geometry, register budget, evaluation budget)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
out=gpurun_out/sweep_jac.log
: > $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('jac_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'step', round(d['ms_per_step'],2), 'strag', round(d['phases_ms_per_step']['straggler'],2), 'trial', round(d['phases_ms_per_step']['trial'],3), 'rmse', d['rmse_px'])" >> $out 2>&1
}
run B200BA_JAC_MINB=4
run B200BA_JAC_MINB=4 B200BA_JAC_THREADS=64
run B200BA_JAC_MINB=4 B200BA_JAC_THREADS=32
run B200BA_JAC_MINB=3
run B200BA_JAC_MINB=3 B200BA_JAC_THREADS=64
run B200BA_JAC_MINB=2 B200BA_JAC_THREADS=64
run B200BA_JAC_MINB=5
run B200BA_JAC_MINB=5 B200BA_JAC_THREADS=64
run B200BA_JAC_MINB=6
run B200BA_JAC_MINB=6 B200BA_JAC_THREADS=64
run B200BA_JAC_MINB=4 B200BA_EVAL_BUDGET=8
run B200BA_JAC_MINB=4 B200BA_EVAL_BUDGET=12
run B200BA_JAC_MINB=4 B200BA_EVAL_BUDGET=24
cat $out
