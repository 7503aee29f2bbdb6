#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
main() {
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "column_block_cholesky or schur_known" > gpurun_out/diag_chol_pytest.log 2>&1
rc=$?
tail -25 gpurun_out/diag_chol_pytest.log
if [ $rc -ne 0 ]; then echo "single-GPU blocked Cholesky test FAILED rc=$rc: stopping"; set +x; return 0; fi
MGPU_DIST_CHOL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tests/mgpu_check.py 2>&1 | grep -v "^W\|OMP_NUM\|^\*" | tail -12
for v in "X=0" "B200BA_DIST_CHOL=0" "B200BA_CHOL_NB=256" "B200BA_CHOL_NB=1024"; do
echo "=== bench 2gpu $v"
env $v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['rmse_px'], d['phases_ms_per_step'])"
done
echo "=== bench 1gpu blocked"
B200BA_DIST_CHOL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['rmse_px'], d['phases_ms_per_step'])"
}
main > gpurun_out/diag_chol.log 2>&1
tail -40 gpurun_out/diag_chol.log
