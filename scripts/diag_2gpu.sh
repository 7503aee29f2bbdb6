#!/bin/bash
# diagnostic: 2-GPU grouped check + per-rank bench phases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
echo "nproc=$(nproc)"; nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
nvidia-smi topo -m | head -8
MGPU_FORCE_GROUPED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_check.py 2>&1 | grep -v "^W\|OMP_NUM" | tail -30
echo "=== bench 2gpu default"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms_per_step'])"
echo "=== plain mgpu"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/mgpu_check.py 2>&1 | grep MGPU
} > gpurun_out/diag_2gpu.log 2>&1
tail -60 gpurun_out/diag_2gpu.log
