#!/usr/bin/env bash
# multi-GPU check: the driver's own launch line for N ranks, own arm and reference arm
set -u
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "== N=$N b200"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -3 | tee $OUT/multi_n${N}.json
echo "== N=$N b200 with mesh gather"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 5 --gather mesh 2>&1 | tail -2 | tee $OUT/multi_n${N}_gather.json
echo "== N=$N reference arm"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/multi_n${N}_ref.json
echo "== N=1 for comparison"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/multi_n1.json
