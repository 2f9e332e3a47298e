#!/usr/bin/env bash
# multi-GPU check: N-rank == 1-rank parity tests (NCCL + peer modes), then the driver's own launch line for N ranks
set -u
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT; TAG=${2:-r02}
nvidia-smi --query-gpu=index,name --format=csv | head -10
nvidia-smi topo -m 2>/dev/null | head -14
echo "== gather parity tests"
timeout 600 python -m pytest tests/test_gpu_gather.py -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gather_n${N}.log
echo "== N=$N b200 (configs[2], gather auto)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
tail -c 1500 $OUT/${TAG}_bench_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3))
    print(json.dumps(d['gather'])[:1500]); print(d['configs']); print(d['config']['sharding'])
except Exception as e: print('bench parse failed',e)
PY
if [ "${REFARM:-1}" = "1" ]; then
echo "== N=$N reference arm"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600 | tee $OUT/${TAG}_bench_n${N}_ref.json
fi
