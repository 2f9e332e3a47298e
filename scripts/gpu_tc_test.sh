#!/usr/bin/env bash
# Quick check of the tcgen05 GEMM path (guards against pipeline hangs with timeouts).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-tc}
echo "== linear tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "linear or gru or temporal" 2>&1 | tail -30 | tee $OUT/${TAG}_linear.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee $OUT/${TAG}_pytest.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee $OUT/${TAG}_bench.json
echo "== bench simt"; GAITB200_LINEAR=simt timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_simt.json
