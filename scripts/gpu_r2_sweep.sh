#!/usr/bin/env bash
# round 2 baseline: GPU test suite, GRU crossover sweep (default / per-step / persistent chunks at any S), LBS rate vs F
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/r02a_pytest_gpu.log
echo "== default"; timeout 600 python scripts/gru_s_sweep.py 64 128 192 256 512 1024 2>&1 | tee $OUT/r02a_sweep_default.jsonl
echo "== per-step path"; GAITB200_GRU_PATH=1 timeout 600 python scripts/gru_s_sweep.py 64 128 192 256 512 1024 2>&1 | tee $OUT/r02a_sweep_perstep.jsonl
echo "== persistent chunks"; GAITB200_GRU_MAXCHUNKED=100000 timeout 600 python scripts/gru_s_sweep.py 256 512 1024 2>&1 | tee $OUT/r02a_sweep_chunked.jsonl
