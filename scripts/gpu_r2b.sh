#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r02b_pytest_gpu.log
echo "== lbs rate at large F"; GAITB200_GRU_PATH=1 timeout 600 python scripts/gru_s_sweep.py 256 512 1024 2>&1 | tee $OUT/r02b_sweep.jsonl
