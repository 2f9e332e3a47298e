#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-gemm}
# first 3 gemm launches after warm-up step: input projection (BN=128) and two recurrent steps (BN=64, split-K)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3' -s 45 -c 3 -f -o $OUT/${TAG}_ncu \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
