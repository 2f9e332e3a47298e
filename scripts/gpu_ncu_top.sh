#!/usr/bin/env bash
# ncu --set full of the two kernels that dominate the step: the recurrence kernel (plain cluster launch: ncu cannot replay
# cooperative launches) and the tensor-core GEMM (input projection BN = 128, then regressor GEMMs BN = 64).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-top}
GAITB200_GRU_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gru_recurrent' -s 3 -c 1 -f -o $OUT/${TAG}_gru \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_gru.log 2>&1
tail -2 $OUT/${TAG}_gru.log | cut -c1-160
GAITB200_GRU_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3' -s 36 -c 4 -f -o $OUT/${TAG}_gemm \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_gemm.log 2>&1
tail -2 $OUT/${TAG}_gemm.log | cut -c1-160
