#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
for v in "" _mma2; do
  echo "== lib$v"
  GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "lbs or smpl or head or three_joint or independent" 2>&1 | tail -2
  GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/lbs_sweep.py 64 128 256 512 1024 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02v_lbs.jsonl
done
echo "== joints-only (skin)"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02v_lbs.jsonl
