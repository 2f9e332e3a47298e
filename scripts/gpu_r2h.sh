#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r02h_pytest.log
for v in "" _mb1 _mb4 _nocons _noconsmb1; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep lbs_us | tee -a $OUT/r02h_lbs.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep lbs_us | tee -a $OUT/r02h_lbs.jsonl
echo "== c4"; timeout 300 python scripts/gru_s_sweep.py 1 2 2>&1 | tail -3
