#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/r02u_pytest.log
echo "== lbs"; timeout 120 python scripts/lbs_sweep.py 64 128 256 512 1024 2>&1 | grep -E "lbs_us|Error" | tee $OUT/r02u_lbs.jsonl
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02u_lbs.jsonl
echo "== trace"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200_trace.so timeout 200 python scripts/lbs_trace.py 64 2>&1 | tail -38 | tee $OUT/r02u_trace.txt
