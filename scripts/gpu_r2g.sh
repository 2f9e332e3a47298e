#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r02g_pytest.log
for v in "" _nmma1; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/lbs_sweep.py 64 128 512 1024 2>&1 | tee -a $OUT/r02g_lbs.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | tee -a $OUT/r02g_lbs.jsonl
for v in "" _jrows1noload; do
  echo "== jreg lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/jreg_time.py 1024 4096 2>&1 | tee -a $OUT/r02g_jreg.jsonl
done
