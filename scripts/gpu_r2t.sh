#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r02t_pytest.log
for v in "" _tf32; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/lbs_sweep.py 64 128 256 512 1024 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02t_lbs.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02t_lbs.jsonl
timeout 200 python scripts/shard_err.py 64 2>&1 | tail -4
