#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lbs or smpl or head" 2>&1 | tail -4 | tee $OUT/r02f_pytest.log
for v in "" _nmma1; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python scripts/lbs_sweep.py 64 128 512 1024 2>&1 | tee -a $OUT/r02f_lbs.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 300 python scripts/lbs_sweep.py 64 512 2>&1 | tee -a $OUT/r02f_lbs.jsonl
for v in "" _jrows1 _jnoload _jrows1noload; do
  echo "== jreg lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python scripts/jreg_time.py 1024 4096 2>&1 | tee -a $OUT/r02f_jreg.jsonl
done
JREG_VARIANT=dense timeout 300 python scripts/jreg_time.py 1024 2>&1 | tee -a $OUT/r02f_jreg.jsonl
JREG_ROWS=9 timeout 300 python scripts/jreg_time.py 1024 2>&1 | tee -a $OUT/r02f_jreg.jsonl
