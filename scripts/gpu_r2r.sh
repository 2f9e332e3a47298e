#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
for v in "" _nw12 _nw12d5 _nw12d3; do
  echo "== jreg lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k joint_regress_stream 2>&1 | tail -1
  GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/jreg_time.py 1024 4096 2>&1 | tee -a $OUT/r02r_jreg.jsonl
  JREG_ROWS=9 GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/jreg_time.py 1024 2>&1 | tee -a $OUT/r02r_jreg.jsonl
done
echo "== racecheck GEMM only"
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 4 python -m pytest tests/test_gpu_parity.py -x -q -k "linear_vs_torch_fp32" > $OUT/r02r_racecheck_gemm.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" $OUT/r02r_racecheck_gemm.txt | tail -3
grep -E "hazard detected|Race reported|and \(|Write Thread|Read Thread|Current Value" $OUT/r02r_racecheck_gemm.txt | head -12 | cut -c1-300
