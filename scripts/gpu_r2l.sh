#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r02l_pytest.log
echo "== lbs"; timeout 120 python scripts/lbs_sweep.py 64 128 256 512 1024 2>&1 | grep -E "lbs_us|Error" | tee $OUT/r02l_lbs.jsonl
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02l_lbs.jsonl
echo "== jreg"; timeout 120 python scripts/jreg_time.py 1024 4096 2>&1 | tee $OUT/r02l_jreg.jsonl
