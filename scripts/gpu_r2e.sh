#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "joint_regress or vpregressor or regressor_vs or smpl" 2>&1 | tail -8 | tee $OUT/r02e_pytest.log
for v in "" _noa _nov _noav; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python scripts/lbs_sweep.py 64 512 2>&1 | tee -a $OUT/r02e_lbs_exp.jsonl
done
echo "== joints-only noa"; LBS_JOINTS_ONLY=1 GAITB200_LIB=$PWD/$PKG/lib/libgaitb200_noa.so timeout 300 python scripts/lbs_sweep.py 64 512 2>&1 | tee -a $OUT/r02e_lbs_exp.jsonl
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'lbs frac',round(d['roofline']['frac'],3)); print(json.dumps(d['stages']['jreg']))"
