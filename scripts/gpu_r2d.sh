#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lbs or smpl or head" 2>&1 | tail -5 | tee $OUT/r02d_pytest.log
for v in "" _v9 _v12 _a4v6 _a2v3; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python scripts/lbs_sweep.py 64 128 512 1024 2>&1 | tee -a $OUT/r02d_lbs_sweep.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 300 python scripts/lbs_sweep.py 64 512 2>&1 | tee -a $OUT/r02d_lbs_sweep.jsonl
echo "== dense weights bench"; timeout 300 python bench.py --steps 10 --warmup 3 --variant dense --no-cpu-baseline --no-extra-configs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dense: value',round(d['value']),'lbs frac',round(d['roofline']['frac'],3), d['stages']['lbs'])"
