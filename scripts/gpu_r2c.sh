#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lbs or smpl or head" 2>&1 | tail -5 | tee $OUT/r02c_pytest.log
for v in "" _a4v9 _a2v12 _a3v9 _a3v6; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 300 python scripts/lbs_sweep.py 64 128 512 1024 2>&1 | tee -a $OUT/r02c_lbs_sweep.jsonl
done
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err; tail -c 600 $OUT/r02c_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'ceil',d['e2e']['d2h_ceiling_gbs'],'roof',round(d['roofline']['frac'],3))
    print({k:v['ms'] for k,v in d['stages'].items()})
    print(json.dumps(d['configs'])[:3000])
except Exception as e: print('bench parse failed',e)
PY
