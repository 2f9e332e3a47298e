#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=video-based-gait-analysis-for-dementia_b200
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r02i_pytest.log
for v in "" _a3v3 _mb2 _nocons; do
  echo "== lib$v"; GAITB200_LIB=$PWD/$PKG/lib/libgaitb200$v.so timeout 120 python scripts/lbs_sweep.py 64 128 512 1024 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02i_lbs.jsonl
done
echo "== joints-only"; LBS_JOINTS_ONLY=1 timeout 120 python scripts/lbs_sweep.py 64 512 2>&1 | grep -E "lbs_us|Error" | tee -a $OUT/r02i_lbs.jsonl
echo "== jreg"; timeout 120 python scripts/jreg_time.py 1024 4096 2>&1 | tee -a $OUT/r02i_jreg.jsonl
JREG_ROWS=9 timeout 120 python scripts/jreg_time.py 1024 2>&1 | tee -a $OUT/r02i_jreg.jsonl
echo "== c4 (weight-stationary GRU)"; timeout 300 python scripts/gru_s_sweep.py 1 2 2>&1 | tail -2 | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02i_bench.json 2> $OUT/r02i_bench.err; tail -c 400 $OUT/r02i_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3))
    print({k:v['ms'] for k,v in d['stages'].items()})
    c=d['configs']; print('c3_n1',round(c['c3_n1']['value']), 'c4',{k:round(v['value']) for k,v in c['c4']['by_frames'].items()}, 'c5', round(c['c5']['value']), 'c1', c['c1'])
    print(c['c3_shard_sizes'])
except Exception as e: print('bench parse failed',e)
PY
