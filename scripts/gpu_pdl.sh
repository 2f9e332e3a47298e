#!/usr/bin/env bash
# A/B of programmatic dependent launch: GPU tests with it on, then the bench with GAITB200_PDL=1 and =0
set -u
TAG=${1:-r02za}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu (PDL on)"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_gpu.log
for P in 1 0; do
  echo "== bench PDL=$P"
  GAITB200_PDL=$P timeout 600 python bench.py --steps 50 --warmup 5 --no-extra-configs --no-cpu-baseline > $OUT/${TAG}_bench_pdl$P.json 2> $OUT/${TAG}_bench_pdl$P.err
  tail -c 300 $OUT/${TAG}_bench_pdl$P.err
  python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_pdl$P.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3), d['roofline'].get('us_per_launch'))
print({k:round(v['ms'],4) for k,v in d['stages'].items()})
PY
done
