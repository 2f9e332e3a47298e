#!/usr/bin/env bash
# final check of a round: tests (plain + poisoned), smoke, bench at N=1 with every sub-record, launch list
set -u
TAG=${1:-r02z}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_gpu.log
echo "== pytest poisoned"; GAITB200_TEST_POISON=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $OUT/${TAG}_pytest_gpu_poison.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 300 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'ceil',round(d['e2e']['d2h_ceiling_gbs'],1),'roof',round(d['roofline']['frac'],3))
print({k:round(v['ms'],4) for k,v in d['stages'].items()})
c=d['configs']; print('c3_n1',round(c['c3_n1']['value']),'c4',{k:round(v['value']) for k,v in c['c4']['by_frames'].items()},'c5',{k:round(v['value']) for k,v in c['c5']['modes'].items()}, 'c1', round(c['c1']['gpu']['value']), round(c['c1']['cpu']['value']))
print('cpu', d['cpu_baseline'])
PY
echo "== reference arm N=1"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
