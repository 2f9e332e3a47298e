#!/usr/bin/env bash
# quick iteration: selected tests + bench + optional ncu of one kernel regex
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}; KSEL=${2:-lbs_kernels}; NCU=${3:-}
echo "== tests ($KSEL)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$KSEL" 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('frames/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'lbs GB/s', round(d['roofline']['achieved']), 'frac', round(d['roofline']['frac'],3), 'single-launch frac', round(d['roofline'].get('frac_single_launch_event_pair', 0),3))
print({k:(v['ms'],v['launches']) for k,v in d['stages'].items()})"
if [ -n "$NCU" ]; then
  echo "== ncu full $NCU"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s 2 -c 2 -f -o $OUT/${TAG}_ncu \
      python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
  tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
fi
if [ -n "${LAUNCHLIST:-}" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
fi
