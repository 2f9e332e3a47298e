#!/usr/bin/env bash
# round-2 evidence run: tests (plain + NaN-poisoned buffers), smoke (plain and under ncu), racecheck on the kernels with
# hand-rolled barrier protocols, rotation / vertex error per shard size, bench, ncu launch list + full captures
set -u
TAG=${1:-r02p}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_gpu.log
echo "== pytest poisoned"; GAITB200_TEST_POISON=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_gpu_poison.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9 | tee $OUT/${TAG}_smoke.log
echo "== smoke under ncu (launch list)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_smoke_ncu.log | cut -c1-200
echo "== shard errors"; timeout 600 python scripts/shard_err.py 2>&1 | tail -12 | tee $OUT/${TAG}_shard_err.txt
echo "== racecheck"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -x -q \
    -k "lbs_kernels_vs_oracle or joint_regress_stream or three_joint_chain or head_vs_oracle" 2>&1 | tail -12 | cut -c1-250 | tee $OUT/${TAG}_racecheck.txt
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 300 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3))
print({k:round(v['ms'],4) for k,v in d['stages'].items()})
c=d['configs']; print('c3_n1',round(c['c3_n1']['value']),'c4',{k:round(v['value']) for k,v in c['c4']['by_frames'].items()},'c5',{k:round(v['value']) for k,v in c['c5']['modes'].items()})
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > $OUT/${TAG}_ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full: lbs, jreg"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smpl_lbs_tc|joint_regress_stream' -s 2 -c 3 -f -o $OUT/${TAG}_lbs_jreg \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > $OUT/${TAG}_ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full: weight-stationary GRU (S=1, T=64)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gru_small' -c 1 -f -o $OUT/${TAG}_gru_small \
    python scripts/gru_s_sweep.py 1 > $OUT/${TAG}_ncu_gru_small.log 2>&1; echo "rc=$?"
ls -la $OUT | grep ${TAG}
