#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and one full capture of the LBS kernel.
# Usage (from the repo root on the box):  bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.csv 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log
echo "== pytest -m gpu, uninitialised buffers poisoned with NaN" ; GAITB200_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_poison.log
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -5 | tee $OUT/${TAG}_bench.json
echo "== bench dense" ; timeout 900 python bench.py --steps 20 --warmup 5 --variant dense --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/${TAG}_bench_dense.json
echo "== ncu launch list"
GAITB200_GRU_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
echo "== ncu full (lbs, joint_regress)"
GAITB200_GRU_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'smpl_lbs|joint_regress' -s 2 -c 4 -f -o $OUT/${TAG}_lbs \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
