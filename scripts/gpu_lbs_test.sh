#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-lbs}
echo "== lbs tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "lbs_kernels" 2>&1 | tail -30 | tee $OUT/${TAG}_lbs.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee $OUT/${TAG}_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee $OUT/${TAG}_bench.json
echo "== ncu full lbs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smpl_lbs_tc' -s 2 -c 2 -f -o $OUT/${TAG}_lbs_tc \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-200
