#!/usr/bin/env bash
# evidence for the late round-2 kernels: smoke under ncu, ncu launch list of a bench run, full captures of the recurrence
# kernel (tcgen05.cp + setmaxnreg), the skinning kernel (register split) and the joint regressor; racecheck on the latter two
set -u
TAG=${1:-r02zc}; OUT=gpurun_out; mkdir -p $OUT
echo "== smoke under ncu (launch list)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_smoke_ncu.log | cut -c1-200
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > $OUT/${TAG}_ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full: recurrence"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gru_recurrent' -s 3 -c 1 -f -o $OUT/${TAG}_gru \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > $OUT/${TAG}_gru.log 2>&1; echo "rc=$?"
echo "== ncu full: lbs, jreg"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smpl_lbs_tc|joint_regress_stream' -s 2 -c 3 -f -o $OUT/${TAG}_lbs_jreg \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > $OUT/${TAG}_ncu_full.log 2>&1; echo "rc=$?"
echo "== racecheck"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -x -q \
    -k "lbs_kernels_vs_oracle or joint_regress_stream or three_joint_chain" 2>&1 | tail -8 | cut -c1-250 | tee $OUT/${TAG}_racecheck.txt
ls -la $OUT | grep ${TAG}
