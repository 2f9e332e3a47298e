#!/usr/bin/env bash
set -u
TAG=${1:-r02q}; OUT=gpurun_out; mkdir -p $OUT
echo "== gru tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "gru or temporal or long_clip or head_vs_oracle" 2>&1 | tail -3
echo "== c4"; timeout 300 python scripts/gru_s_sweep.py 1 2 2>&1 | tail -2 | cut -c1-200
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, json
from gaitb200 import synthetic
from gaitb200.head import GaitHead
head = GaitHead(synthetic.make_smpl_data(seed=0), synthetic.make_mean_params(), synthetic.make_regressor_state(seed=0), synthetic.make_gru_state(seed=0)).cuda()
for T in (16, 900):
    head.capture(1, T); head.input.copy_(synthetic.make_features(1, T, seed=1))
    for _ in range(3): head.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): head.step()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"T": T, "ms_per_step": round(ms, 4), "frames_per_s": round(T / ms * 1e3)}))
PY
echo "== racecheck (kernels without inter-CTA spin flags)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "lbs_kernels_vs_oracle or joint_regress_stream or three_joint_chain" > $OUT/${TAG}_racecheck_full.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" $OUT/${TAG}_racecheck_full.txt | tail -3
grep -E "Race reported|hazard|at .*kernel|in .*\.cu" $OUT/${TAG}_racecheck_full.txt | head -24 | cut -c1-260 | tee $OUT/${TAG}_racecheck.txt
echo "== ncu full: jreg"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'joint_regress_stream' -s 2 -c 2 -f -o $OUT/${TAG}_jreg \
    python scripts/jreg_time.py 1024 > $OUT/${TAG}_ncu_jreg.log 2>&1; echo "rc=$?"
