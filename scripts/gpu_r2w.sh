#!/usr/bin/env bash
set -u
echo "== gru tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "gru or temporal or long_clip or head_vs_oracle or repeatable" 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, json
from gaitb200 import synthetic
from gaitb200.head import GaitHead
head = GaitHead(synthetic.make_smpl_data(seed=0), synthetic.make_mean_params(), synthetic.make_regressor_state(seed=0), synthetic.make_gru_state(seed=0)).cuda()
for S, T in ((1, 16), (1, 64), (1, 900), (2, 16), (2, 450)):
    head.capture(S, T); head.input.copy_(synthetic.make_features(S, T, seed=1))
    for _ in range(3): head.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): head.step()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    st = head.profile_stages(iters=3)
    print(json.dumps({"S": S, "T": T, "ms_per_step": round(ms, 4), "frames_per_s": round(S * T / ms * 1e3), "gru_ms": round(st["gru"]["ms"], 4),
                      "us_per_recurrence_step": round(st["gru"]["ms"] * 1e3 / T, 2)}))
PY
