#!/usr/bin/env python
"""GRU stage (input projection + recurrence) and whole-step device time vs the number of sequences S (T = 16).
Run once per variant; the variant comes from the environment (GAITB200_GRU_PATH / GAITB200_GRU_MAXCHUNKED)."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L, synthetic
from gaitb200.head import GaitHead

L.require_device()
head = GaitHead(synthetic.make_smpl_data(seed=0, variant="sparse"), synthetic.make_mean_params(),
                synthetic.make_regressor_state(seed=0), synthetic.make_gru_state(seed=0)).cuda()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 512, 1024]
tag = {k: os.environ[k] for k in ("GAITB200_GRU_PATH", "GAITB200_GRU_MAXCHUNKED") if k in os.environ}
for S in sizes:
    T = 16
    p = head.plan(S, T)
    p["x"].copy_(synthetic.make_features(S, T, seed=1).cuda())
    res = head.profile_stages(iters=5, flush=lambda: flush_buf.zero_())
    F = S * T
    tot = sum(v["ms"] for v in res.values())
    print(json.dumps({"variant": tag, "S": S, "F": F, "gru_ms": round(res["gru"]["ms"], 4), "gru_launches": res["gru"]["launches"],
                      "gru_Mframes_s": round(F / res["gru"]["ms"] / 1e3, 3),
                      "stages_us": {k: round(v["ms"] * 1e3, 1) for k, v in res.items()}, "sum_ms": round(tot, 4),
                      "lbs_gbs": round((F * 166512 + 661440) / res["lbs"]["ms"] * 1e-6)}), flush=True)
    head._slots = []; head._plan = None
    torch.cuda.empty_cache()
