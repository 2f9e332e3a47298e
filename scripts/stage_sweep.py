#!/usr/bin/env python
"""Per-stage device times of the head at several batch sizes (fixed launch overheads vs streaming rates)."""
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
sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 512]
for S in sizes:
    T = 16
    p = head.plan(S, T)
    p["x"].copy_(synthetic.make_features(S, T, seed=1).cuda())
    res = head.profile_stages(iters=10, flush=lambda: flush_buf.zero_())
    F = S * T
    lbs_bytes = F * 166512 + 661440
    print(f"S={S:4d} F={F:5d}: " + "  ".join(f"{k} {v['ms'] * 1e3:7.1f}us" for k, v in res.items())
          + f" | lbs {lbs_bytes / res['lbs']['ms'] * 1e-6:6.0f} GB/s")
