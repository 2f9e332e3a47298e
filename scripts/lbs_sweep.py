#!/usr/bin/env python
"""LBS kernel rate vs frames per launch (back-to-back launches on two alternating buffer sets + single-launch event pairs).
GAITB200_LIB selects the library build; prints one JSON line per size."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L, synthetic
from gaitb200.head import GaitHead

L.require_device()
jo = os.environ.get("LBS_JOINTS_ONLY") == "1"
head = GaitHead(synthetic.make_smpl_data(seed=0, variant="sparse"), synthetic.make_mean_params(),
                synthetic.make_regressor_state(seed=0), synthetic.make_gru_state(seed=0), write_mesh=not jo, joints_mode="skin").cuda()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512, 1024]
peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else 6548.5
for S in sizes:
    T = 16
    head.plan(S, T, slots=2)
    for p in head._slots:
        p["x"].copy_(synthetic.make_features(S, T, seed=1).cuda())
    F = S * T
    b2b = head.time_stage_back_to_back("lbs", launches=8, repeats=5)
    res = head.profile_stages(iters=5, flush=lambda: flush_buf.zero_())
    nbytes = F * (166512 - (6890 * 12 - 21 * 12 if jo else 0)) + 661440
    print(json.dumps({"lib": os.path.basename(os.environ.get("GAITB200_LIB", "default")), "F": F, "lbs_us_b2b": round(b2b * 1e3, 2),
                      "gbs_b2b": round(nbytes / b2b * 1e-6), "frac_b2b": round(nbytes / b2b * 1e-6 / peak, 4),
                      "lbs_us_single": round(res["lbs"]["ms"] * 1e3, 2), "frac_single": round(nbytes / res["lbs"]["ms"] * 1e-6 / peak, 4),
                      "blend_us": round(res["blend"]["ms"] * 1e3, 1)}), flush=True)
    head._slots = []; head._plan = None
    torch.cuda.empty_cache()
