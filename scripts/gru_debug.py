#!/usr/bin/env python
"""Where does a GRU mismatch start? per-step error of gru_forward against torch.nn.GRU (GPU box only)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200.temporal import gru_forward

S, T, H = (int(a) for a in (sys.argv[1:4] + ["1", "16", "2048"][len(sys.argv) - 1:]))
torch.manual_seed(3)
gru = torch.nn.GRU(H, H).eval()
x = torch.randn(S, T, H) * 0.5
with torch.no_grad():
    ref, _ = gru(x.permute(1, 0, 2))
ref = ref.permute(1, 0, 2)
gru = gru.cuda()
for rep in range(3):
    y, _ = gru_forward(gru, x.cuda())
    e = (y.cpu() - ref).abs()
    print(f"rep {rep}: max err {e.max():.3e}; per step:", " ".join(f"{e[:, t].max():.1e}" for t in range(T)))
    bad = (e > 2e-5).nonzero()
    if len(bad):
        t0 = int(bad[:, 1].min())
        units = bad[bad[:, 1] == t0][:, 2]
        print(f"   first bad step {t0}: {len(units)} units, clusters(32) {sorted(set((units // 32).tolist()))[:20]}, seqs {sorted(set(bad[bad[:, 1] == t0][:, 0].tolist()))[:10]}")
