#!/usr/bin/env python
"""Determinism stress of the persistent GRU path: the kernels are deterministic, so any bitwise difference between
repetitions is a race.  Perturbs timing / cache state between repetitions (GPU box only)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200.temporal import gru_forward

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
torch.manual_seed(3)
H = 2048
gru = torch.nn.GRU(H, H).eval().cuda()
junk = torch.randn(64 << 20, device="cuda")
for (S, T) in [(1, 16), (3, 5), (8, 16), (64, 16)]:
    x = (torch.randn(S, T, H) * 0.5).cuda()
    base = gru_forward(gru, x)[0].clone()
    bad = 0
    for r in range(reps):
        if r % 3 == 0:
            junk.mul_(1.0001)                      # evict L2, change what is resident
        elif r % 3 == 1:
            torch.cuda.synchronize()
        y = gru_forward(gru, x)[0]
        if not torch.equal(y, base):
            d = (y - base).abs()
            idx = (d > 0).nonzero()
            t0 = int(idx[:, 1].min())
            units = idx[idx[:, 1] == t0][:, 2]
            if bad < 5:
                print(f"  S={S} T={T} rep {r}: max diff {float(d.max()):.3e}, first differing step {t0}, {len(units)} units there, "
                      f"k-blocks(32) {sorted(set((units // 32).tolist()))[:12]}, seqs {sorted(set(idx[idx[:, 1] == t0][:, 0].tolist()))[:8]}")
            bad += 1
    print(f"S={S} T={T}: {bad} of {reps} repetitions differ from the first")
