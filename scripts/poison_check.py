#!/usr/bin/env python
"""Fill every torch.empty() CUDA float buffer with NaN and report which head outputs / plan buffers contain NaN afterwards."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
real_empty = torch.empty
def empty(*a, **k):
    t = real_empty(*a, **k)
    if t.is_cuda and t.is_floating_point() and t.numel():
        t.fill_(float("nan"))
    return t
torch.empty = empty
from gaitb200 import synthetic
from gaitb200.head import GaitHead

data = synthetic.make_smpl_data(seed=0, variant="sparse")
head = GaitHead(data, synthetic.make_mean_params(), synthetic.make_regressor_state(seed=0, decoder_gain=0.3),
                synthetic.make_gru_state(seed=0), write_mesh="--joints-only" not in sys.argv).cuda()
S, T = 3, 5
out = head(synthetic.make_features(S, T, seed=1234).cuda())
for k, v in out.items():
    n = int(torch.isnan(v).sum())
    print(f"out {k:10s} shape {tuple(v.shape)} nan {n}" + (f" first idx {torch.isnan(v).nonzero()[0].tolist()}" if n else ""))
for k, v in head._plan.items():
    if torch.is_tensor(v) and v.is_floating_point():
        n = int(torch.isnan(v).sum())
        if n:
            idx = torch.isnan(v).nonzero()
            print(f"plan {k:10s} shape {tuple(v.shape)} nan {n} first {idx[0].tolist()} last {idx[-1].tolist()}")
