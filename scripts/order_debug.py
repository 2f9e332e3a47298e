#!/usr/bin/env python
"""Reproduce an order-dependent mismatch: GRU tests first, then the head; report per-stage errors (GPU box only)."""
import sys, gc
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from gaitb200 import synthetic, _lib as L
from gaitb200.temporal import gru_forward
from gaitb200.head import GaitHead
from oracle.head import GaitHeadOracle

cfgs = [dict(S=2, T=5, I=32, H=32, layers=1, bi=False), dict(S=3, T=7, I=48, H=20, layers=2, bi=True),
        dict(S=4, T=6, I=3072, H=300, layers=2, bi=True), dict(S=1, T=16, I=2048, H=2048, layers=1, bi=False),
        dict(S=64, T=16, I=2048, H=2048, layers=1, bi=False), dict(S=5, T=9, I=96, H=256, layers=2, bi=True),
        dict(S=64, T=4, I=40, H=64, layers=1, bi=False), dict(S=37, T=1, I=64, H=128, layers=1, bi=False),
        dict(S=70, T=5, I=64, H=128, layers=1, bi=False), dict(S=200, T=3, I=64, H=128, layers=1, bi=False)]
which = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 and sys.argv[1] else []

def run_gru(cfg):
    torch.manual_seed(3)
    gru = torch.nn.GRU(cfg["I"], cfg["H"], num_layers=cfg["layers"], bidirectional=cfg["bi"]).eval()
    x = torch.randn(cfg["S"], cfg["T"], cfg["I"]) * 0.5
    with torch.no_grad():
        ref, _ = gru(x.permute(1, 0, 2))
    y, _ = gru_forward(gru.cuda(), x.cuda())
    return float((y.cpu() - ref.permute(1, 0, 2)).abs().max())

for i in which:
    print("gru cfg", i, "err", run_gru(cfgs[i]), "prepared entries", len(L._prepared))
data = synthetic.make_smpl_data(seed=0, variant="sparse")
mean = synthetic.make_mean_params()
rs = synthetic.make_regressor_state(seed=0, decoder_gain=0.3)
gs = synthetic.make_gru_state(seed=0)
head = GaitHead(data, mean, rs, gs).cuda()
oracle = GaitHeadOracle(data, mean, rs, gs)
feats = synthetic.make_features(3, 5, seed=1234)
ref = oracle(feats)
with torch.no_grad():
    yref = oracle.encoder(feats)
    sref = torch.cat(oracle.regressor.iterate(yref.reshape(15, -1)), 1)
for rep in range(2):
    out = head(feats.cuda())
    y = head.encoder(feats.cuda())
    st = head.regressor.iterate(y.reshape(15, -1))
    st2 = head.regressor.iterate(yref.reshape(15, -1).cuda())
    print(f"rep {rep}: rotmat {float((out['rotmat'].cpu() - ref['rotmat']).abs().max()):.3e}  encoder {float((y.cpu() - yref).abs().max()):.3e}  "
          f"state {float((st[:, :157].cpu() - sref).abs().max()):.3e}  state(exact y) {float((st2[:, :157].cpu() - sref).abs().max()):.3e}  prepared {len(L._prepared)}")
