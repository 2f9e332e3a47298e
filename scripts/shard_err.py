#!/usr/bin/env python
"""Head output error against the FP32 oracle and against the same oracle run in FP64 (GPU box; test infrastructure).
Shows how much of the distance to the FP32 oracle is the oracle's own rounding."""
import sys, copy
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from gaitb200 import synthetic
from gaitb200.head import GaitHead
from oracle.head import GaitHeadOracle

def main():
    data = synthetic.make_smpl_data(seed=0, variant="sparse")
    mean = synthetic.make_mean_params()
    rs = synthetic.make_regressor_state(seed=0, decoder_gain=0.3)
    gs = synthetic.make_gru_state(seed=0)
    head = GaitHead(data, mean, rs, gs).cuda()
    o32 = GaitHeadOracle(data, mean, rs, gs)
    o64 = copy.deepcopy(o32).double()
    for S in [int(a) for a in sys.argv[1:]] or [64, 100, 200]:
        feats = synthetic.make_features(S, 16, seed=9)
        out = head(feats.cuda())
        r32 = o32(feats)
        torch.set_default_dtype(torch.float64)          # the oracle creates a few constants (eye, zeros) in the default dtype
        r64 = o64(feats.double())
        torch.set_default_dtype(torch.float32)
        for k in ("rotmat", "verts", "kp_3d", "theta"):
            a = out[k].cpu().double()
            e1 = (a - r32[k].double()).abs().max().item()
            e2 = (a - r64[k]).abs().max().item()
            e3 = (r32[k].double() - r64[k]).abs().max().item()
            print(f"S={S:4d} {k:8s} ours-vs-f32 {e1:.3e}  ours-vs-f64 {e2:.3e}  f32-vs-f64 {e3:.3e}")

def stages(S=64):
    """Per-stage error with exact (FP64-computed, FP32-rounded) inputs: which stage contributes the distance."""
    from oracle import geometry as OG
    data = synthetic.make_smpl_data(seed=0, variant="sparse")
    mean = synthetic.make_mean_params()
    rs = synthetic.make_regressor_state(seed=0, decoder_gain=0.3)
    gs = synthetic.make_gru_state(seed=0)
    head = GaitHead(data, mean, rs, gs).cuda()
    o32 = GaitHeadOracle(data, mean, rs, gs)
    o64 = copy.deepcopy(o32).double()
    feats = synthetic.make_features(S, 16, seed=9)
    def rep(name, a, r32, r64):
        a, r32 = a.cpu().double(), r32.double()
        print(f"{name:10s} ours-vs-f64 max {(a - r64).abs().max():.3e} rms {(a - r64).pow(2).mean().sqrt():.3e} | "
              f"f32-vs-f64 max {(r32 - r64).abs().max():.3e} rms {(r32 - r64).pow(2).mean().sqrt():.3e} | scale {r64.abs().mean():.3e}")
    torch.set_default_dtype(torch.float64)
    y64 = o64.encoder(feats.double())
    torch.set_default_dtype(torch.float32)
    y32 = o32.encoder(feats)
    y = head.encoder(feats.cuda())
    rep("encoder", y, y32, y64)
    yin = y64.float().reshape(S * 16, -1)
    torch.set_default_dtype(torch.float64)
    st64 = torch.cat(o64.regressor.iterate(yin.double()), 1)
    torch.set_default_dtype(torch.float32)
    st32 = torch.cat(o32.regressor.iterate(yin), 1)
    st = head.regressor.iterate(yin.cuda())[:, :157]
    rep("regressor", st, st32, st64)
    p6 = st64[:, :144].float()
    from gaitb200 import geometry as GG
    R = GG.rot6d_to_rotmat(p6.cuda())
    rep("rot6d", R.reshape(-1, 9), OG.rot6d_to_rotmat(p6).reshape(-1, 9), OG.rot6d_to_rotmat(p6.double()).reshape(-1, 9))


if __name__ == "__main__":
    if sys.argv[1:2] == ["stages"]:
        stages(int(sys.argv[2]) if len(sys.argv) > 2 else 64)
        sys.exit(0)
    main()
