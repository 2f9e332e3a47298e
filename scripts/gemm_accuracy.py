#!/usr/bin/env python
"""Measure the error of gait_linear paths against FP64 (GPU box only)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from gaitb200 import _lib as L

def run(M, N, K, positive, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    if positive:
        A, W = A.abs(), W.abs()
    ref = (A.double().cuda() @ W.double().cuda().T)
    out = torch.empty(M, N, device="cuda")
    Ad, Wd = A.cuda(), W.cuda()
    L.call("gait_linear", Ad.data_ptr(), K, Wd.data_ptr(), K, None, None, 0, out.data_ptr(), N, M, N, K, L.stream_ptr())
    torch.cuda.synchronize()
    err = (out.double() - ref).abs()
    f32 = (Ad @ Wd.T).double()
    err32 = (f32 - ref).abs()
    scale = ref.abs().mean().item()
    return err.max().item() / scale, err.mean().item() / scale, err32.max().item() / scale, err32.mean().item() / scale

if __name__ == "__main__":
    L.require_device()
    torch.backends.cuda.matmul.allow_tf32 = False
    print("mode", os.environ.get("GAITB200_LINEAR", "auto"), os.environ.get("GAITB200_TC_MODE", "0"))
    print(f"{'M':>5} {'N':>6} {'K':>6} pos | gait max/mean rel err | torch fp32 max/mean rel err")
    for (M, N, K) in [(128, 128, 32), (128, 128, 256), (256, 256, 1024), (256, 256, 2048), (256, 512, 8192), (64, 6144, 2048)]:
        for pos in (False, True):
            a, b, c, d = run(M, N, K, pos)
            print(f"{M:5d} {N:6d} {K:6d} {int(pos)}   | {a:.3e} {b:.3e} | {c:.3e} {d:.3e}")
