#!/usr/bin/env python
"""Back-to-back timing of gait_linear on the head's GEMM shapes with prepared weights (GPU box only)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L

L.require_device()
shapes = [(1024, 6144, 2048), (1024, 1024, 2048), (1024, 1024, 1024), (1024, 20672, 224), (64, 6144, 2048)]
for (M, N, K) in shapes:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    for prep in (False, True):
        if prep:
            L.prepare_weight(W)
        run = lambda: L.call("gait_linear", A.data_ptr(), K, W.data_ptr(), K, None, None, 0, C.data_ptr(), N, M, N, K, L.stream_ptr())
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(5):
            e0.record()
            for _ in range(20):
                run()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        print(f"M={M:5d} N={N:6d} K={K:5d} prepared={int(prep)}  {best * 1e3:8.1f} us  {2 * M * N * K / best / 1e9:7.1f} TFLOP/s(fp32-equivalent)")
    L.release_weight(W)
