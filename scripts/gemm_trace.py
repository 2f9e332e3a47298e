#!/usr/bin/env python
"""Pipeline timing of the tensor-core GEMM (CTA 0): where does a k-block's time go?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L

L.require_device()
for (M, N, K) in [(1024, 6144, 2048), (64, 6144, 2048), (1024, 1024, 1024)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    if "--prepared" in sys.argv:
        L.prepare_weight(W)
    tr = torch.zeros(64 * 4, dtype=torch.int64, device="cuda")
    for rep in range(3):
        tr.zero_()
        L.call("gait_debug_linear_trace", tr.data_ptr())
        L.call("gait_linear", A.data_ptr(), K, W.data_ptr(), K, None, None, 0, C.data_ptr(), N, M, N, K, L.stream_ptr())
        torch.cuda.synchronize()
    L.call("gait_debug_linear_trace", None)
    t = tr.cpu().view(64, 4)
    t0 = int(t[0, 0])
    print(f"--- M={M} N={N} K={K}: kb: stage_free  landed(+L)  converted(+C)  mma_issued | period")
    prev = None
    for kb in range(min(K // 32, 40)):
        a, b, c, d = [int(x) - t0 for x in t[kb]]
        per = "" if prev is None else a - prev
        print(f"{kb:3d} {a:8d} {b - a:8d} {c - b:8d} {d - c:8d}   {per}")
        prev = a
