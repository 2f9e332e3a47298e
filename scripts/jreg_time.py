#!/usr/bin/env python
"""Time the streaming joint-regressor kernel (17 rows) at several frame counts; GAITB200_LIB selects the build."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L, synthetic
L.require_device()
lib = L.load()
V, rows = 6890, int(os.environ.get("JREG_ROWS", "17"))
variant = os.environ.get("JREG_VARIANT", "sparse")
jr = torch.as_tensor(synthetic.make_smpl_data(seed=0, variant=variant)["J_regressor_h36m"]).cuda()[:rows].contiguous()
packed = torch.empty(lib.gait_joint_regress_pack_bytes(V, rows) // 4, device="cuda")
L.call("gait_joint_regress_pack", jr.data_ptr(), packed.data_ptr(), V, rows, L.stream_ptr())
for F in [int(a) for a in sys.argv[1:]] or [1024, 4096]:
    bufs = [torch.randn(F, V, 3, device="cuda") for _ in range(2 if F <= 4096 else 1)]
    out = torch.empty(F, rows, 3, device="cuda")
    run = lambda b: L.call("gait_joint_regress_packed", b.data_ptr(), packed.data_ptr(), out.data_ptr(), F, V, rows, L.stream_ptr())
    for b in bufs: run(b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(8): run(bufs[i % len(bufs)])
        e.record(); e.synchronize()
        best = min(best, a.elapsed_time(e) / 8)
    # the same 8 launches replayed from a CUDA graph (no host launch cost between kernels)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(8): run(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    bestg = 1e9
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); e.synchronize()
        bestg = min(bestg, a.elapsed_time(e) / 8)
    nbytes = F * (82680 + 12 * rows) + 27560 * rows
    print(json.dumps({"lib": os.path.basename(os.environ.get("GAITB200_LIB", "default")), "variant": variant, "F": F, "rows": rows, "us": round(best * 1e3, 2), "us_graph": round(bestg * 1e3, 2), "frac_graph": round(nbytes / bestg * 1e-6 / 6548.5, 4),
                      "gbs": round(nbytes / best * 1e-6), "frac": round(nbytes / best * 1e-6 / 6548.5, 4)}), flush=True)
