#!/usr/bin/env python
"""Does programmatic dependent launch survive stream capture?  Captures the regressor stage (13 launches), the SMPL part
(chain, blend, skinning, joints) and the whole step into CUDA graphs, times replays, and dumps the graph of the regressor so
that the edge types can be read.  Run once with GAITB200_PDL=1 and once with =0."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L, synthetic
from gaitb200.head import GaitHead

L.require_device()
head = GaitHead(synthetic.make_smpl_data(seed=0, variant="sparse"), synthetic.make_mean_params(),
                synthetic.make_regressor_state(seed=0), synthetic.make_gru_state(seed=0)).cuda()
S, T = 64, 16
head.plan(S, T, slots=1)
p = head._plan
p["x"].copy_(synthetic.make_features(S, T, seed=1).cuda())
head._launch(p)
torch.cuda.synchronize()
stages = dict(head._stages(p))


def graph_of(fns, reps, dump=None):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    if dump:
        g.enable_debug_mode()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    if dump:
        g.debug_dump(dump)
    return g


def time_graph(g, reps, n=30):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            g.replay()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) / n / reps)
    return best


out = {"pdl": os.environ.get("GAITB200_PDL", "1")}
os.makedirs("gpurun_out", exist_ok=True)
if os.environ.get("PDL_CHECK_QUICK") == "1":
    g = graph_of([stages[n] for n in ("pose_chain", "blend", "lbs", "joints")], 4)
    out["smpl_part_us_x4"] = round(time_graph(g, 4) * 1e3, 2)
    g = graph_of([f for _, f in head._stages(p)], 1)
    out["step_us"] = [round(time_graph(g, 1, n=60) * 1e3, 2) for _ in range(3)]
    print(json.dumps(out), flush=True)
    sys.exit(0)
g = graph_of([stages["regressor"]], 8)
out["regressor_us_x8"] = round(time_graph(g, 8) * 1e3, 2)
g = graph_of([stages[n] for n in ("pose_chain", "blend", "lbs", "joints")], 4)
out["smpl_part_us_x4"] = round(time_graph(g, 4) * 1e3, 2)
g = graph_of([stages[n] for n in ("pose_chain", "blend")], 4)
out["chain_blend_us_x4"] = round(time_graph(g, 4) * 1e3, 2)
g = graph_of([stages[n] for n in ("blend", "lbs")], 4)
out["blend_lbs_us_x4"] = round(time_graph(g, 4) * 1e3, 2)
g = graph_of([stages[n] for n in ("lbs", "joints")], 4)
out["lbs_joints_us_x4"] = round(time_graph(g, 4) * 1e3, 2)
g = graph_of([stages["blend"]], 8)
out["blend_us_x8"] = round(time_graph(g, 8) * 1e3, 2)
g = graph_of([stages["lbs"]], 8)
out["lbs_us_x8_same_buffers"] = round(time_graph(g, 8) * 1e3, 2)
g = graph_of([f for _, f in head._stages(p)], 1)
out["step_us"] = round(time_graph(g, 1) * 1e3, 2)
print(json.dumps(out), flush=True)
