#!/usr/bin/env python
"""Timeline of the persistent recurrent GRU kernel (CTA 0): per-step phases and the k-block pipeline of step 2."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gaitb200 import _lib as L

L.require_device()
S, T, H = 64, 16, 2048
torch.manual_seed(0)
gru = torch.nn.GRU(H, H).cuda().eval()
x = torch.randn(S, T, H, device="cuda") * 0.5
y = torch.empty(S, T, H, device="cuda"); out = torch.empty_like(y)
nbytes = L.load().gait_gru_workspace_bytes(S, T, H)
ws = torch.empty(nbytes // 4 + 1, device="cuda")
tr = torch.zeros(1280, dtype=torch.int64, device="cuda")
w = {k: v.detach() for k, v in gru.named_parameters()}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(4):
    tr.zero_()
    L.call("gait_debug_gru_trace", tr.data_ptr())
    ev[0].record()
    L.call("gait_gru_layer", x.data_ptr(), H, w["weight_ih_l0"].data_ptr(), w["weight_hh_l0"].data_ptr(),
           w["bias_ih_l0"].data_ptr(), w["bias_hh_l0"].data_ptr(), None, y.data_ptr(), H, x.data_ptr(), H, out.data_ptr(), H,
           None, S, T, H, H, 0, ws.data_ptr(), nbytes, L.stream_ptr())
    ev[1].record()
    torch.cuda.synchronize()
    print(f"rep {rep}: gru layer (input projection + recurrence) {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
L.call("gait_debug_gru_trace", None)
t = tr.cpu()
st = t[:256].view(32, 8)
t0 = int(st[0, 5])
print("step: bar_wait_start bar_passed first_h_landed last_mma_issued drained P_ready gates_done stored | step period (cycles)")
prev = None
for s in range(T):
    v = [int(st[s, i]) - t0 if int(st[s, i]) else 0 for i in (6, 0, 1, 2, 3, 4, 7, 5)]
    per = "" if prev is None else v[7] - prev
    print(f"{s:3d} " + " ".join(f"{a:9d}" for a in v) + f"   {per}")
    prev = v[7]
kb = t[256:768].view(64, 8)
k0 = int(kb[0, 0])
print("step 2 k-blocks (cycles; + = relative to W issue): W_issue | h_issue+ convW_start+ convH_start+ conv_done(warp12)+ mma:acc_free+ mma:conv_seen+ mma_issued+ | period")
prev = None
for i in range(32):
    v = [int(x) - k0 for x in kb[i]]
    a = v[0]
    print(f"{i:3d} {a:8d} | " + " ".join(f"{v[j] - a:8d}" for j in (1, 4, 5, 6, 2, 3, 7)) + f"   {'' if prev is None else a - prev}")
    prev = a

sk = t[1024:1280].view(128, 2)
st0 = int(sk[:, 0].min())
stored = sorted(int(x) - st0 for x in sk[:, 0])
passed = sorted(int(x) - st0 for x in sk[:, 1])
print("all CTAs, globaltimer ns from the first CTA's store of h_3: stored min/median/max", stored[0], stored[64], stored[-1],
      "| flags of step 4 passed min/median/max", passed[0], passed[64], passed[-1])
print("stored per CTA (ns):", [int(x) - st0 for x in sk[:, 0]])
