"""TemporalEncoder: the GRU encoder over 2048-d backbone features that BASELINE.json's
north_star names (VIBE lib/models/vibe.py signature; the class itself is absent from the
reference tree - SURVEY.md fact 3).  The GRU kernel is general in input size, hidden size,
layers and direction so it also serves the reference's only in-tree GRU,
BidirectionalModel.rnn (lib/models/layers/gait_feat_encoder.py:51-57,88: 3072->300, 2 layers,
bidirectional).

Parameters live in a torch.nn.GRU / nn.Linear (same state_dict keys: gru.weight_ih_l0, ...);
the arithmetic runs in the sm_100a kernels behind gait_gru_layer / gait_linear.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


def gru_forward(gru: nn.GRU, x: torch.Tensor, resid: torch.Tensor | None = None, return_hn: bool = False):
    """nn.GRU forward (h0 = 0) on x laid out (S,T,I) contiguous.  Returns (y (S,T,H*dirs), out)
    where out = y + resid when `resid` (S,T,H) is given and the GRU is single-direction; with `return_hn` a third
    value h_n (num_layers*dirs, S, H) in torch's order (layer-major, forward before reverse)."""
    if gru.training and gru.dropout > 0 and gru.num_layers > 1:
        raise L.GaitLibraryError("GRU kernels implement eval() semantics (no inter-layer dropout); call .eval()")
    x = L.f32(x, "x")
    S, T, I = x.shape
    H, dirs = gru.hidden_size, 2 if gru.bidirectional else 1
    dev = x.device
    nbytes = L.load().gait_gru_workspace_bytes(S, T, H)
    ws = torch.empty(max(nbytes, 4) // 4, device=dev)
    inp, out = x, None
    hn = torch.empty(gru.num_layers * dirs, S, H, device=dev) if return_hn else None
    for layer in range(gru.num_layers):
        y = torch.empty(S, T, H * dirs, device=dev)
        last = layer == gru.num_layers - 1
        for d in range(dirs):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            w_ih_p = getattr(gru, "weight_ih" + sfx)
            if not gru.training and w_ih_p.is_contiguous():
                L.prepare_weight(w_ih_p)                      # input-projection weights: TF32 lo part split off once
            w_ih = L.f32(w_ih_p.detach(), "weight_ih")
            w_hh_p = getattr(gru, "weight_hh" + sfx)
            if not gru.training and w_hh_p.is_contiguous() and T > 1 and L.load().gait_gru_plan(S, T, H) == 0:
                # per-step path (many sequences): the recurrent GEMM runs T-1 times over W_hh, so its lo part is split off
                # once as well (measured at 1024 sequences: 174 -> 134 us per step); the persistent kernels read W_hh raw
                L.prepare_weight(w_hh_p)
            w_hh = L.f32(w_hh_p.detach(), "weight_hh")
            if gru.bias:
                b_ih = L.f32(getattr(gru, "bias_ih" + sfx).detach(), "bias_ih")
                b_hh = L.f32(getattr(gru, "bias_hh" + sfx).detach(), "bias_hh")
            else:
                b_ih = b_hh = torch.zeros(3 * H, device=dev)
            fuse_res = last and resid is not None and dirs == 1
            if fuse_res:
                out = torch.empty(S, T, H, device=dev)
            L.call("gait_gru_layer", L.ptr(inp), inp.shape[-1], L.ptr(w_ih), L.ptr(w_hh), L.ptr(b_ih), L.ptr(b_hh),
                   None, y.data_ptr() + 4 * d * H, H * dirs,
                   L.ptr(resid) if fuse_res else None, resid.shape[-1] if fuse_res else 0,
                   L.ptr(out) if fuse_res else None, H if fuse_res else 0,
                   None if hn is None else hn[layer * dirs + d].data_ptr(),
                   S, T, inp.shape[-1], H, d, L.ptr(ws), nbytes, L.stream_ptr())
        inp = y
    if return_hn:
        return inp, out, hn
    return inp, out


class TemporalEncoder(nn.Module):
    def __init__(self, n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False, use_residual=True,
                 input_size=2048):
        super().__init__()
        self.gru = nn.GRU(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional,
                          num_layers=n_layers)
        self.linear = None
        if bidirectional:
            self.linear = nn.Linear(hidden_size * 2, input_size)
        elif add_linear:
            self.linear = nn.Linear(hidden_size, input_size)
        self.use_residual = use_residual
        self.input_size = input_size

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        L.purge_prepared()                       # prepared-weight entries of GRU weights that moved (.to / .cuda / .float)
        return r

    @torch.no_grad()
    def forward(self, x):
        """x (N,T,F) -> (N,T,F') : GRU over T, optional relu+Linear, residual when F' == F."""
        x = L.f32(x, "x")
        n, t, f = x.shape
        if self.linear is None:
            want_res = self.use_residual and self.gru.hidden_size == self.input_size
            y, out = gru_forward(self.gru, x, resid=x if want_res else None)
            return out if want_res else y
        y, _ = gru_forward(self.gru, x)
        st = L.stream_ptr()
        L.call("gait_relu", L.ptr(y), L.ptr(y), y.numel(), st)
        w = L.f32(self.linear.weight.detach(), "linear.weight")
        b = L.f32(self.linear.bias.detach(), "linear.bias")
        out = torch.empty(n, t, w.shape[0], device=x.device)
        res = x if (self.use_residual and w.shape[0] == self.input_size) else None
        L.call("gait_linear", L.ptr(y), y.shape[-1], L.ptr(w), w.shape[1], L.ptr(b), L.ptr(res),
               0 if res is None else f, L.ptr(out), w.shape[0], n * t, w.shape[0], y.shape[-1], st)
        return out
