"""Layers of lib/models/layers that sit directly next to the regression path, same class names,
constructor arguments and state_dict keys as the reference; the arithmetic runs in csrc/heads.cu.

  * LocallyConnected2d   lib/models/layers/locallyconnected2d.py:22-49  (kernel_size = 1, output_size = [J, 1])
  * KeypointAttention    lib/models/layers/keypoint_attention.py:22-55  (use_conv = False, act = 'softmax')
  * BidirectionalModel   lib/models/layers/gait_feat_encoder.py:10-104  (use_pareFeat = True - the branch that runs as shipped)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .temporal import gru_forward


def _linear(x, lin: nn.Linear, out=None, ldc=None):
    """torch.nn.Linear on (M,K) rows through gait_linear; `out` may be a column slice of a wider buffer (ldc)."""
    x = L.f32(x, "x")
    w, b = L.f32(lin.weight.detach(), "weight"), L.f32(lin.bias.detach(), "bias")
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device)
        ldc = N
    L.call("gait_linear", L.ptr(x), K, L.ptr(w), K, L.ptr(b), None, 0, out.data_ptr(), ldc, M, N, K, L.stream_ptr())
    return out


def _act(x, kind, slope=0.0):
    L.call("gait_activation", L.ptr(x), L.ptr(x), x.numel(), kind, float(slope), L.stream_ptr())
    return x


class LocallyConnected2d(nn.Module):
    """locallyconnected2d.py:22-49.  Only what the reference instantiates is implemented: kernel_size = 1, stride = 1,
    output_size = [J, 1] (pare.py:422-430, gait_feat_encoder.py:43-49); anything else raises."""

    def __init__(self, in_channels, out_channels, output_size, kernel_size, stride, bias=False):
        super().__init__()
        output_size = tuple(output_size) if isinstance(output_size, (list, tuple)) else (output_size, output_size)
        if kernel_size != 1 or stride != 1 or output_size[1] != 1:
            raise NotImplementedError("gaitb200.LocallyConnected2d implements kernel_size=1, stride=1, output_size=[J,1]")
        self.weight = nn.Parameter(torch.randn(1, out_channels, in_channels, output_size[0], output_size[1], kernel_size ** 2))
        if bias:
            self.bias = nn.Parameter(torch.randn(1, out_channels, output_size[0], output_size[1]))
        else:
            self.register_parameter('bias', None)
        self.kernel_size, self.stride = (kernel_size, kernel_size), (stride, stride)

    def run(self, x, resid=None, out_layout="NOJ"):
        """x (N,C,J,1) or any expanded view of it -> (N,O,J,1) [out_layout 'NOJ'] or (N,J,O) ['NJO'].
        With `resid` (same layout as the output) also returns out + resid."""
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32):
            raise L.GaitLibraryError("LocallyConnected2d: expected a float32 CUDA tensor (no CPU path)")
        _, O, C, J, _, _ = self.weight.shape
        if x.dim() != 4 or x.shape[1] != C or x.shape[2] != J or x.shape[3] != 1:
            raise ValueError(f"expected (N,{C},{J},1), got {tuple(x.shape)}")
        N = x.shape[0]
        w = L.f32(self.weight.detach(), "weight")
        b = None if self.bias is None else L.f32(self.bias.detach(), "bias")
        if out_layout == "NOJ":
            out = torch.empty(N, O, J, 1, device=x.device)
            so = (O * J, J, 1)
        else:
            out = torch.empty(N, J, O, device=x.device)
            so = (J * O, 1, O)
        out2 = None
        if resid is not None:
            resid = L.f32(resid, "resid")
            out2 = torch.empty_like(out)
        L.call("gait_locally_connected", x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), L.ptr(w), C * J, J, 1,
               L.ptr(b), J, 1, L.ptr(out), so[0], so[1], so[2], L.ptr(resid), L.ptr(out2), N, C, O, J, L.stream_ptr())
        return out if resid is None else (out, out2)

    @torch.no_grad()
    def forward(self, x):
        return self.run(x)


class KeypointAttention(nn.Module):
    """keypoint_attention.py:22-55 with use_conv=False, act='softmax' (how pare.py:237-243 builds it)."""

    def __init__(self, use_conv=False, in_channels=(256, 64), out_channels=(256, 64), act='softmax', use_scale=False):
        super().__init__()
        if use_conv or act != 'softmax':
            raise NotImplementedError("gaitb200.KeypointAttention implements use_conv=False, act='softmax' (pare.py:237-243)")
        self.use_conv, self.in_channels, self.out_channels, self.act, self.use_scale = use_conv, in_channels, out_channels, act, use_scale

    @torch.no_grad()
    def forward(self, features, heatmaps):
        f, h = L.f32(features, "features"), L.f32(heatmaps, "heatmaps")
        B, J, Hh, Ww = h.shape
        C = f.numel() // (B * Hh * Ww)
        scale = 1.0 / np.sqrt(Hh * Ww) if self.use_scale else 1.0
        out = torch.empty(B, C, J, device=f.device)
        L.call("gait_keypoint_attention", L.ptr(f), L.ptr(h), float(scale), L.ptr(out), B, C, J, Hh * Ww, C * J, J, 1, L.stream_ptr())
        return out


class BidirectionalModel(nn.Module):
    """gait_feat_encoder.py:10-104: per-joint PARE features (+ camera-parameter embedding) -> 2-layer bidirectional GRU ->
    walking speed / step length MLPs on the final hidden states and a gait-phase MLP on every frame.
    Same constructor and state_dict keys.  As shipped the reference only runs with use_pareFeat=True (`xc` is undefined
    otherwise, :102,104), which is what is implemented here."""

    def __init__(self, seqlen, input_size=128, num_joints=24, num_outputs=3, estime_phase=True, fc_size=40, num_layers=2,
                 use_pareFeat=False):
        super().__init__()
        if not use_pareFeat:
            raise NotImplementedError("BidirectionalModel: only use_pareFeat=True runs in the reference (gait_feat_encoder.py:102,104)")
        self.estim_phase, self.num_outputs, self.num_layers, self.use_pareFeat = estime_phase, num_outputs, num_layers, use_pareFeat
        h_size, fc_size = 300, 100
        self.input_size = input_size * num_joints
        self.dropout = nn.Dropout(0.2)
        self.cparam_mpl = LocallyConnected2d(in_channels=3, out_channels=128, output_size=[num_joints, 1], kernel_size=1, stride=1)
        self.rnn = nn.GRU(input_size=self.input_size, hidden_size=h_size, num_layers=num_layers, batch_first=True, bidirectional=True)
        if num_outputs > 0:
            self.speed_mlp = nn.Sequential(nn.Linear(h_size * 2 * num_layers, fc_size), nn.LeakyReLU(0.05, inplace=True), nn.Linear(fc_size, 1))
            self.step_mlp = nn.Sequential(nn.Linear(h_size * 2 * num_layers, fc_size), nn.LeakyReLU(0.05, inplace=True), nn.Linear(fc_size, 2))
        if estime_phase:
            self.phase_mlp = nn.Sequential(nn.Linear(h_size * 2, fc_size), nn.LeakyReLU(0.05, inplace=True), nn.Linear(fc_size, 4), nn.Tanh())

    @torch.no_grad()
    def forward(self, x, cparams=None):
        if self.training:
            raise L.GaitLibraryError("BidirectionalModel kernels implement eval() semantics (dropout off); call .eval()")
        assert cparams is not None
        x, cparams = L.f32(x, "x"), L.f32(cparams, "cparams")
        b, n, cf = cparams.shape
        J = self.cparam_mpl.weight.shape[3]
        # cparams (b*n, cf, 1, 1) broadcast over the joints (stride 0), residual add fused: xs = x + xc
        cp = cparams.reshape(b * n, cf, 1, 1).expand(b * n, cf, J, 1)
        xc, xs = self.cparam_mpl.run(cp, resid=x.reshape(b * n, -1, J, 1))
        xc, xs = xc.reshape(b, n, -1), xs.reshape(b, n, -1)
        y_seq, _, hn = gru_forward(self.rnn, xs, return_hn=True)
        h = hn.permute(1, 0, 2).reshape(b, -1).contiguous()
        y = None
        if self.num_outputs:
            y = torch.empty(b, 3, device=x.device)
            _linear(_act(_linear(h, self.speed_mlp[0]), 0, 0.05), self.speed_mlp[2], out=y[:, 0:1], ldc=3)
            _linear(_act(_linear(h, self.step_mlp[0]), 0, 0.05), self.step_mlp[2], out=y[:, 1:3], ldc=3)
        if self.estim_phase:
            p = _act(_linear(_act(_linear(y_seq.reshape(b * n, -1), self.phase_mlp[0]), 0, 0.05), self.phase_mlp[2]), 1)
            return y, p.reshape(b, n, 4), xc
        return y, None, xc
