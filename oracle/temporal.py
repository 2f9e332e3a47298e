"""Oracle TemporalEncoder: VIBE's GRU encoder over 2048-d backbone features.

The class is ABSENT from the reference tree; BASELINE.json's
north_star names the VIBE-lineage API, so the oracle *defines* it (VIBE
lib/models/vibe.py signature) on top of torch.nn.GRU - the arithmetic the
reference's only GRU delegates to (lib/models/layers/gait_feat_encoder.py:51-57,88).
Test infrastructure: see oracle/__init__.py.
"""
import torch.nn as nn
import torch.nn.functional as F


class TemporalEncoder(nn.Module):
    def __init__(self, n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False,
                 use_residual=True, input_size=2048):
        super().__init__()
        self.gru = nn.GRU(input_size=input_size, hidden_size=hidden_size,
                          bidirectional=bidirectional, num_layers=n_layers)
        self.linear = None
        if bidirectional:
            self.linear = nn.Linear(hidden_size * 2, input_size)
        elif add_linear:
            self.linear = nn.Linear(hidden_size, input_size)
        self.use_residual = use_residual
        self.input_size = input_size

    def forward(self, x):
        n, t, f = x.shape
        x = x.permute(1, 0, 2)                      # NTF -> TNF
        y, _ = self.gru(x)
        if self.linear is not None:
            y = self.linear(F.relu(y).view(-1, y.size(-1))).view(t, n, f)
        if self.use_residual and y.shape[-1] == self.input_size:
            y = y + x
        return y.permute(1, 0, 2)                   # TNF -> NTF
