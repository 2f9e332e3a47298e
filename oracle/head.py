"""Oracle composition of the whole hot path: backbone features -> TemporalEncoder ->
Regressor (MLP loop, rot6d->R, SMPL, projection, axis-angle) -> Kinect-25 joints.

This is the CPU FP32 arm bench.py times (cpu_baseline / --impl reference) and the
end-to-end parity target.  Test infrastructure: see oracle/__init__.py.
"""
import torch
import torch.nn as nn

from .kp_utils import gather_indices
from .regressor import Regressor
from .temporal import TemporalEncoder


class GaitHeadOracle(nn.Module):
    def __init__(self, smpl_data, mean_params, regressor_state=None, gru_state=None, **encoder_kw):
        super().__init__()
        self.encoder = TemporalEncoder(**encoder_kw)
        self.regressor = Regressor(smpl_data, mean_params)
        if gru_state is not None:
            self.encoder.gru.load_state_dict(gru_state)
        if regressor_state is not None:
            self.regressor.load_state_dict(regressor_state, strict=False)
        self.kinect_idx = torch.tensor(gather_indices('spin2', 'kinectv2'))
        self.eval()

    @torch.no_grad()
    def forward(self, features, J_regressor=None):
        s, t = features.shape[:2]
        y = self.encoder(features).reshape(s * t, -1)
        out = self.regressor(y, J_regressor=J_regressor)[-1]
        res = {k: v.reshape(s, t, *v.shape[1:]) for k, v in out.items()}
        if J_regressor is None and out['kp_3d'].shape[1] == 29:
            res['kinect25'] = res['kp_3d'][:, :, self.kinect_idx]
        return res
