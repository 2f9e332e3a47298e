"""gaitb200 - B200-native (sm_100a) regression head for MAX-GRNet.

Mirrors the reference's module API for the hot path (TemporalEncoder, Regressor,
VPRegressor/SMPLRegressor, SMPL/SMPLHead, the geometry free functions, convert_kps)
on top of hand-written CUDA kernels reached through the C-ABI in include/gaitb200.h.
There is no CPU path: every op raises if the CUDA library is missing or a tensor
is not on a CUDA device.
"""
__version__ = "0.1.0"
