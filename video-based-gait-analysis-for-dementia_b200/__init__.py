"""gaitb200 - B200-native (sm_100a) regression head for MAX-GRNet.

Mirrors the reference's module API for the hot path (TemporalEncoder, Regressor,
VPRegressor/SMPLRegressor, SMPL/SMPLHead, the geometry free functions, convert_kps)
on top of hand-written CUDA kernels reached through the C-ABI in include/gaitb200.h.
There is no CPU path: every op raises if the CUDA library is missing or a tensor
is not on a CUDA device.

Import is cheap and works without a GPU (the library is loaded on first use);
sub-modules: geometry, kp_utils, smpl, regressor, temporal, head, sharding, synthetic.
"""
__version__ = "0.1.0"

_LAZY = {
    "TemporalEncoder": "temporal", "Regressor": "regressor", "VPRegressor": "regressor",
    "SMPLRegressor": "regressor", "SMPL": "smpl", "SMPLHead": "smpl", "SMPLOutput": "smpl",
    "GaitHead": "head", "convert_kps": "kp_utils",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f"gaitb200.{_LAZY[name]}"), name)
    raise AttributeError(f"module 'gaitb200' has no attribute {name!r}")
