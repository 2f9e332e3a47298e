"""Mirror of the joint re-indexing the hot path uses from lib/data_utils/kp_utils.py
(convert_kps :26-36; spin2 names :211-242; kinectv2 names :904-931; skeleton :933-942).

The name tables are host constants; the copy itself runs on the GPU (one gather kernel),
so the Kinect-25 joints never have to leave the device in the 29-joint layout first.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def get_spin2_joint_names():
    return [
        'hip', 'lhip (SMPL)', 'rhip (SMPL)', 'spine (SMPL)', 'lknee', 'rknee', 'Spine (H36M)',
        'lankle', 'rankle', 'spine2', 'leftFoot', 'rightFoot', 'neck', 'lcollar', 'rcollar',
        'Head (H36M)', 'lshoulder', 'rshoulder', 'lelbow', 'relbow', 'lwrist', 'rwrist',
        'leftHand', 'rightHand', 'leftThumb', 'leftHandTip', 'rightThumb', 'rightHandTip', 'thorax',
    ]


def get_kinectv2_joint_names():
    return [
        'hip', 'Spine (H36M)', 'neck', 'Head (H36M)', 'lshoulder', 'lelbow', 'lwrist', 'leftHand',
        'rshoulder', 'relbow', 'rwrist', 'rightHand', 'lhip (SMPL)', 'lknee', 'lankle', 'leftFoot',
        'rhip (SMPL)', 'rknee', 'rankle', 'rightFoot', 'thorax', 'leftHandTip', 'leftThumb',
        'rightHandTip', 'rightThumb',
    ]


def get_kinectv2_skeleton():
    """Bone list (pairs of kinectv2 joint indices): trunk, arms, hands, legs, feet."""
    trunk = [(0, 1), (20, 2), (1, 20), (2, 3)]
    arms = [(20, 4), (20, 8), (4, 5), (8, 9), (5, 6), (9, 10)]
    hands = [(6, 7), (10, 11), (7, 21), (11, 23), (6, 22), (10, 24)]
    legs = [(0, 12), (0, 16), (12, 13), (16, 17), (13, 14), (17, 18)]
    feet = [(14, 15), (18, 19)]
    return np.array(trunk + arms + hands + legs + feet)


_TABLES = {'spin2': get_spin2_joint_names, 'kinectv2': get_kinectv2_joint_names}


def gather_indices(src: str, dst: str):
    """dst-length list of source indices (-1 where dst names a joint src lacks -> zeros)."""
    if src not in _TABLES or dst not in _TABLES:
        raise NameError(f"name 'get_{src if src not in _TABLES else dst}_joint_names' is not defined")
    s, d = _TABLES[src](), _TABLES[dst]()
    return [s.index(n) if n in s else -1 for n in d]


SPIN2_TO_KINECTV2 = gather_indices('spin2', 'kinectv2')

_idx_cache = {}


def convert_kps(joints2d: torch.Tensor, src: str, dst: str) -> torch.Tensor:
    """kp_utils.py:26-36 on the device: (N, len(src), 3) CUDA FP32 -> (N, len(dst), 3); joints
    missing from `src` are zero.  (The reference's numpy version returns float64 zeros-initialised
    arrays; values are identical.)"""
    idx = gather_indices(src, dst)
    j = L.f32(joints2d, "joints2d")
    if j.dim() != 3 or j.shape[1] != len(_TABLES[src]()) or j.shape[2] != 3:
        raise ValueError(f"expected (N,{len(_TABLES[src]())},3) joints in '{src}' order, got {tuple(j.shape)}")
    key = (src, dst, j.device)
    if key not in _idx_cache:
        _idx_cache[key] = torch.tensor(idx, dtype=torch.int32, device=j.device)
    out = torch.empty(j.shape[0], len(idx), 3, device=j.device, dtype=torch.float32)
    L.call("gait_gather_joints", L.ptr(j), j.shape[1], L.ptr(_idx_cache[key]), len(idx), L.ptr(out), j.shape[0],
           L.stream_ptr())
    return out
