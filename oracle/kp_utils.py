"""Oracle restatement of the joint re-indexing in lib/data_utils/kp_utils.py
(convert_kps :26-36; spin2 names :211-242; kinectv2 names :904-931). numpy.

Test infrastructure: see oracle/__init__.py.  Pinned by tests/golden/kp_utils.npz.
"""
import numpy as np

SPIN2_NAMES = [
    'hip', 'lhip (SMPL)', 'rhip (SMPL)', 'spine (SMPL)', 'lknee', 'rknee', 'Spine (H36M)',
    'lankle', 'rankle', 'spine2', 'leftFoot', 'rightFoot', 'neck', 'lcollar', 'rcollar',
    'Head (H36M)', 'lshoulder', 'rshoulder', 'lelbow', 'relbow', 'lwrist', 'rwrist',
    'leftHand', 'rightHand', 'leftThumb', 'leftHandTip', 'rightThumb', 'rightHandTip', 'thorax',
]
KINECTV2_NAMES = [
    'hip', 'Spine (H36M)', 'neck', 'Head (H36M)', 'lshoulder', 'lelbow', 'lwrist', 'leftHand',
    'rshoulder', 'relbow', 'rwrist', 'rightHand', 'lhip (SMPL)', 'lknee', 'lankle', 'leftFoot',
    'rhip (SMPL)', 'rknee', 'rankle', 'rightFoot', 'thorax', 'leftHandTip', 'leftThumb',
    'rightHandTip', 'rightThumb',
]
_TABLES = {'spin2': SPIN2_NAMES, 'kinectv2': KINECTV2_NAMES}


def convert_kps(joints, src, dst):
    """kp_utils.py:26-36 - name-matched copy into a zero float64 (N, len(dst), 3) array."""
    s, d = _TABLES[src], _TABLES[dst]
    out = np.zeros((joints.shape[0], len(d), 3))
    for i, name in enumerate(d):
        if name in s:
            out[:, i] = joints[:, s.index(name)]
    return out


def gather_indices(src, dst):
    """dst-length list of source indices (-1 where the name is missing in src)."""
    s, d = _TABLES[src], _TABLES[dst]
    return [s.index(n) if n in s else -1 for n in d]
