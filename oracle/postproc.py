"""Oracle restatement of the post-processing that follows / re-uses the regression head:

  * OneEuroFilter                      lib/utils/one_euro_filter.py:5-46
  * smooth_pose                        lib/utils/smooth_pose.py:28-116
  * convert_crop_cam_to_orig_img       lib/utils/demo_utils.py:176-193
  * convert_crop_coords_to_orig_img    lib/utils/demo_utils.py:196-209

numpy / torch CPU.  Test infrastructure: see oracle/__init__.py.  Pinned by tests/golden/postproc.npz, produced by
executing the reference's own modules (tests/golden/make_golden_post.py).
"""
import math

import numpy as np
import torch

from . import geometry as G
from .kp_utils import convert_kps
from .smpl import SMPL


def one_euro_filter(x, min_cutoff=1.0, beta=0.0, d_cutoff=1.0):
    """The filter as smooth_pose.py:51-56,84-88 drives it: x (T, ...) sampled at t = 0,1,2,...; x_hat[0] = x[0],
    dx_prev = 0.0, t_prev starts at zeros, so t_e = 1 at every step.  Arithmetic stays in x.dtype (numpy keeps
    float32 arrays float32 against Python scalars), one_euro_filter.py:5-46."""
    x = np.asarray(x)
    one = np.ones_like(x[0])
    hat = np.zeros_like(x)
    hat[0] = x[0]
    x_prev, dx_prev, t_prev = x[0], 0.0, np.zeros_like(x[0])
    for i in range(1, x.shape[0]):
        t = one * i
        t_e = t - t_prev
        r = 2 * math.pi * float(d_cutoff) * t_e
        a_d = r / (r + 1)
        dx = (x[i] - x_prev) / t_e
        dx_hat = a_d * dx + (1 - a_d) * dx_prev
        cutoff = float(min_cutoff) + float(beta) * np.abs(dx_hat)
        r = 2 * math.pi * cutoff * t_e
        a = r / (r + 1)
        x_hat = a * x[i] + (1 - a) * x_prev
        x_prev, dx_prev, t_prev = x_hat, dx_hat, t
        hat[i] = x_hat
    return hat


def smooth_pose(smpl_data, pred_pose, pred_betas, min_cutoff=0.004, beta=0.7, kinectv2=False):
    """smooth_pose.py:28-116: filter the (T,72) axis-angle or (T,96) quaternion poses, then one SMPL forward per frame
    (pose2rot=True) with the FIRST frame's betas (smooth_pose.py:73,97).  Returns (verts, pose_hat, joints3d)."""
    T = pred_betas.shape[0]
    if pred_pose.shape[-1] == 72:
        q, shape = 3, pred_pose.shape
    elif pred_pose.shape[-1] == 96:
        q, shape = 4, pred_pose.shape
    else:
        raise ValueError(f"Invalid pred_pose format: {pred_pose.shape}")
    pose = pred_pose.reshape(T, 24, q)
    hat = one_euro_filter(pose, min_cutoff=min_cutoff, beta=beta)
    smpl = SMPL(smpl_data)
    smpl.kinectv2 = kinectv2
    verts, joints = [], []
    b0 = torch.from_numpy(pred_betas[0]).unsqueeze(0)
    with torch.no_grad():
        for i in range(T):
            aa = torch.from_numpy(hat[i]) if q == 3 else G.quaternion_to_angle_axis(torch.from_numpy(hat[i].reshape(-1, 4)).float())
            so = smpl(betas=b0, body_pose=aa[1:].unsqueeze(0), global_orient=aa[0:1].unsqueeze(0))
            verts.append(so.vertices.numpy())
            joints.append(so.joints.numpy())
    j = np.vstack(joints)
    if kinectv2:
        j = convert_kps(j, 'spin2', 'kinectv2')
    return np.vstack(verts), hat.reshape(shape), j


def convert_crop_cam_to_orig_img(cam, bbox, img_width, img_height):
    """demo_utils.py:176-193: weak-perspective camera of the crop -> [sx, sy, tx, ty] in the original image."""
    cx, cy, h = bbox[:, 0], bbox[:, 1], bbox[:, 2]
    hw, hh = img_width / 2., img_height / 2.
    sx = cam[:, 0] * (1. / (img_width / h))
    sy = cam[:, 0] * (1. / (img_height / h))
    tx = ((cx - hw) / hw / sx) + cam[:, 1]
    ty = ((cy - hh) / hh / sy) + cam[:, 2]
    return np.stack([sx, sy, tx, ty]).T


def convert_crop_coords_to_orig_img(bbox, keypoints, crop_size):
    """demo_utils.py:196-209: keypoints in [-1,1] crop units -> original image pixels (array dtype of `keypoints` kept:
    the reference updates it in place)."""
    cx, cy, h = bbox[:, 0], bbox[:, 1], bbox[:, 2]
    keypoints = 0.5 * crop_size * (keypoints + 1.0)
    keypoints *= h[..., None, None] / crop_size
    keypoints[:, :, 0] = (cx - h / 2)[..., None] + keypoints[:, :, 0]
    keypoints[:, :, 1] = (cy - h / 2)[..., None] + keypoints[:, :, 1]
    return keypoints
