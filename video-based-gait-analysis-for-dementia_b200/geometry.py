"""Mirror of the hot-path free functions of the reference's lib/utils/geometry.py.

Same names, argument meaning, output shapes and error behaviour; every function runs one
hand-written sm_100a kernel through the C-ABI (include/gaitb200.h).  Inputs must be FP32
CUDA tensors - there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib as L


def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    """geometry.py:395-410 - (B,6)/(B*24,6)/(B,144) -> (N,3,3); F.normalize eps = 1e-6."""
    x = L.f32(x, "x").reshape(-1, 6)
    out = torch.empty(x.shape[0], 3, 3, device=x.device, dtype=torch.float32)
    L.call("gait_rot6d_to_rotmat", L.ptr(x), 1, 6, L.ptr(out), x.shape[0], 1e-6, L.stream_ptr())
    return out


def rot6d_to_rotmat_spin(x: torch.Tensor) -> torch.Tensor:
    """geometry.py:368-387 - same construction with F.normalize's default eps (1e-12)."""
    x = L.f32(x, "x").reshape(-1, 6)
    out = torch.empty(x.shape[0], 3, 3, device=x.device, dtype=torch.float32)
    L.call("gait_rot6d_to_rotmat", L.ptr(x), 1, 6, L.ptr(out), x.shape[0], 1e-12, L.stream_ptr())
    return out


def rotmat_to_rot6d(x: torch.Tensor) -> torch.Tensor:
    """geometry.py:389-393 - (N,3,3) -> (N,3,2)."""
    x = L.f32(x, "x").reshape(-1, 3, 3)
    out = torch.empty(x.shape[0], 3, 2, device=x.device, dtype=torch.float32)
    L.call("gait_rotmat_to_rot6d", L.ptr(x), L.ptr(out), x.shape[0], L.stream_ptr())
    return out


def rotation_matrix_to_quaternion(rotation_matrix: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """geometry.py:213-293 - (N,3,4) or (N,3,3) -> (N,4) in (w,x,y,z)."""
    if not torch.is_tensor(rotation_matrix):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(rotation_matrix)))
    if len(rotation_matrix.shape) > 3:
        raise ValueError("Input size must be a three dimensional tensor. Got {}".format(rotation_matrix.shape))
    if rotation_matrix.dim() != 3 or tuple(rotation_matrix.shape[-2:]) not in ((3, 4), (3, 3)):
        raise ValueError("Input size must be a N x 3 x 4 or N x 3 x 3 tensor. Got {}".format(rotation_matrix.shape))
    r = L.f32(rotation_matrix, "rotation_matrix")
    out = torch.empty(r.shape[0], 4, device=r.device, dtype=torch.float32)
    L.call("gait_rotmat_to_quaternion", L.ptr(r), r.shape[-1], L.ptr(out), r.shape[0], float(eps), L.stream_ptr())
    return out


def quaternion_to_angle_axis(quaternion: torch.Tensor) -> torch.Tensor:
    """geometry.py:159-210 - (...,4) -> (...,3)."""
    if not torch.is_tensor(quaternion):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(quaternion)))
    if not quaternion.shape[-1] == 4:
        raise ValueError("Input must be a tensor of shape Nx4 or 4. Got {}".format(quaternion.shape))
    q = L.f32(quaternion, "quaternion")
    out = torch.empty(*q.shape[:-1], 3, device=q.device, dtype=torch.float32)
    L.call("gait_quaternion_to_axis_angle", L.ptr(q), L.ptr(out), q.numel() // 4, L.stream_ptr())
    return out


def rotation_matrix_to_angle_axis(rotation_matrix: torch.Tensor) -> torch.Tensor:
    """geometry.py:68-97 - (N,3,3) or (N,3,4) -> (N,3); NaN entries are zeroed."""
    if not torch.is_tensor(rotation_matrix):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(rotation_matrix)))
    if rotation_matrix.dim() != 3 or tuple(rotation_matrix.shape[-2:]) not in ((3, 4), (3, 3)):
        raise ValueError("Input size must be a N x 3 x 4 or N x 3 x 3 tensor. Got {}".format(rotation_matrix.shape))
    r = L.f32(rotation_matrix, "rotation_matrix")
    out = torch.empty(r.shape[0], 3, device=r.device, dtype=torch.float32)
    L.call("gait_rotmat_to_axis_angle", L.ptr(r), r.shape[-1], L.ptr(out), r.shape[0], 1, 3, 0, L.stream_ptr())
    return out


def quat2mat(quat: torch.Tensor) -> torch.Tensor:
    """geometry.py:38-65 - (B,4) (w,x,y,z), normalised inside -> (B,3,3)."""
    q = L.f32(quat, "quat").reshape(-1, 4)
    out = torch.empty(q.shape[0], 3, 3, device=q.device, dtype=torch.float32)
    L.call("gait_quat2mat", L.ptr(q), L.ptr(out), q.shape[0], L.stream_ptr())
    return out


def batch_rodrigues(axisang: torch.Tensor) -> torch.Tensor:
    """geometry.py:23-35 - (N,3) axis-angle -> (N,9) via the half-angle quaternion."""
    a = L.f32(axisang, "axisang").reshape(-1, 3)
    out = torch.empty(a.shape[0], 9, device=a.device, dtype=torch.float32)
    L.call("gait_batch_rodrigues", L.ptr(a), L.ptr(out), a.shape[0], 1, L.stream_ptr())
    return out


def batch_rodrigues_smplx(rot_vecs: torch.Tensor) -> torch.Tensor:
    """smplx lbs.batch_rodrigues (what SMPL.forward(pose2rot=True) uses) - (N,3) -> (N,3,3)."""
    a = L.f32(rot_vecs, "rot_vecs").reshape(-1, 3)
    out = torch.empty(a.shape[0], 3, 3, device=a.device, dtype=torch.float32)
    L.call("gait_batch_rodrigues", L.ptr(a), L.ptr(out), a.shape[0], 0, L.stream_ptr())
    return out


def convert_weak_perspective_to_perspective(weak_perspective_camera, focal_length=5000., img_res=224):
    """geometry.py:427-446 - [s,tx,ty] -> [tx,ty,2f/(res*s+1e-9)]."""
    cam = L.f32(weak_perspective_camera, "weak_perspective_camera").reshape(-1, 3)
    out = torch.empty_like(cam)
    L.call("gait_weak_perspective_to_translation", L.ptr(cam), L.ptr(out), cam.shape[0], float(focal_length),
           float(img_res), L.stream_ptr())
    return out


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """geometry.py:448-479 - points (B,N,3), rotation (B,3,3), translation (B,3),
    camera_center (B,2) -> (B,N,2)."""
    pts = L.f32(points, "points")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError(f"points must be (B,N,3), got {tuple(pts.shape)}")
    b, n = pts.shape[:2]
    rot = None if rotation is None else L.f32(rotation, "rotation").expand(b, 3, 3).contiguous()
    trans = L.f32(translation, "translation").reshape(b, 3)
    cen = None if camera_center is None else L.f32(camera_center, "camera_center").reshape(b, 2)
    out = torch.empty(b, n, 2, device=pts.device, dtype=torch.float32)
    L.call("gait_perspective_projection", L.ptr(pts), L.ptr(rot), L.ptr(trans), L.ptr(cen), float(focal_length), 1.0,
           L.ptr(out), b, n, L.stream_ptr())
    return out


def projection(pred_joints, pred_camera):
    """geometry.py:412-425 - weak-perspective camera with the hard-coded 5000/224, result / 112."""
    pts = L.f32(pred_joints, "pred_joints")
    if pts.dim() != 3 or pts.shape[-1] != 3:
        raise ValueError(f"pred_joints must be (B,N,3), got {tuple(pts.shape)}")
    b, n = pts.shape[:2]
    trans = convert_weak_perspective_to_perspective(pred_camera, 5000., 224.)
    out = torch.empty(b, n, 2, device=pts.device, dtype=torch.float32)
    L.call("gait_perspective_projection", L.ptr(pts), None, L.ptr(trans), None, 5000., 224. / 2., L.ptr(out), b, n,
           L.stream_ptr())
    return out
