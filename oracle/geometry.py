"""Oracle restatement of lib/utils/geometry.py (hot-path functions only). Torch CPU FP32.

Test infrastructure: see oracle/__init__.py.  Each function names the reference
lines it follows; tests/test_oracle_golden.py checks every one against vectors
produced by the reference's own module.
"""
import torch
import torch.nn.functional as F


def rot6d_to_rotmat(x):
    """geometry.py:395-410 - Gram-Schmidt on a (3,2)-interleaved 6-vector, columns stacked."""
    m = x.reshape(-1, 3, 2)
    a1, a2 = m[..., 0], m[..., 1]
    b1 = F.normalize(a1, dim=1, eps=1e-6)
    proj = (b1 * a2).sum(dim=1, keepdim=True)
    b2 = F.normalize(a2 - proj * b1, dim=-1, eps=1e-6)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def rot6d_to_rotmat_spin(x):
    """geometry.py:368-387 - same construction with F.normalize's default eps (1e-12)."""
    m = x.view(-1, 3, 2)
    a1, a2 = m[:, :, 0], m[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def rotmat_to_rot6d(x):
    """geometry.py:389-393 - first two columns, interleaved."""
    r = x.reshape(-1, 3, 3)
    return torch.stack((r[:, :, 0], r[:, :, 1]), dim=-1)


def quat2mat(quat):
    """geometry.py:38-65 - normalise (w,x,y,z), expand to a 3x3 matrix."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q.unbind(dim=1)
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    rows = [w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
            2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
            2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def batch_rodrigues(axisang):
    """geometry.py:23-35 - half-angle quaternion route; returns (N, 9)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1).unsqueeze(-1)
    unit = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * unit], dim=1)
    return quat2mat(quat).view(-1, 9)


def rotation_matrix_to_quaternion(rotation_matrix, eps=1e-6):
    """geometry.py:213-293 - four-branch (w,x,y,z) extraction on the transposed matrix."""
    if not torch.is_tensor(rotation_matrix):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(rotation_matrix)}")
    if rotation_matrix.dim() > 3:
        raise ValueError(f"Input size must be a three dimensional tensor. Got {rotation_matrix.shape}")
    if rotation_matrix.shape[-2:] not in ((3, 4), (3, 3)):
        raise ValueError(f"Input size must be a N x 3 x 4 or N x 3 x 3 tensor. Got {rotation_matrix.shape}")
    m = rotation_matrix.transpose(1, 2)
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    d2 = m22 < eps
    d0_gt_d1 = m00 > m11
    d0_lt_nd1 = m00 < -m11

    t0 = 1 + m00 - m11 - m22
    q0 = torch.stack([m[:, 1, 2] - m[:, 2, 1], t0, m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], -1)
    t1 = 1 - m00 + m11 - m22
    q1 = torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t1, m[:, 1, 2] + m[:, 2, 1]], -1)
    t2 = 1 - m00 - m11 + m22
    q2 = torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t2], -1)
    t3 = 1 + m00 + m11 + m22
    q3 = torch.stack([t3, m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], -1)

    c0 = (d2 & d0_gt_d1).view(-1, 1).type_as(q0)
    c1 = (d2 & ~d0_gt_d1).view(-1, 1).type_as(q0)
    c2 = (~d2 & d0_lt_nd1).view(-1, 1).type_as(q0)
    c3 = (~d2 & ~d0_lt_nd1).view(-1, 1).type_as(q0)
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    denom = t0.view(-1, 1) * c0 + t1.view(-1, 1) * c1 + t2.view(-1, 1) * c2 + t3.view(-1, 1) * c3
    q = q / torch.sqrt(denom.expand(-1, 4))
    return q * 0.5


def quaternion_to_angle_axis(quaternion):
    """geometry.py:159-210 - atan2 branch on the sign of w; k=2 where sin^2 == 0."""
    if not torch.is_tensor(quaternion):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(quaternion)}")
    if quaternion.shape[-1] != 4:
        raise ValueError(f"Input must be a tensor of shape Nx4 or 4. Got {quaternion.shape}")
    q1, q2, q3 = quaternion[..., 1], quaternion[..., 2], quaternion[..., 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = torch.sqrt(s2)
    c = quaternion[..., 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    return torch.stack((q1 * k, q2 * k, q3 * k), dim=-1)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """geometry.py:68-97 - (N,3,3) gets a dummy 4th column, then quaternion route; NaN -> 0."""
    if rotation_matrix.shape[1:] == (3, 3):
        r = rotation_matrix.reshape(-1, 3, 3)
        hom = torch.tensor([0, 0, 1], dtype=torch.float32).reshape(1, 3, 1).expand(r.shape[0], -1, -1)
        rotation_matrix = torch.cat([r, hom], dim=-1)
    aa = quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix))
    aa[torch.isnan(aa)] = 0.0
    return aa


def convert_weak_perspective_to_perspective(cam, focal_length=5000., img_res=224):
    """geometry.py:427-446 - [s,tx,ty] -> [tx,ty,2f/(res*s+1e-9)]."""
    return torch.stack([cam[:, 1], cam[:, 2], 2 * focal_length / (img_res * cam[:, 0] + 1e-9)], dim=-1)


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """geometry.py:448-479 - rotate, translate, divide by z, apply K; returns (B,N,2)."""
    b = points.shape[0]
    K = torch.zeros(b, 3, 3)
    K[:, 0, 0] = focal_length
    K[:, 1, 1] = focal_length
    K[:, 2, 2] = 1.
    K[:, :-1, -1] = camera_center
    p = torch.einsum('bij,bkj->bki', rotation, points) + translation.unsqueeze(1)
    p = p / p[:, :, -1].unsqueeze(-1)
    p = torch.einsum('bij,bkj->bki', K, p)
    return p[:, :, :-1]


def projection(pred_joints, pred_camera):
    """geometry.py:412-425 - weak-perspective camera with hard-coded 5000/224, result /112."""
    t = torch.stack([pred_camera[:, 1], pred_camera[:, 2],
                     2 * 5000. / (224. * pred_camera[:, 0] + 1e-9)], dim=-1)
    b = pred_joints.shape[0]
    kp = perspective_projection(pred_joints, torch.eye(3).unsqueeze(0).expand(b, -1, -1), t,
                                5000., torch.zeros(b, 2))
    return kp / (224. / 2.)
