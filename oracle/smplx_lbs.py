"""Oracle restatement of the smplx==0.1.26 arithmetic the reference calls.

PARITY UNPINNED: smplx (requirements.txt:13) is a third-party dependency that is
neither under /root/reference nor installable offline, and the reference holds no
test or golden vector for it.  This restates its published algorithm
(smplx/lbs.py: blend_shapes, vertices2joints, batch_rodrigues, transform_mat,
batch_rigid_transform, lbs; smplx/vertex_joint_selector.py; body_models.SMPL.forward)
and is anchored on the reference's call sites lib/models/smpl.py:8-10,111,113 and
the analytic known-answer tests in tests/test_oracle_kat.py.  Torch CPU FP32.
Test infrastructure: see oracle/__init__.py.
"""
from collections import namedtuple

import torch
import torch.nn.functional as F

SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose"])
SMPLOutput.__new__.__defaults__ = (None,) * 6


def blend_shapes(betas, shape_disps):
    """lbs.py blend_shapes: einsum('bl,mkl->bmk')."""
    return torch.einsum('bl,mkl->bmk', betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    """lbs.py vertices2joints: einsum('bik,ji->bjk')."""
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


def batch_rodrigues(rot_vecs):
    """lbs.py batch_rodrigues: angle = ||a + 1e-8||, R = I + sin K + (1-cos) K^2."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    cos = torch.cos(angle).unsqueeze(1)
    sin = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    zeros = torch.zeros(n, 1, dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """lbs.py batch_rigid_transform: 24-joint chain G_i = G_parent(i) @ [[R_i, J_i - J_parent],[0,1]];
    returns posed joints and A_i = G_i with the rest-pose joint removed."""
    b, nj = joints.shape[:2]
    j = joints.unsqueeze(-1)
    rel = j.clone()
    rel[:, 1:] -= j[:, parents[1:]]
    top = torch.cat([rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)], dim=2)
    bottom = torch.tensor([0., 0., 0., 1.]).view(1, 1, 4).expand(top.shape[0], -1, -1)
    local = torch.cat([top, bottom], dim=1).view(b, nj, 4, 4)
    chain = [local[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], local[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    j_h = F.pad(j, [0, 0, 0, 1])
    A = G - F.pad(torch.matmul(G, j_h), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True):
    """lbs.py lbs: shape blend -> joints -> pose blend -> chain -> W.A -> apply."""
    b = max(betas.shape[0], pose.shape[0])
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3)
    if pose2rot:
        rot_mats = batch_rodrigues(pose.reshape(-1, 3)).view(b, -1, 3, 3)
        pose_feature = (rot_mats[:, 1:] - ident).view(b, -1)
    else:
        rot_mats = pose.reshape(b, -1, 3, 3)
        pose_feature = (rot_mats[:, 1:] - ident).reshape(b, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(b, -1, 3)
    J_t, A = batch_rigid_transform(rot_mats, J, parents)
    W = lbs_weights.unsqueeze(0).expand(b, -1, -1)
    nj = J_regressor.shape[0]
    T = torch.matmul(W, A.view(b, nj, 16)).view(b, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(b, v_posed.shape[1], 1)], dim=2)
    verts = torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_t


class SMPLX_SMPL(torch.nn.Module):
    """body_models.SMPL restated for the arguments the reference passes
    (create_transl=False, gender neutral, betas/body_pose/global_orient given).
    Buffer names match smplx so checkpoints keyed `...smpl.<name>` line up."""

    def __init__(self, data: dict, batch_size: int = 1):
        super().__init__()
        t = lambda a, dt=torch.float32: torch.as_tensor(a, dtype=dt)
        self.faces = data["faces"]
        self.register_buffer("faces_tensor", t(data["faces"], torch.long))
        self.register_parameter("betas", torch.nn.Parameter(torch.zeros(batch_size, 10)))
        self.register_parameter("global_orient", torch.nn.Parameter(torch.zeros(batch_size, 3)))
        self.register_parameter("body_pose", torch.nn.Parameter(torch.zeros(batch_size, 69)))
        self.register_buffer("v_template", t(data["v_template"]))
        self.register_buffer("shapedirs", t(data["shapedirs"])[:, :, :10])
        self.register_buffer("J_regressor", t(data["J_regressor"]))
        self.register_buffer("posedirs", t(data["posedirs"]))
        self.register_buffer("parents", t(data["parents"], torch.long))
        self.register_buffer("lbs_weights", t(data["lbs_weights"]))
        self.register_buffer("extra_joints_idxs", t(data["landmark_verts"], torch.long))

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, **kwargs):
        global_orient = self.global_orient if global_orient is None else global_orient
        body_pose = self.body_pose if body_pose is None else body_pose
        betas = self.betas if betas is None else betas
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        verts, joints = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs,
                            self.J_regressor, self.parents, self.lbs_weights, pose2rot=pose2rot)
        # vertex_joint_selector.py: append landmark vertices as joints 24..44
        joints = torch.cat([joints, torch.index_select(verts, 1, self.extra_joints_idxs)], dim=1)
        return SMPLOutput(vertices=verts, joints=joints, full_pose=None, betas=betas,
                          global_orient=global_orient, body_pose=body_pose)
