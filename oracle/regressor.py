"""Oracle restatement of spin.Regressor (lib/models/spin.py:210-295) and
pare.VPRegressor / pare.SMPLRegressor (lib/models/pare.py:24-142). Torch CPU FP32.

Test infrastructure: see oracle/__init__.py.  Pinned by tests/golden/{regressor,vpregressor}.npz.
"""
import torch
import torch.nn as nn

from . import geometry as G
from .smpl import SMPL, SMPLHead, H36M_TO_J14


class Regressor(nn.Module):
    """spin.py:210-295 - iterative [x|pose6d|beta|cam] -> fc1 -> fc2 -> 3 decoders, residual
    updates, no non-linearity; dropout is the identity in eval()."""

    def __init__(self, smpl_data: dict, mean_params: dict):
        super().__init__()
        npose = 24 * 6
        self.fc1 = nn.Linear(512 * 4 + npose + 13, 1024)
        self.drop1 = nn.Dropout()
        self.fc2 = nn.Linear(1024, 1024)
        self.drop2 = nn.Dropout()
        self.decpose = nn.Linear(1024, npose)
        self.decshape = nn.Linear(1024, 10)
        self.deccam = nn.Linear(1024, 3)
        for m in (self.decpose, self.decshape, self.deccam):
            nn.init.xavier_uniform_(m.weight, gain=0.01)
        self.smpl = SMPL(smpl_data, batch_size=64)
        self.register_buffer('init_pose', torch.as_tensor(mean_params['pose'], dtype=torch.float32).unsqueeze(0))
        self.register_buffer('init_shape', torch.as_tensor(mean_params['shape'], dtype=torch.float32).unsqueeze(0))
        self.register_buffer('init_cam', torch.as_tensor(mean_params['cam'], dtype=torch.float32).unsqueeze(0))

    def iterate(self, x, init_pose=None, init_shape=None, init_cam=None, n_iter=3):
        """spin.py:244-265 - the MLP loop alone; returns (pose6d, betas, cam)."""
        b = x.shape[0]
        pose = self.init_pose.expand(b, -1) if init_pose is None else init_pose
        shape = self.init_shape.expand(b, -1) if init_shape is None else init_shape
        cam = self.init_cam.expand(b, -1) if init_cam is None else init_cam
        for _ in range(n_iter):
            h = self.drop2(self.fc2(self.drop1(self.fc1(torch.cat([x, pose, shape, cam], 1)))))
            pose = self.decpose(h) + pose
            shape = self.decshape(h) + shape
            cam = self.deccam(h) + cam
        return pose, shape, cam

    def forward(self, x, init_pose=None, init_shape=None, init_cam=None, n_iter=3, J_regressor=None):
        b = x.shape[0]
        pose, shape, cam = self.iterate(x, init_pose, init_shape, init_cam, n_iter)
        rotmat = G.rot6d_to_rotmat(pose).view(b, 24, 3, 3)
        so = self.smpl(betas=shape, body_pose=rotmat[:, 1:], global_orient=rotmat[:, 0].unsqueeze(1), pose2rot=False)
        verts, joints = so.vertices, so.joints
        if J_regressor is not None:
            joints = torch.matmul(J_regressor[None, :].expand(b, -1, -1), verts)[:, H36M_TO_J14, :]
        kp2d = G.projection(joints, cam)
        aa = G.rotation_matrix_to_angle_axis(rotmat.reshape(-1, 3, 3)).reshape(-1, 72)
        return [{'theta': torch.cat([cam, aa, shape], dim=1), 'verts': verts, 'kp_2d': kp2d,
                 'kp_3d': joints, 'rotmat': rotmat}]


def _smpl_stage(head: SMPLHead, rotmat, shape, cam, batch_size, J_regressor):
    """Shared body of pare.py:52-76 and pare.py:108-131."""
    so = head(rotmat=rotmat, shape=shape, cam=cam, normalize_joints2d=True)
    aa = G.rotation_matrix_to_angle_axis(rotmat.reshape(-1, 3, 3)).reshape(-1, 72)
    seqlen = int(aa.shape[0] / batch_size)
    if J_regressor is not None:
        v = so['smpl_vertices'].reshape(batch_size * seqlen, -1, 3)
        j = torch.matmul(J_regressor[None, :].expand(v.shape[0], -1, -1), v)
        if J_regressor.shape[0] < 24:
            j = j[:, H36M_TO_J14, :]
        so['smpl_joints3d'] = j
    return so, aa, seqlen


class VPRegressor(nn.Module):
    """pare.py:24-91 - the regressor object GRNet owns (lib/models/grnet.py:82-85,171)."""

    def __init__(self, smpl_data: dict, focal_length=5000., img_res=224):
        super().__init__()
        self.smpl = SMPLHead(smpl_data, focal_length=focal_length, img_res=img_res)

    def get_body_joints(self, patt_output, batch_size=1, J_regressor=None):
        """pare.py:38-50 as written calls SMPLHead with SMPL keywords and cannot run; restated
        with the evident intent: identity root orientation, joints flattened per frame."""
        pose = patt_output['pred_pose']
        rot = pose.clone()
        rot[:, 0] = torch.eye(3)
        so = self.smpl.smpl(betas=patt_output['pred_shape'], body_pose=rot[:, 1:],
                            global_orient=rot[:, 0].unsqueeze(1), pose2rot=False)
        return so.joints.reshape(pose.shape[0], pose.shape[1], -1)

    def forward(self, patt_output, batch_size=1, J_regressor=None):
        so, aa, seqlen = _smpl_stage(self.smpl, patt_output['pred_pose'], patt_output['pred_shape'],
                                     patt_output['pred_cam'], batch_size, J_regressor)
        rotmat = patt_output['pred_pose']
        out = [{
            'theta': torch.cat([patt_output['pred_cam'], aa, patt_output['pred_shape']], dim=1).reshape(batch_size, seqlen, -1),
            'verts': so['smpl_vertices'].reshape(batch_size, seqlen, -1, 3),
            'kp_2d': so['smpl_joints2d'].reshape(batch_size, seqlen, -1, 2),
            'kp_3d': so['smpl_joints3d'].reshape(batch_size, seqlen, -1, 3),
            'rotmat': rotmat.reshape(batch_size, seqlen, -1, 3, 3),
        }]
        if 'pred_avg' in patt_output and 'pred_phase' in patt_output:
            out[-1].update({'pred_avg': patt_output['pred_avg'], 'pred_phase': patt_output['pred_phase']})
        return out


class SMPLRegressor(nn.Module):
    """pare.py:93-142 - same stage keyed on 'pred_rotmat', returning a plain dict."""

    def __init__(self, smpl_data: dict, focal_length=5000., img_res=224):
        super().__init__()
        self.smpl = SMPLHead(smpl_data, focal_length=focal_length, img_res=img_res)

    def forward(self, patt_output, batch_size=1, J_regressor=None):
        so, aa, seqlen = _smpl_stage(self.smpl, patt_output['pred_rotmat'], patt_output['pred_shape'],
                                     patt_output['pred_cam'], batch_size, J_regressor)
        return {
            'kp_2d': so['smpl_joints2d'].reshape(batch_size, seqlen, -1, 2),
            'kp_3d': so['smpl_joints3d'].reshape(batch_size, seqlen, -1, 3),
            'rotmat': patt_output['pred_rotmat'].reshape(batch_size, seqlen, -1, 3, 3),
            'verts': so['smpl_vertices'].reshape(batch_size, seqlen, -1, 3),
        }
