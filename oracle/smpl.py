"""Oracle restatement of lib/models/smpl.py:16-191 (joint tables, SMPL, SMPLHead). Torch CPU FP32.

Test infrastructure: see oracle/__init__.py.  Pinned by tests/golden/smpl_wrapper.npz, produced
by executing the reference's own smpl.py on top of ``oracle.smplx_lbs`` (the smplx stand-in).
"""
import torch
import torch.nn as nn

from . import geometry as G
from .smplx_lbs import SMPLX_SMPL, SMPLOutput, vertices2joints

# smpl.py:16-36 - joint name -> index into [45 smplx joints | 9 J_regressor_extra joints]
JOINT_MAP = {
    'OP Nose': 24, 'OP Neck': 12, 'OP RShoulder': 17, 'OP RElbow': 19, 'OP RWrist': 21,
    'OP LShoulder': 16, 'OP LElbow': 18, 'OP LWrist': 20, 'OP MidHip': 0, 'OP RHip': 2,
    'OP RKnee': 5, 'OP RAnkle': 8, 'OP LHip': 1, 'OP LKnee': 4, 'OP LAnkle': 7, 'OP REye': 25,
    'OP LEye': 26, 'OP REar': 27, 'OP LEar': 28, 'OP LBigToe': 29, 'OP LSmallToe': 30,
    'OP LHeel': 31, 'OP RBigToe': 32, 'OP RSmallToe': 33, 'OP RHeel': 34, 'Right Ankle': 8,
    'Right Knee': 5, 'Right Hip': 45, 'Left Hip': 46, 'Left Knee': 4, 'Left Ankle': 7,
    'Right Wrist': 21, 'Right Elbow': 19, 'Right Shoulder': 17, 'Left Shoulder': 16,
    'Left Elbow': 18, 'Left Wrist': 20, 'Neck (LSP)': 47, 'Top of Head (LSP)': 48,
    'Pelvis (MPII)': 49, 'Thorax (MPII)': 50, 'Spine (H36M)': 51, 'Jaw (H36M)': 52,
    'Head (H36M)': 53, 'Nose': 24, 'Left Eye': 26, 'Right Eye': 25, 'Left Ear': 28,
    'Right Ear': 27, 'Left Foot': 10, 'Right Foot': 11, 'Left Thumb': 35, 'Right Thumb': 40,
}
# smpl.py:37-87 - the 49-joint "spin" order
JOINT_NAMES = [
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder', 'OP LElbow',
    'OP LWrist', 'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle', 'OP LHip', 'OP LKnee',
    'OP LAnkle', 'OP REye', 'OP LEye', 'OP REar', 'OP LEar', 'OP LBigToe', 'OP LSmallToe',
    'OP LHeel', 'OP RBigToe', 'OP RSmallToe', 'OP RHeel', 'Right Ankle', 'Right Knee',
    'Right Hip', 'Left Hip', 'Left Knee', 'Left Ankle', 'Right Wrist', 'Right Elbow',
    'Right Shoulder', 'Left Shoulder', 'Left Elbow', 'Left Wrist', 'Neck (LSP)',
    'Top of Head (LSP)', 'Pelvis (MPII)', 'Thorax (MPII)', 'Spine (H36M)', 'Jaw (H36M)',
    'Head (H36M)', 'Nose', 'Left Thumb', 'Right Thumb', 'Left Foot', 'Right Foot',
]
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]   # smpl.py:93
H36M_TO_J14 = H36M_TO_J17[:14]                                                # smpl.py:94


class SMPL(SMPLX_SMPL):
    """smpl.py:97-130 - adds J_regressor_extra joints; 29-joint 'spin2' set when kinectv2
    (24 SMPL joints + L thumb/middle + R thumb/middle landmark vertices + thorax), else the
    49-joint 'spin' set."""
    extra = True
    kinectv2 = True

    def __init__(self, data: dict, batch_size: int = 1):
        super().__init__(data, batch_size=batch_size)
        self.register_buffer('J_regressor_extra', torch.as_tensor(data['J_regressor_extra'], dtype=torch.float32))
        self.joint_map = torch.tensor([JOINT_MAP[n] for n in JOINT_NAMES], dtype=torch.long)

    def forward(self, *args, **kwargs):
        kwargs['get_skin'] = True
        out = super().forward(*args, **kwargs)
        joints = out.joints
        if self.extra:
            extra = vertices2joints(self.J_regressor_extra, out.vertices)
            if self.kinectv2:
                lh = out.joints[:, [35, 37], :]
                rh = out.joints[:, [40, 42], :]
                thorax = extra[:, JOINT_MAP['Thorax (MPII)'] - out.joints.shape[-2], None]
                joints = torch.cat([out.joints[:, :24], lh, rh, thorax], dim=1)
            else:
                joints = torch.cat([out.joints, extra], dim=1)[:, self.joint_map, :]
        return SMPLOutput(vertices=out.vertices, global_orient=out.global_orient, body_pose=out.body_pose,
                          joints=joints, betas=out.betas, full_pose=out.full_pose)


class SMPLHead(nn.Module):
    """smpl.py:137-191 - SMPL + weak-perspective -> perspective camera + projection."""

    def __init__(self, data: dict, focal_length=5000., img_res=224):
        super().__init__()
        self.smpl = SMPL(data)
        self.focal_length = focal_length
        self.img_res = img_res

    def forward(self, rotmat, shape, cam=None, normalize_joints2d=False):
        so = self.smpl(betas=shape, body_pose=rotmat[:, 1:].contiguous(),
                       global_orient=rotmat[:, 0].unsqueeze(1).contiguous(), pose2rot=False)
        out = {'smpl_vertices': so.vertices, 'smpl_joints3d': so.joints}
        if cam is not None:
            b = so.joints.shape[0]
            cam_t = G.convert_weak_perspective_to_perspective(cam, focal_length=self.focal_length, img_res=self.img_res)
            j2d = G.perspective_projection(so.joints, rotation=torch.eye(3).unsqueeze(0).expand(b, -1, -1),
                                           translation=cam_t, focal_length=self.focal_length,
                                           camera_center=torch.zeros(b, 2))
            if normalize_joints2d:
                j2d = j2d / (self.img_res / 2.)
            out['smpl_joints2d'] = j2d
        return out
