"""The final-prediction part of lib/models/pare.py::PareHead: the step of the live MAX-GRNet model
immediately upstream of the regression path - it turns the convolutional feature maps into pred_pose / pred_shape / pred_cam.

  * _get_local_feats        pare.py:318-326  smpl_final_layer (1x1 conv) + KeypointAttention pooling of both feature maps
  * _pare_get_final_preds   pare.py:328-375  per-joint pose MLP (LocallyConnected2d), shape / cam MLPs, residual iterations
  * forward                 pare.py:261-289  init handling, rot6d -> rotation matrices, output dict

Parameter names match PareHead's (pose_mlp.weight, shape_mlp.{weight,bias}, cam_mlp.{weight,bias}, smpl_final_layer.{weight,
bias}, init_pose/shape/cam), so `load_state_dict(pare_head.state_dict(), strict=False)` takes a trained head.  The
convolutional branches of PareHead (deconv layers, keypoint_final_layer) are out of scope and stay in the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import geometry as G
from .layers import KeypointAttention, LocallyConnected2d, _linear


class PareFinalHead(nn.Module):
    def __init__(self, num_joints=24, num_features_pare=128, num_features_smpl=64, num_camera_params=3, mean_params=None,
                 iterative_regression=False):
        super().__init__()
        self.num_joints, self.num_iterations, self.iterative_regression = num_joints, 1, iterative_regression   # pare.py:174
        self.pose_mlp_inp_dim = num_features_pare
        self.shape_mlp_inp_dim = num_joints * num_features_smpl
        self.smpl_final_layer = nn.Conv2d(num_features_pare, num_features_smpl, kernel_size=1)                  # pare.py:206-212
        self.shape_mlp = nn.Linear(self.shape_mlp_inp_dim, 10)                                                   # pare.py:226 (1 layer)
        self.cam_mlp = nn.Linear(self.shape_mlp_inp_dim, num_camera_params)
        self.pose_mlp = LocallyConnected2d(in_channels=num_features_pare, out_channels=6, output_size=[num_joints, 1],
                                           kernel_size=1, stride=1)                                              # pare.py:422-430
        self.keypoint_attention = KeypointAttention(use_conv=False, in_channels=(num_features_pare, num_features_smpl),
                                                    out_channels=(num_features_pare, num_features_smpl), act='softmax', use_scale=False)
        if mean_params is None:
            mean_params = {"pose": np.tile(np.array([1., 0, 0, 1, 0, 0], np.float32), num_joints), "shape": np.zeros(10, np.float32),
                           "cam": np.array([0.9, 0, 0], np.float32)}
        self.register_buffer('init_pose', torch.as_tensor(np.asarray(mean_params['pose'][:]), dtype=torch.float32).unsqueeze(0))
        self.register_buffer('init_shape', torch.as_tensor(np.asarray(mean_params['shape'][:]), dtype=torch.float32).unsqueeze(0))
        self.register_buffer('init_cam', torch.as_tensor(np.asarray(mean_params['cam']), dtype=torch.float32).unsqueeze(0))

    @torch.no_grad()
    def _get_local_feats(self, smpl_feats, part_attention, output=None):
        """pare.py:318-326.  The 1x1 smpl_final_layer is applied AFTER the attention pooling instead of before it:
        pooling is linear and the softmax weights of a joint sum to one, so conv(pool(x)) == pool(conv(x)) (bias
        included) in exact arithmetic, and the (B,64,H,W) map is never materialised."""
        point_local_feat = self.keypoint_attention(smpl_feats, part_attention)            # (B, C, J)
        B, C, J = point_local_feat.shape
        w = L.f32(self.smpl_final_layer.weight.detach().reshape(-1, C), "smpl_final_layer.weight")
        b = L.f32(self.smpl_final_layer.bias.detach(), "smpl_final_layer.bias")
        O = w.shape[0]
        cam_shape_feats = torch.empty(B, O, J, device=point_local_feat.device)
        L.call("gait_locally_connected", L.ptr(point_local_feat), C * J, J, 1, L.ptr(w), C, 1, 0, L.ptr(b), 1, 0,
               L.ptr(cam_shape_feats), O * J, J, 1, None, None, B, C, O, J, L.stream_ptr())
        return point_local_feat, cam_shape_feats

    @torch.no_grad()
    def _pare_get_final_preds(self, pose_feats, cam_shape_feats, init_pose, init_shape, init_cam, iter_now=True):
        """pare.py:328-375 -> pred_pose (N,J,6), pred_shape (N,10), pred_cam (N,3)."""
        N = pose_feats.shape[0]
        J = self.num_joints
        shape_feats = L.f32(cam_shape_feats, "cam_shape_feats").reshape(N, -1)
        x = L.f32(pose_feats, "pose_feats").reshape(N, -1, J, 1)
        if self.iterative_regression and iter_now:
            if init_pose.shape[-1] != 6:
                init_pose = init_pose.reshape(N, 6, -1).transpose(2, 1)                   # mean pose: (N, 6*J) viewed (N,6,J)
            pred_pose = L.f32(init_pose.expand(N, J, 6) if init_pose.dim() == 3 else init_pose, "init_pose")
            pred_shape, pred_cam = init_shape.expand(N, -1), init_cam.expand(N, -1)
            for _ in range(self.num_iterations):
                _, pred_pose = self.pose_mlp.run(x, resid=pred_pose, out_layout="NJO")   # residual fused
                pred_cam = _linear(shape_feats, self.cam_mlp) + pred_cam
                pred_shape = _linear(shape_feats, self.shape_mlp) + pred_shape
        else:
            pred_pose = self.pose_mlp.run(x, out_layout="NJO")
            pred_cam = _linear(shape_feats, self.cam_mlp)
            pred_shape = _linear(shape_feats, self.shape_mlp)
        return pred_pose, pred_shape, pred_cam

    @torch.no_grad()
    def forward(self, point_local_feat, cam_shape_feats, output, inits=None, gt_segm=None):
        """pare.py:261-289."""
        batch_size = point_local_feat.shape[0]
        if inits is None:
            init_pose, init_shape = self.init_pose.expand(batch_size, -1), self.init_shape.expand(batch_size, -1)
            init_cam, iter_now = self.init_cam.expand(batch_size, -1), False
        else:
            init_pose, init_shape, init_cam, iter_now = inits['pred_rot6d'], inits['pred_shape'], inits['pred_cam'], True
        pred_pose, pred_shape, pred_cam = self._pare_get_final_preds(point_local_feat, cam_shape_feats, init_pose, init_shape,
                                                                     init_cam, iter_now=iter_now)
        pred_rotmat = G.rot6d_to_rotmat(pred_pose).reshape(batch_size, 24, 3, 3)
        output.update({'pred_rotmat': pred_rotmat, 'pred_cam': pred_cam, 'pred_shape': pred_shape, 'pred_rot6d': pred_pose,
                       'pred_pose': pred_rotmat})
        return output
