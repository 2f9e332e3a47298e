"""Oracle restatement of the heads next to the regression path.  Torch CPU FP32.

  * locally_connected      lib/models/layers/locallyconnected2d.py:39-49
  * keypoint_attention     lib/models/layers/keypoint_attention.py:34-55 (use_conv=False, act='softmax')
  * pare_final             lib/models/pare.py:261-289, 318-375 (PareHead: local features + final-prediction MLPs)
  * bidirectional_model    lib/models/layers/gait_feat_encoder.py:80-104 (use_pareFeat=True)

Test infrastructure: see oracle/__init__.py.  Pinned by tests/golden/heads.npz, produced by executing the reference's own
classes (tests/golden/make_golden_heads.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import geometry as G


def locally_connected(x, weight, bias=None):
    """x (N,C,J,1), weight (1,O,C,J,1,1), bias (1,O,J,1) -> (N,O,J,1); kernel_size 1 so the unfolds are identities."""
    out = (x.unsqueeze(1).unsqueeze(-1) * weight).sum([2, -1])
    if bias is not None:
        out = out + bias
    return out


def keypoint_attention(features, heatmaps, use_scale=False):
    B, J, H, W = heatmaps.shape
    if use_scale:
        heatmaps = heatmaps * (1.0 / np.sqrt(H * W))
    nh = F.softmax(heatmaps.reshape(B, J, -1), dim=-1)
    feats = features.reshape(B, -1, H * W)
    return torch.matmul(nh, feats.transpose(2, 1)).transpose(2, 1)


def pare_final(sd, smpl_feats, part_attn, inits=None, iterative=False):
    """sd: state dict with pose_mlp.weight, shape_mlp.*, cam_mlp.*, smpl_final_layer.*, init_*.  Returns the output dict."""
    cam_shape_map = F.conv2d(smpl_feats, sd["smpl_final_layer.weight"], sd["smpl_final_layer.bias"])
    plf = keypoint_attention(smpl_feats, part_attn)
    csf = keypoint_attention(cam_shape_map, part_attn)
    N = plf.shape[0]
    pose_feats = plf.unsqueeze(-1)
    shape_feats = torch.flatten(csf, start_dim=1)
    if inits is None:
        init_pose, init_shape, init_cam, iter_now = sd["init_pose"].expand(N, -1), sd["init_shape"].expand(N, -1), sd["init_cam"].expand(N, -1), False
    else:
        init_pose, init_shape, init_cam, iter_now = inits["pred_rot6d"], inits["pred_shape"], inits["pred_cam"], True
    if init_pose.shape[-1] == 6:
        init_pose = init_pose.transpose(2, 1).unsqueeze(-1)
    else:
        init_pose = init_pose.reshape(N, 6, -1).unsqueeze(-1)
    lin = lambda x, k: F.linear(x, sd[k + ".weight"], sd[k + ".bias"])
    if iterative and iter_now:
        pred_pose = locally_connected(pose_feats, sd["pose_mlp.weight"]) + init_pose
        pred_cam = lin(shape_feats, "cam_mlp") + init_cam
        pred_shape = lin(shape_feats, "shape_mlp") + init_shape
    else:
        pred_pose = locally_connected(pose_feats, sd["pose_mlp.weight"])
        pred_cam, pred_shape = lin(shape_feats, "cam_mlp"), lin(shape_feats, "shape_mlp")
    pred_pose = pred_pose.squeeze(-1).transpose(2, 1)
    rotmat = G.rot6d_to_rotmat(pred_pose).reshape(N, 24, 3, 3)
    return {"point_local_feat": plf, "cam_shape_feats": csf, "pred_rotmat": rotmat, "pred_cam": pred_cam, "pred_shape": pred_shape,
            "pred_rot6d": pred_pose, "pred_pose": rotmat}


def bidirectional_model(sd, x, cparams, num_layers=2, h_size=300):
    """gait_feat_encoder.py:80-104 with eval-mode dropout.  Returns (y (B,3), p (B,T,4), xc (B,T,3072))."""
    b, n, cf = cparams.shape
    xc = locally_connected(cparams.reshape(b * n, cf, 1, 1).expand(b * n, cf, 24, 1), sd["cparam_mpl.weight"]).reshape(b, n, -1)
    x = x + xc
    rnn = torch.nn.GRU(input_size=x.shape[-1], hidden_size=h_size, num_layers=num_layers, batch_first=True, bidirectional=True)
    rnn.load_state_dict({k[len("rnn."):]: v for k, v in sd.items() if k.startswith("rnn.")})
    with torch.no_grad():
        xo, h = rnn(x)
    h = h.permute(1, 0, 2).reshape(b, -1)
    mlp = lambda z, name: F.linear(F.leaky_relu(F.linear(z, sd[name + ".0.weight"], sd[name + ".0.bias"]), 0.05),
                                   sd[name + ".2.weight"], sd[name + ".2.bias"])
    y = torch.cat((mlp(h, "speed_mlp"), mlp(h, "step_mlp")), dim=-1)
    p = torch.tanh(mlp(xo, "phase_mlp"))
    return y, p, xc
