"""Mirror of spin.Regressor (lib/models/spin.py:210-295) and pare.VPRegressor / SMPLRegressor
(lib/models/pare.py:24-142) on the sm_100a kernels.

Same constructor keywords, forward signatures, output dict keys/shapes (list of one dict for
Regressor / VPRegressor) and state_dict keys (fc1/fc2/decpose/decshape/deccam, init_* buffers,
smpl.* buffers).  Inference only: the kernels implement eval() semantics (dropout = identity).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import geometry as G
from .smpl import SMPL, SMPLHead, SMPL_MEAN_PARAMS, SMPL_MODEL_DIR, H36M_TO_J14

_STATE = 157        # 144 pose6d + 10 betas + 3 cam
_STATE_LD = 160


def _regress_joints(J_regressor, verts, subset14):
    """pare.py:70-76 / spin.py:279-282: (J',V).(F,V,3), then optional H36M_TO_J14 pick."""
    Jr = L.f32(J_regressor.to(verts.device), "J_regressor")
    out = L.joint_regress(verts, Jr)
    if subset14:
        out = out[:, H36M_TO_J14, :]
    return out


class Regressor(nn.Module):
    """spin.py:210-295 - 3x [x | pose6d | betas | cam] -> fc1 -> fc2 -> {decpose, decshape, deccam}
    residual updates, rot6d -> R, SMPL, projection, R -> axis-angle."""

    def __init__(self, smpl_mean_params=SMPL_MEAN_PARAMS, smpl_model_dir=SMPL_MODEL_DIR):
        super().__init__()
        npose = 24 * 6
        self.fc1 = nn.Linear(512 * 4 + npose + 13, 1024)
        self.drop1 = nn.Dropout()
        self.fc2 = nn.Linear(1024, 1024)
        self.drop2 = nn.Dropout()
        self.decpose = nn.Linear(1024, npose)
        self.decshape = nn.Linear(1024, 10)
        self.deccam = nn.Linear(1024, 3)
        nn.init.xavier_uniform_(self.decpose.weight, gain=0.01)
        nn.init.xavier_uniform_(self.decshape.weight, gain=0.01)
        nn.init.xavier_uniform_(self.deccam.weight, gain=0.01)
        self.smpl = SMPL(smpl_model_dir, batch_size=64, create_transl=False)
        mean_params = smpl_mean_params if isinstance(smpl_mean_params, dict) else np.load(smpl_mean_params)
        self.register_buffer('init_pose', torch.from_numpy(np.asarray(mean_params['pose'][:], dtype=np.float32)).unsqueeze(0))
        self.register_buffer('init_shape', torch.from_numpy(np.asarray(mean_params['shape'][:], dtype=np.float32)).unsqueeze(0))
        self.register_buffer('init_cam', torch.from_numpy(np.asarray(mean_params['cam'], dtype=np.float32)).unsqueeze(0))
        self._packed = None
        self._packed_key = None
        self._folded = {}

    # ------------------------------------------------------------ packed weights
    def _prepare(self):
        ps = (self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.decpose.weight, self.decpose.bias,
              self.decshape.weight, self.decshape.bias, self.deccam.weight, self.deccam.bias,
              self.init_pose, self.init_shape, self.init_cam)
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.fc1.weight.device
        if dev.type != "cuda":
            raise L.GaitLibraryError("Regressor weights are on %s; move the module to a CUDA device (no CPU path)" % dev)
        with torch.no_grad():
            din = self.fc1.in_features - _STATE
            dh = self.fc1.out_features
            w1 = self.fc1.weight.detach().float()
            W1s = torch.zeros(dh, _STATE_LD, device=dev)
            W1s[:, :_STATE] = w1[:, din:]
            init = torch.zeros(1, _STATE_LD, device=dev)
            init[:, :_STATE] = torch.cat([self.init_pose, self.init_shape, self.init_cam], dim=1).float()
            self._packed = {
                "din": din, "dh": dh,
                "W1x": w1[:, :din].contiguous(), "W1s": W1s, "b1": self.fc1.bias.detach().float().contiguous(),
                "W2": self.fc2.weight.detach().float().contiguous(), "b2": self.fc2.bias.detach().float().contiguous(),
                "Wd": torch.cat([self.decpose.weight, self.decshape.weight, self.deccam.weight], 0).detach().float().contiguous(),
                "bd": torch.cat([self.decpose.bias, self.decshape.bias, self.deccam.bias], 0).detach().float().contiguous(),
                "init": init,
            }
        for name in ("W1x", "W1s", "W2", "Wd"):      # constant GEMM weights: TF32 lo parts split off once
            L.prepare_weight(self._packed[name])
        self._packed_key = key
        return self._packed

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._folded = {}
        r = super()._apply(fn, *a, **k)
        L.purge_prepared()                       # prepared-weight entries of tensors that moved (.to / .cuda / .float)
        return r

    def fold(self, n_iter=3):
        """Opt-in weight folding.  spin.py:244-265 puts no non-linearity between fc1, fc2 and the decoders, and dropout is the
        identity in eval(), so with the shared mean-parameter init the loop is one affine map of x:
            u = A x + c,  s_{k+1} = (I + B) s_k + u   =>   s_n = G A x + (G c + (I + B)^n s_0),  G = sum_{k<n} (I + B)^k
        with D = Wd W2, A = D W1x, B = D W1s, c = D b1 + Wd b2 + bd.  The products are formed once in FP64 and rounded to FP32;
        returns {"Wf": (157,Din), "bf": (157,)} (re-folded when a weight changes)."""
        pk = self._prepare()
        key = (self._packed_key, int(n_iter))
        hit = self._folded.get("key")
        if hit == key:
            return self._folded
        with torch.no_grad():
            d = lambda t: t.double()
            W1x, W1s, W2, Wd = d(pk["W1x"]), d(pk["W1s"][:, :_STATE]), d(pk["W2"]), d(pk["Wd"])
            D = Wd @ W2
            A, B = D @ W1x, D @ W1s
            c = D @ d(pk["b1"]) + Wd @ d(pk["b2"]) + d(pk["bd"])
            T = torch.eye(_STATE, dtype=torch.float64, device=A.device) + B
            G = torch.zeros_like(T)
            P = torch.eye(_STATE, dtype=torch.float64, device=A.device)
            for _ in range(int(n_iter)):
                G += P
                P = T @ P
            s0 = d(pk["init"][0, :_STATE])
            self._folded = {"key": key, "Wf": (G @ A).float().contiguous(), "bf": (G @ c + P @ s0).float().contiguous()}
        L.prepare_weight(self._folded["Wf"])
        return self._folded

    def iterate_folded(self, x, n_iter=3):
        """iterate() for the shared init through the folded affine map (one GEMM): same state up to FP32 rounding."""
        if self.training:
            raise L.GaitLibraryError("Regressor kernels implement eval() semantics (dropout = identity); call .eval()")
        fk, pk = self.fold(n_iter), self._prepare()
        x = L.f32(x, "x")
        if x.dim() != 2 or x.shape[1] != pk["din"]:
            raise ValueError(f"x must be (N,{pk['din']}), got {tuple(x.shape)}")
        F = x.shape[0]
        state = torch.empty(F, _STATE_LD, device=x.device)
        nbytes = L.load().gait_hmr_folded_workspace_bytes(F)
        ws = torch.empty(max(nbytes, 4) // 4, device=x.device)
        L.call("gait_hmr_regressor_folded", L.ptr(x), x.stride(0), L.ptr(fk["Wf"]), L.ptr(fk["bf"]), L.ptr(state), F,
               pk["din"], L.ptr(ws), nbytes, L.stream_ptr())
        return state

    def iterate(self, x, init_pose=None, init_shape=None, init_cam=None, n_iter=3):
        """spin.py:244-265 - the MLP loop alone -> state (F,160) = [pose6d 144 | betas 10 | cam 3 | pad]."""
        if self.training:
            raise L.GaitLibraryError("Regressor kernels implement eval() semantics (dropout = identity); call .eval()")
        pk = self._prepare()
        x = L.f32(x, "x")
        if x.dim() != 2 or x.shape[1] != pk["din"]:
            raise ValueError(f"x must be (N,{pk['din']}), got {tuple(x.shape)}")
        F = x.shape[0]
        dev = x.device
        if init_pose is None and init_shape is None and init_cam is None:
            init, rows = pk["init"], 1
        else:
            e = lambda t, d: (d.expand(F, -1) if t is None else L.f32(t, "init").expand(F, -1))
            init = torch.zeros(F, _STATE_LD, device=dev)
            init[:, :_STATE] = torch.cat([e(init_pose, self.init_pose), e(init_shape, self.init_shape),
                                          e(init_cam, self.init_cam)], dim=1)
            rows = F
        state = torch.empty(F, _STATE_LD, device=dev)
        nbytes = L.load().gait_hmr_workspace_bytes(F, pk["dh"])
        ws = torch.empty(max(nbytes, 4) // 4, device=dev)
        L.call("gait_hmr_regressor", L.ptr(x), x.stride(0), L.ptr(pk["W1x"]), L.ptr(pk["W1s"]), L.ptr(pk["b1"]),
               L.ptr(pk["W2"]), L.ptr(pk["b2"]), L.ptr(pk["Wd"]), L.ptr(pk["bd"]), L.ptr(init), rows, int(n_iter),
               L.ptr(state), F, pk["din"], pk["dh"], L.ptr(ws), nbytes, L.stream_ptr())
        return state

    @torch.no_grad()
    def forward(self, x, init_pose=None, init_shape=None, init_cam=None, n_iter=3, J_regressor=None):
        batch_size = x.shape[0]
        state = self.iterate(x, init_pose, init_shape, init_cam, n_iter)
        pred_pose, pred_shape, pred_cam = state[:, :144], state[:, 144:154], state[:, 154:157]
        pred_rotmat = G.rot6d_to_rotmat(pred_pose).view(batch_size, 24, 3, 3)
        if J_regressor is None:
            # projection() = SMPLHead projection with focal 5000 / res 224, divided by 112 (geometry.py:412-425)
            res = self.smpl.run(pred_rotmat, pred_shape, cam=pred_cam, focal_length=5000., img_res=224.,
                                kp2d_divisor=224. / 2.)
            pred_vertices, pred_joints, kp2d = res["vertices"], res["joints"], res["joints2d"]
        else:
            res = self.smpl.run(pred_rotmat, pred_shape)
            pred_vertices = res["vertices"]
            pred_joints = _regress_joints(J_regressor, pred_vertices, subset14=True)
            kp2d = G.projection(pred_joints, pred_cam)
        theta = torch.empty(batch_size, 85, device=x.device)
        L.call("gait_pack_theta", L.ptr(pred_rotmat), L.ptr(pred_cam), state.stride(0), L.ptr(pred_shape),
               state.stride(0), L.ptr(theta), batch_size, L.stream_ptr())
        return [{'theta': theta, 'verts': pred_vertices, 'kp_2d': kp2d, 'kp_3d': pred_joints, 'rotmat': pred_rotmat}]


def _smpl_stage(head: SMPLHead, rotmat, shape, cam, batch_size, J_regressor):
    """Shared body of pare.py:52-76 and pare.py:108-131."""
    rotmat = L.f32(rotmat, "pred_pose")
    shape = L.f32(shape, "pred_shape")
    cam = L.f32(cam, "pred_cam")
    so = head(rotmat=rotmat, shape=shape, cam=cam, normalize_joints2d=True)
    F = rotmat.shape[0]
    seqlen = int(F / batch_size)
    theta = torch.empty(F, 85, device=rotmat.device)
    L.call("gait_pack_theta", L.ptr(rotmat), L.ptr(cam), cam.stride(0), L.ptr(shape), shape.stride(0), L.ptr(theta), F,
           L.stream_ptr())
    if J_regressor is not None:
        v = so['smpl_vertices'].reshape(batch_size * seqlen, -1, 3)
        so['smpl_joints3d'] = _regress_joints(J_regressor, v, subset14=J_regressor.shape[0] < 24)
    return so, theta, seqlen


class VPRegressor(nn.Module):
    """pare.py:24-91 - the regressor object GRNet owns (lib/models/grnet.py:82-85,171)."""

    def __init__(self, focal_length=5000., img_res=224, smpl_model_dir=SMPL_MODEL_DIR):
        super().__init__()
        self.smpl = SMPLHead(focal_length=focal_length, img_res=img_res, smpl_model_dir=smpl_model_dir)

    @torch.no_grad()
    def get_body_joints(self, patt_output, batch_size=1, J_regressor=None):
        """pare.py:38-50 (identity root orientation, joints flattened per frame).  As written the
        reference passes SMPL keywords to SMPLHead and cannot run; this keeps its evident intent."""
        pose = L.f32(patt_output['pred_pose'], "pred_pose")
        rot = pose.clone()
        rot[:, 0] = torch.eye(3, device=pose.device)
        res = self.smpl.smpl.run(rot, patt_output['pred_shape'])
        return res["joints"].reshape(pose.shape[0], pose.shape[1], -1)

    @torch.no_grad()
    def forward(self, patt_output, batch_size=1, J_regressor=None):
        so, theta, seqlen = _smpl_stage(self.smpl, patt_output['pred_pose'], patt_output['pred_shape'],
                                        patt_output['pred_cam'], batch_size, J_regressor)
        output = [{
            'theta': theta.reshape(batch_size, seqlen, -1),
            'verts': so['smpl_vertices'].reshape(batch_size, seqlen, -1, 3),
            'kp_2d': so['smpl_joints2d'].reshape(batch_size, seqlen, -1, 2),
            'kp_3d': so['smpl_joints3d'].reshape(batch_size, seqlen, -1, 3),
            'rotmat': patt_output['pred_pose'].reshape(batch_size, seqlen, -1, 3, 3),
        }]
        if 'pred_avg' in patt_output.keys() and 'pred_phase' in patt_output.keys():
            output[-1].update({'pred_avg': patt_output['pred_avg'], 'pred_phase': patt_output['pred_phase']})
        return output


class SMPLRegressor(nn.Module):
    """pare.py:93-142 - same stage keyed on 'pred_rotmat', returning a plain dict."""

    def __init__(self, focal_length=5000., img_res=224, smpl_model_dir=SMPL_MODEL_DIR):
        super().__init__()
        self.smpl = SMPLHead(focal_length=focal_length, img_res=img_res, smpl_model_dir=smpl_model_dir)

    @torch.no_grad()
    def forward(self, patt_output, batch_size=1, J_regressor=None):
        so, _, seqlen = _smpl_stage(self.smpl, patt_output['pred_rotmat'], patt_output['pred_shape'],
                                    patt_output['pred_cam'], batch_size, J_regressor)
        return {
            'kp_2d': so['smpl_joints2d'].reshape(batch_size, seqlen, -1, 2),
            'kp_3d': so['smpl_joints3d'].reshape(batch_size, seqlen, -1, 3),
            'rotmat': patt_output['pred_rotmat'].reshape(batch_size, seqlen, -1, 3, 3),
            'verts': so['smpl_vertices'].reshape(batch_size, seqlen, -1, 3),
        }
