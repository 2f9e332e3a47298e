"""Mirror of lib/models/smpl.py (SMPL, SMPLHead, joint tables) on the sm_100a kernels.

``SMPL`` keeps the reference's constructor / forward keywords, class switches (``extra``,
``kinectv2``), ``SMPLOutput`` fields and the smplx buffer names so that checkpoints keyed
``...smpl.<buffer>`` load.  The arithmetic smplx==0.1.26 performs underneath
(lbs.py: blend shapes, vertices2joints, Rodrigues, batch_rigid_transform, skinning;
vertex_joint_selector.py) runs in hand-written CUDA kernels through include/gaitb200.h:

    pose chain (1 warp / frame)  ->  blend GEMM (tcgen05)  ->  LBS (tcgen05, thorax row fused)  ->  assembly/projection

There is no CPU path: inputs must be FP32 CUDA tensors.
"""
from __future__ import annotations

import os
from collections import namedtuple
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import geometry as G

SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose"])
SMPLOutput.__new__.__defaults__ = (None,) * 6
ModelOutput = SMPLOutput

# lib/models/smpl.py:16-36 - joint name -> index into [45 smplx joints | 9 J_regressor_extra joints]
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
# lib/models/smpl.py:37-87 - the 49-joint "spin" order
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
JOINT_IDS = {JOINT_NAMES[i]: i for i in range(len(JOINT_NAMES))}
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]   # smpl.py:93
H36M_TO_J14 = H36M_TO_J17[:14]                                                # smpl.py:94

SMPL_DATA_DIR = 'data/smpl_data'
SMPL_MODEL_DIR = SMPL_DATA_DIR
JOINT_REGRESSOR_TRAIN_EXTRA = os.path.join(SMPL_DATA_DIR, 'J_regressor_extra.npy')
SMPL_MEAN_PARAMS = os.path.join(SMPL_DATA_DIR, 'smpl_mean_params.npz')

NUM_JOINTS = 24
NUM_SMPLX_JOINTS = 45      # 24 + 21 landmark vertices (smplx VertexJointSelector)
_THORAX_ROW = JOINT_MAP['Thorax (MPII)'] - NUM_SMPLX_JOINTS          # 5
_KINECT_HAND_JOINTS = [35, 37, 40, 42]                                # smpl.py:115-116


def load_smpl_data(model_path) -> dict:
    """Accept a dict of arrays, an .npz file, the reference's own SMPL_NEUTRAL.pkl (lib/models/smpl.py:102 unpickles it through
    smplx + chumpy; here a stand-in unpickler reads the same file without either, scripts/convert_smpl_pkl.py), or a
    directory holding SMPL_NEUTRAL*.npz / SMPL_NEUTRAL.pkl (+ J_regressor_extra.npy, as in data/smpl_data/)."""
    if isinstance(model_path, dict):
        return model_path
    p = Path(model_path)

    def read(f):
        if f.suffix == ".pkl":
            import importlib.util
            spec = importlib.util.spec_from_file_location("_convert_smpl_pkl", Path(__file__).resolve().parent.parent / "scripts" / "convert_smpl_pkl.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.smpl_pkl_to_dict(f)
        return dict(np.load(f, allow_pickle=False))

    if p.is_dir():
        cands = sorted(p.glob("SMPL_NEUTRAL*.npz")) or sorted(p.glob("SMPL_NEUTRAL*.pkl"))
        if not cands:
            raise FileNotFoundError(f"no SMPL_NEUTRAL*.npz / SMPL_NEUTRAL*.pkl under {p}")
        data = read(cands[0])
        extra = p / "J_regressor_extra.npy"
        if "J_regressor_extra" not in data and extra.exists():
            data["J_regressor_extra"] = np.load(extra)
        return data
    data = read(p)
    extra = p.parent / "J_regressor_extra.npy"
    if "J_regressor_extra" not in data and extra.exists():
        data["J_regressor_extra"] = np.load(extra)
    return data


class _VertexJointSelector(nn.Module):
    """Holds smplx's `extra_joints_idxs` buffer under the same state_dict key."""

    def __init__(self, idxs):
        super().__init__()
        self.register_buffer('extra_joints_idxs', torch.as_tensor(np.asarray(idxs), dtype=torch.long))


class SMPL(nn.Module):
    """lib/models/smpl.py:97-130 over smplx.SMPL: vertices + joints.  With ``extra`` and
    ``kinectv2`` (the defaults) joints are the 29-joint 'spin2' set (24 SMPL joints, L/R thumb
    and middle-finger landmark vertices, MPII thorax); with ``kinectv2 = False`` the 49-joint
    'spin' set; with ``extra = False`` smplx's 45 joints."""
    extra = True
    kinectv2 = True

    def __init__(self, model_path=SMPL_MODEL_DIR, batch_size=1, create_transl=False, **kwargs):
        super().__init__()
        data = load_smpl_data(model_path)
        t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a), dtype=dt)
        self.batch_size = batch_size
        self.faces = np.asarray(data["faces"])
        self.register_buffer("faces_tensor", t(data["faces"], torch.long))
        self.betas = nn.Parameter(torch.zeros(batch_size, 10))
        self.global_orient = nn.Parameter(torch.zeros(batch_size, 3))
        self.body_pose = nn.Parameter(torch.zeros(batch_size, 69))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(batch_size, 3))
        self.register_buffer("v_template", t(data["v_template"]))
        self.register_buffer("shapedirs", t(data["shapedirs"])[:, :, :10].contiguous())
        self.register_buffer("J_regressor", t(data["J_regressor"]))
        self.register_buffer("posedirs", t(data["posedirs"]))
        self.register_buffer("parents", t(data["parents"], torch.long))
        self.register_buffer("lbs_weights", t(data["lbs_weights"]))
        self.vertex_joint_selector = _VertexJointSelector(data["landmark_verts"])
        if "J_regressor_extra" not in data:
            raise KeyError("SMPL data lacks 'J_regressor_extra' (J_regressor_extra.npy, smpl.py:104)")
        self.register_buffer('J_regressor_extra', t(data['J_regressor_extra']))
        self.joint_map = torch.tensor([JOINT_MAP[n] for n in JOINT_NAMES], dtype=torch.long)   # plain tensor, smpl.py:106
        self._packed = None
        self._packed_key = None

    # ---------------------------------------------------------------- packed device operands
    def _key(self):
        bufs = (self.v_template, self.shapedirs, self.J_regressor, self.posedirs, self.parents, self.lbs_weights,
                self.J_regressor_extra, self.vertex_joint_selector.extra_joints_idxs)
        return tuple((b.data_ptr(), b._version, str(b.device)) for b in bufs)

    def _prepare(self):
        """Derived operands, rebuilt when the buffers move or change:
        J_template/J_shapedirs (J_regressor folded through the shape blend - the joint regression
        is linear in betas), the packed blend basis (3V,224), int32 index tables."""
        key = self._key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.v_template.device
        if dev.type != "cuda":
            raise L.GaitLibraryError("SMPL buffers are on %s; move the module to a CUDA device (no CPU path)" % dev)
        L.require_device()
        V = self.v_template.shape[0]
        st = L.stream_ptr()
        Jreg = self.J_regressor.contiguous()
        vt = self.v_template.contiguous()
        J_template = torch.empty(1, NUM_JOINTS, 3, device=dev)
        L.call("gait_joint_regress", L.ptr(vt), L.ptr(Jreg), L.ptr(J_template), 1, V, NUM_JOINTS, st)
        sd = self.shapedirs.permute(2, 0, 1).contiguous()                  # (10, V, 3)
        Jsd = torch.empty(10, NUM_JOINTS, 3, device=dev)
        L.call("gait_joint_regress", L.ptr(sd), L.ptr(Jreg), L.ptr(Jsd), 10, V, NUM_JOINTS, st)
        basis_t = torch.zeros(3 * V, 224, device=dev)
        basis_t[:, :207] = self.posedirs.t()
        basis_t[:, 207:217] = self.shapedirs.reshape(3 * V, 10)
        basis_t[:, 217] = vt.reshape(-1)
        lbs_w = self.lbs_weights.contiguous()
        wpack = torch.empty(L.load().gait_smpl_lbs_pack_bytes(V) // 4, device=dev)
        L.call("gait_smpl_lbs_pack", L.ptr(lbs_w), L.ptr(wpack), V, st)
        lm = self.vertex_joint_selector.extra_joints_idxs.to(torch.int32).contiguous()
        n_lm = lm.numel()
        i32 = lambda xs: torch.tensor(list(xs), dtype=torch.int32, device=dev)
        self._packed = {
            "V": V,
            "J_template": J_template.reshape(NUM_JOINTS, 3),
            "J_shapedirs": Jsd.permute(1, 2, 0).contiguous(),             # (24,3,10)
            "basis_t": basis_t,
            "parents": self.parents.to(torch.int32).contiguous(),
            "lbs_weights": lbs_w,
            "lbs_wpack": wpack,                                           # tensor-core LBS operand (lbs_tc.cu)
            "vtiles": 4 * ((V + 127) // 128),   # partial sums of the fused regressor row (gait_smpl_lbs_jx_parts)
             "ldv": 384 * ((V + 127) // 128),  # padded v_posed row (TMA bulk rows)
            "landmarks": lm, "n_landmarks": n_lm,
            "extra_all": self.J_regressor_extra.contiguous(),
            "extra_thorax": self.J_regressor_extra[_THORAX_ROW:_THORAX_ROW + 1].contiguous(),
            # joint maps over virtual joints [24 chain | landmarks | extra rows given]
            "map_kinect": i32(list(range(24)) + _KINECT_HAND_JOINTS + [NUM_JOINTS + n_lm]),
            "map_spin": i32(int(j) for j in self.joint_map),
            "map_smplx": i32(range(NUM_JOINTS + n_lm)),
        }
        L.prepare_weight(self._packed["basis_t"])   # constant blend-GEMM operand: TF32 lo part split off once
        self._packed_key = key
        return self._packed

    def _prepare_reduced(self):
        """Operands of the joints-only path that never forms the mesh (gait_smpl_reduced_joints): the blend-basis rows of the
        landmark vertices, and the thorax row of J_regressor_extra folded through the skinning weights,
        P_j = sum_v jx[v] W[v,j] basis[3v..3v+2,:] (24 x 3 x 224) and s_j = sum_v jx[v] W[v,j] - both regressions run on this
        library's own joint-regression kernel.  red_basis (3 n_lm + 72 -> padded to a multiple of 4, 224)."""
        pk = self._prepare()
        if "red_basis" in pk:
            return pk
        dev, V, st = pk["basis_t"].device, pk["V"], L.stream_ptr()
        lm, n_lm = pk["landmarks"].long(), pk["n_landmarks"]
        rows = (lm[:, None] * 3 + torch.arange(3, device=dev)[None, :]).reshape(-1)
        jxw = (pk["extra_thorax"].reshape(V, 1) * pk["lbs_weights"]).t().contiguous()            # (24, V): jx[v] W[v,j]
        basis_fvc = pk["basis_t"].reshape(V, 3, 224).permute(2, 0, 1).contiguous()              # (224, V, 3)
        Pk = torch.empty(224, NUM_JOINTS, 3, device=dev)
        L.call("gait_joint_regress", L.ptr(basis_fvc), L.ptr(jxw), L.ptr(Pk), 224, V, NUM_JOINTS, st)
        ones = torch.ones(1, V, 3, device=dev)
        sj = torch.empty(1, NUM_JOINTS, 3, device=dev)
        L.call("gait_joint_regress", L.ptr(ones), L.ptr(jxw), L.ptr(sj), 1, V, NUM_JOINTS, st)
        n_rows = 3 * n_lm + 3 * NUM_JOINTS
        red = torch.zeros((n_rows + 3) // 4 * 4, 224, device=dev)
        red[:3 * n_lm] = pk["basis_t"][rows]
        red[3 * n_lm:n_rows] = Pk.permute(1, 2, 0).reshape(3 * NUM_JOINTS, 224)
        pk["red_basis"], pk["red_ld"] = red, red.shape[0]
        pk["red_s"] = sj[0, :, 0].contiguous()
        pk["lm_weights"] = pk["lbs_weights"][lm].contiguous()
        L.prepare_weight(pk["red_basis"])
        return pk

    def _apply(self, fn, *a, **k):
        self._packed = None
        r = super()._apply(fn, *a, **k)
        L.purge_prepared()                       # prepared-weight entries of tensors that moved (.to / .cuda / .float)
        return r

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    # ---------------------------------------------------------------- the SMPL stage
    def run(self, rotmat, betas, cam=None, focal_length=5000., img_res=224., kp2d_divisor=1.0, want_verts=True,
            gather=None):
        """rotmat (F,24,3,3), betas (F,10) -> dict(vertices, joints[, joints2d][, gathered]).
        One pose-chain, one blend GEMM, one LBS (+ one joint-regression launch in the 49-joint mode)
        and one assembly launch."""
        pk = self._prepare()
        R = L.f32(rotmat, "rotmat").reshape(-1, NUM_JOINTS, 3, 3)
        F = R.shape[0]
        betas = L.f32(betas, "betas")
        if betas.shape[0] != F:
            betas = betas.expand(F, -1).contiguous()
        dev, V, st = R.device, pk["V"], L.stream_ptr()
        Jp = torch.empty(F, NUM_JOINTS, 3, device=dev)
        coef = torch.empty(F, 224, device=dev)
        lib = L.load()
        aop = torch.empty(lib.gait_smpl_lbs_aop_bytes(F) // 4, device=dev)
        L.call("gait_smpl_pose_chain", L.ptr(R), L.ptr(betas), betas.stride(0), L.ptr(pk["J_template"]),
               L.ptr(pk["J_shapedirs"]), L.ptr(pk["parents"]), None, L.ptr(Jp), L.ptr(coef), L.ptr(aop), F, st)
        ldv, vtiles = pk["ldv"], pk["vtiles"]
        v_posed = torch.empty(F, ldv, device=dev)
        L.call("gait_smpl_blend", L.ptr(coef), L.ptr(pk["basis_t"]), L.ptr(v_posed), ldv, F, 3 * V, st)
        verts = torch.empty(F, V, 3, device=dev)
        extra, extra_parts, extra_stride = None, 1, 0
        if self.extra and self.kinectv2:
            # thorax row of J_regressor_extra fused into the skinning kernel as per-tile partial sums
            jmap = pk["map_kinect"]
            extra = torch.empty(vtiles, F, 1, 3, device=dev)
            extra_parts, extra_stride = vtiles, F * 3
            L.call("gait_smpl_lbs_tc", L.ptr(v_posed), ldv, L.ptr(aop), L.ptr(pk["lbs_wpack"]),
                   L.ptr(pk["extra_thorax"]), L.ptr(verts), L.ptr(extra), F, V, st)
        else:
            L.call("gait_smpl_lbs_tc", L.ptr(v_posed), ldv, L.ptr(aop), L.ptr(pk["lbs_wpack"]), None, L.ptr(verts),
                   None, F, V, st)
            if self.extra:
                Jx, jmap = pk["extra_all"], pk["map_spin"]
                extra = torch.empty(1, F, Jx.shape[0], 3, device=dev)
                L.joint_regress(verts, Jx, extra.view(F, Jx.shape[0], 3))          # 9 rows, one streaming pass (jreg.cu)
            else:
                jmap = pk["map_smplx"]
        J = jmap.numel()
        joints = torch.empty(F, J, 3, device=dev)
        kp2d = torch.empty(F, J, 2, device=dev) if cam is not None else None
        camc = None if cam is None else L.f32(cam, "cam")
        gidx = gat = None
        if gather is not None:
            gidx = gather.to(device=dev, dtype=torch.int32).contiguous()
            gat = torch.empty(F, gidx.numel(), 3, device=dev)
        L.call("gait_joints_assemble", L.ptr(Jp), L.ptr(verts), V, L.ptr(pk["landmarks"]), pk["n_landmarks"],
               L.ptr(extra), 0 if extra is None else extra.shape[2], extra_parts, extra_stride, L.ptr(jmap), J,
               L.ptr(joints), L.ptr(camc),
               0 if camc is None else camc.stride(0), float(focal_length), float(img_res), float(kp2d_divisor),
               L.ptr(kp2d), L.ptr(gidx), 0 if gidx is None else gidx.numel(), L.ptr(gat), F, st)
        out = {"vertices": verts, "joints": joints}
        if kp2d is not None:
            out["joints2d"] = kp2d
        if gat is not None:
            out["gathered"] = gat
        return out

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, **kwargs):
        """smplx.SMPL.forward keywords as the reference passes them (smpl.py:108-111;
        spin.py:269-274; smooth_pose.py:72-76).  pose2rot=False: body_pose (F,23,3,3),
        global_orient (F,1,3,3) rotation matrices; pose2rot=True: axis-angle (F,69)/(F,3)."""
        kwargs['get_skin'] = True
        global_orient = self.global_orient if global_orient is None else global_orient
        body_pose = self.body_pose if body_pose is None else body_pose
        betas = self.betas if betas is None else betas
        F = max(betas.shape[0], body_pose.shape[0])
        if pose2rot:
            go = global_orient.reshape(-1, 3)
            full = torch.cat([go, body_pose.reshape(go.shape[0], -1)], dim=1)
            R = G.batch_rodrigues_smplx(full.detach().reshape(-1, 3)).view(-1, NUM_JOINTS, 3, 3)
        else:
            R = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(-1, NUM_JOINTS - 1, 3, 3)], dim=1)
        if R.shape[0] != F:
            R = R.expand(F, -1, -1, -1)
        res = self.run(R.detach(), betas.detach())
        return SMPLOutput(vertices=res["vertices"], global_orient=global_orient, body_pose=body_pose,
                          joints=res["joints"], betas=betas, full_pose=None)


def get_smpl_faces(model_path=SMPL_MODEL_DIR):
    """lib/models/smpl.py:133-135."""
    return np.asarray(load_smpl_data(model_path)["faces"])


class SMPLHead(nn.Module):
    """lib/models/smpl.py:137-191 - SMPL + weak-perspective -> perspective camera + projection
    (the projection runs inside the joint-assembly kernel)."""

    def __init__(self, focal_length=5000., img_res=224, smpl_model_dir=SMPL_MODEL_DIR):
        super().__init__()
        self.smpl = SMPL(smpl_model_dir, create_transl=False)
        self.focal_length = focal_length
        self.img_res = img_res

    def forward(self, rotmat, shape, cam=None, normalize_joints2d=False):
        res = self.smpl.run(rotmat, shape, cam=cam, focal_length=self.focal_length, img_res=self.img_res,
                            kp2d_divisor=(self.img_res / 2.) if normalize_joints2d else 1.0)
        output = {'smpl_vertices': res["vertices"], 'smpl_joints3d': res["joints"]}
        if cam is not None:
            output['smpl_joints2d'] = res["joints2d"]
        return output
