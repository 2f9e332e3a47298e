#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code in this container.

Run here (CPU container, /root/reference mounted):  python tests/golden/make_golden.py
Nothing in tests/, smoke() or bench.py reads /root/reference at run time - only this script.

What runs unmodified from /root/reference:
  * lib/utils/geometry.py and lib/data_utils/kp_utils.py (import cleanly);
  * lib/models/smpl.py, lib/models/spin.py (Regressor), lib/models/pare.py (VPRegressor,
    SMPLRegressor) - executed from their source files with STUB modules standing in for imports
    that are missing here: `smplx` (-> oracle.smplx_lbs, the restated third-party arithmetic),
    `turtle` (needs tkinter; the reference only does `from turtle import forward` and never
    uses it), `yacs` (a 10-line attribute-dict CfgNode), and the constant
    lib.core.config.VIBE_DATA_DIR that spin.py:12 imports but config.py never defines.
    The package __init__ files (lib/models/__init__.py pulls HRNet->yacs, layers/__init__ pulls
    timm) are bypassed by registering empty package modules first.
  So the goldens pin every line of the reference's wrappers (joint selection, projection,
  theta packing, J_regressor branch, MLP loop); the smplx lbs arithmetic underneath is the
  oracle's restatement and stays "parity unpinned" (see oracle/__init__.py).

The synthetic SMPL data / weights come from gaitb200.synthetic (seeded); the goldens store a
checksum of them so drift in the generator is detected.
"""
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))

from gaitb200 import synthetic  # noqa: E402
from oracle import smplx_lbs  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(1)


def checksum(arrs: dict) -> float:
    return float(sum(np.abs(np.asarray(v, dtype=np.float64)).sum() for v in arrs.values()))


# --------------------------------------------------------------------------- geometry
def gen_geometry():
    import lib.utils.geometry as rg
    g = torch.Generator().manual_seed(11)
    ident6 = torch.tensor([1., 0, 0, 1, 0, 0])
    rot6d = torch.cat([
        ident6 + 0.5 * torch.randn(48, 6, generator=g),
        torch.randn(12, 6, generator=g) * 3.0,
        ident6[None],
        torch.zeros(1, 6),                                   # both columns zero (eps clamp)
        torch.tensor([[1., 2., 0., 0., 0., 0.]]),            # a2 parallel to a1
        torch.tensor([[1e-7, 0., 0., 1e-7, 0., 0.]]),        # below eps
    ])
    R = rg.rot6d_to_rotmat(rot6d)
    # rotations that exercise all four quaternion branches + near-pi + identity
    aa = torch.cat([
        torch.randn(40, 3, generator=g) * 1.2,
        torch.tensor([[np.pi, 0, 0], [0, np.pi, 0], [0, 0, np.pi], [0, 0, 0], [1e-5, 0, 0],
                      [3.1, 0.1, 0.0], [0.1, 3.1, 0.0], [0.0, 0.1, 3.1], [2.2, 2.2, 0.0],
                      [0, 2.9, 1.0], [-3.0, 0.2, 0.4], [1e-9, -1e-9, 1e-9]], dtype=torch.float32),
    ])
    Rr = rg.batch_rodrigues(aa).view(-1, 3, 3)
    Rall = torch.cat([R[:40], Rr, torch.eye(3)[None], torch.diag(torch.tensor([1., -1., -1.]))[None],
                      torch.diag(torch.tensor([-1., 1., -1.]))[None], torch.diag(torch.tensor([-1., -1., 1.]))[None]])
    quat = rg.rotation_matrix_to_quaternion(Rall)
    qin = torch.cat([quat, torch.tensor([[1., 0, 0, 0], [-1., 0, 0, 0], [0., 1, 0, 0], [-0.5, 0.5, -0.5, 0.5]]),
                     torch.randn(8, 4, generator=g)])
    pts = torch.randn(6, 29, 3, generator=g) * 0.5
    cam = torch.stack([0.5 + torch.rand(6, generator=g), 0.2 * torch.randn(6, generator=g),
                       0.2 * torch.randn(6, generator=g)], dim=1)
    rot = rg.rot6d_to_rotmat(ident6 + 0.3 * torch.randn(6, 6, generator=g))
    trans = torch.cat([0.3 * torch.randn(6, 2, generator=g), 20 + 5 * torch.rand(6, 1, generator=g)], dim=1)
    center = torch.randn(6, 2, generator=g) * 10
    np.savez_compressed(
        OUT / "geometry.npz",
        rot6d=rot6d.numpy(), rot6d_to_rotmat=R.numpy(),
        rot6d_to_rotmat_spin=rg.rot6d_to_rotmat_spin(rot6d[:60].clone()).numpy(),
        rotmat_to_rot6d=rg.rotmat_to_rot6d(R).numpy(),
        aa=aa.numpy(), batch_rodrigues=Rr.reshape(-1, 9).numpy(),
        Rall=Rall.numpy(), rotation_matrix_to_quaternion=quat.numpy(),
        rotation_matrix_to_angle_axis=rg.rotation_matrix_to_angle_axis(Rall).numpy(),
        qin=qin.numpy(), quaternion_to_angle_axis=rg.quaternion_to_angle_axis(qin).numpy(),
        quat2mat=rg.quat2mat(qin).numpy(),
        pts=pts.numpy(), cam=cam.numpy(), projection=rg.projection(pts, cam).numpy(),
        convert_weak_perspective_to_perspective=rg.convert_weak_perspective_to_perspective(cam).numpy(),
        cwp_1000_256=rg.convert_weak_perspective_to_perspective(cam, focal_length=1000., img_res=256).numpy(),
        pp_rot=rot.numpy(), pp_trans=trans.numpy(), pp_center=center.numpy(),
        perspective_projection=rg.perspective_projection(pts, rot, trans, 1234.5, center).numpy(),
    )
    print("geometry.npz written")


# --------------------------------------------------------------------------- kp_utils
def gen_kp_utils():
    import lib.data_utils.kp_utils as rk
    rng = np.random.default_rng(5)
    j = rng.standard_normal((7, 29, 3)).astype(np.float32)
    out = rk.convert_kps(j, 'spin2', 'kinectv2')
    src, dst = rk.get_spin2_joint_names(), rk.get_kinectv2_joint_names()
    gather = np.array([src.index(n) if n in src else -1 for n in dst])
    np.savez_compressed(OUT / "kp_utils.npz", joints=j, spin2_to_kinectv2=out, gather=gather,
                        spin2_names=np.array(src), kinectv2_names=np.array(dst),
                        kinectv2_skeleton=rk.get_kinectv2_skeleton())
    print("kp_utils.npz written; gather =", gather.tolist())


# --------------------------------------------------------------------------- model wrappers
class _CN(dict):
    """Minimal yacs.config.CfgNode stand-in (attribute dict) for lib/core/config.py:27-60."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return self

    def merge_from_file(self, f):
        pass


def _install_stubs(data_dir: Path):
    # turtle: the reference does `from turtle import forward` (spin.py:5, pare.py:5) and never uses it
    t = types.ModuleType("turtle"); t.forward = lambda *a, **k: None; sys.modules["turtle"] = t
    y = types.ModuleType("yacs"); yc = types.ModuleType("yacs.config"); yc.CfgNode = _CN
    y.config = yc; sys.modules["yacs"] = y; sys.modules["yacs.config"] = yc

    # smplx stand-in: constructor signature of smplx.SMPL as the reference calls it
    # (smpl.py:102 via spin.py:227-231, smpl.py:144): SMPL(model_path, batch_size=.., create_transl=False)
    class _SMPL(smplx_lbs.SMPLX_SMPL):
        def __init__(self, model_path, batch_size=1, create_transl=True, **kw):
            data = dict(np.load(Path(model_path) / "SMPL_NEUTRAL_synthetic.npz"))
            super().__init__(data, batch_size=batch_size)

        def forward(self, *a, get_skin=True, **kw):
            return super().forward(*a, **kw)

    sx = types.ModuleType("smplx"); sx.SMPL = _SMPL
    sxu = types.ModuleType("smplx.utils"); sxu.SMPLOutput = smplx_lbs.SMPLOutput; sxu.ModelOutput = smplx_lbs.SMPLOutput
    sxl = types.ModuleType("smplx.lbs"); sxl.vertices2joints = smplx_lbs.vertices2joints
    sx.utils, sx.lbs = sxu, sxl
    sys.modules.update({"smplx": sx, "smplx.utils": sxu, "smplx.lbs": sxl})

    import lib.core.config as rc
    rc.VIBE_DATA_DIR = "data/vibe_data"            # spin.py:12 imports it; config.py never defines it

    def pkg(name, path):
        m = types.ModuleType(name); m.__path__ = [str(path)]; sys.modules[name] = m; return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec); sys.modules[name] = m; spec.loader.exec_module(m); return m

    import lib  # noqa: F401  (namespace/regular package from /root/reference)
    models = pkg("lib.models", REF / "lib/models")
    layers = pkg("lib.models.layers", REF / "lib/models/layers")
    layers.LocallyConnected2d = load("lib.models.layers.locallyconnected2d",
                                     REF / "lib/models/layers/locallyconnected2d.py").LocallyConnected2d
    layers.KeypointAttention = load("lib.models.layers.keypoint_attention",
                                    REF / "lib/models/layers/keypoint_attention.py").KeypointAttention
    models.smpl = load("lib.models.smpl", REF / "lib/models/smpl.py")
    models.spin = load("lib.models.spin", REF / "lib/models/spin.py")
    models.pare = load("lib.models.pare", REF / "lib/models/pare.py")
    return models


def gen_models():
    smpl_data = synthetic.make_smpl_data(seed=0, variant="sparse")
    mean = synthetic.make_mean_params()
    reg_state = synthetic.make_regressor_state(seed=0, decoder_gain=0.3)   # larger gain: non-trivial poses
    with tempfile.TemporaryDirectory() as td:
        d = Path(td) / "data/smpl_data"
        d.mkdir(parents=True)
        np.savez(d / "SMPL_NEUTRAL_synthetic.npz", **smpl_data)
        np.save(d / "J_regressor_extra.npy", smpl_data["J_regressor_extra"])
        np.savez(d / "smpl_mean_params.npz", **mean)
        cwd = os.getcwd()
        os.chdir(td)                                   # SMPL_DATA_DIR is the relative 'data/smpl_data'
        try:
            m = _install_stubs(d)
            # tables (smpl.py:16-94)
            np.savez_compressed(
                OUT / "smpl_tables.npz",
                joint_names=np.array(m.smpl.JOINT_NAMES), joint_map_keys=np.array(list(m.smpl.JOINT_MAP.keys())),
                joint_map_vals=np.array(list(m.smpl.JOINT_MAP.values())),
                joint_map_49=np.array([m.smpl.JOINT_MAP[n] for n in m.smpl.JOINT_NAMES]),
                h36m_to_j17=np.array(m.smpl.H36M_TO_J17), h36m_to_j14=np.array(m.smpl.H36M_TO_J14))

            rot6d, betas, cam = synthetic.make_pose_inputs(4, seed=3, noise=0.4)
            import lib.utils.geometry as rg
            rotmat = rg.rot6d_to_rotmat(rot6d).view(4, 24, 3, 3)
            jh36m = torch.from_numpy(smpl_data["J_regressor_h36m"])
            out = {"rot6d": rot6d.numpy(), "betas": betas.numpy(), "cam": cam.numpy(), "rotmat": rotmat.numpy(),
                   "data_checksum": np.array(checksum(smpl_data))}

            # --- SMPL wrapper, both joint sets, pose2rot False/True (smpl.py:97-130)
            smpl = m.smpl.SMPL(str(d), batch_size=1, create_transl=False)
            with torch.no_grad():
                for kin in (True, False):
                    smpl.kinectv2 = kin
                    so = smpl(betas=betas[:2], body_pose=rotmat[:2, 1:], global_orient=rotmat[:2, 0:1], pose2rot=False)
                    tag = "kin" if kin else "spin"
                    out[f"smpl_{tag}_vertices"] = so.vertices.numpy()
                    out[f"smpl_{tag}_joints"] = so.joints.numpy()
                smpl.kinectv2 = True
                aa = 0.4 * torch.randn(2, 72, generator=torch.Generator().manual_seed(9))
                so = smpl(betas=betas[:2], body_pose=aa[:, 3:], global_orient=aa[:, :3], pose2rot=True)
                out["smpl_aa"] = aa.numpy()
                out["smpl_aa_vertices"] = so.vertices.numpy()
                out["smpl_aa_joints"] = so.joints.numpy()

                # --- SMPLHead (smpl.py:137-191)
                head = m.smpl.SMPLHead(focal_length=5000., img_res=224, smpl_model_dir=str(d))
                ho = head(rotmat[:2], betas[:2], cam=cam[:2], normalize_joints2d=True)
                out["head_joints2d_norm"] = ho["smpl_joints2d"].numpy()
                ho = head(rotmat[:2], betas[:2], cam=cam[:2], normalize_joints2d=False)
                out["head_joints2d"] = ho["smpl_joints2d"].numpy()
                out["head_joints3d"] = ho["smpl_joints3d"].numpy()
            np.savez_compressed(OUT / "smpl_wrapper.npz", **out)

            # --- VPRegressor / SMPLRegressor (pare.py:24-142), B=2, T=2
            vp = m.pare.VPRegressor()
            sr = m.pare.SMPLRegressor()
            patt = {"pred_pose": rotmat, "pred_shape": betas, "pred_cam": cam}
            vout = {"rotmat": rotmat.numpy(), "betas": betas.numpy(), "cam": cam.numpy(),
                    "data_checksum": np.array(checksum(smpl_data))}
            with torch.no_grad():
                o = vp(dict(patt), batch_size=2)[-1]
                for k, v in o.items():
                    vout[f"vp_{k}"] = v.numpy()
                o = vp(dict(patt), batch_size=2, J_regressor=jh36m)[-1]
                vout["vp_h36m_kp_3d"] = o["kp_3d"].numpy()
                o = sr({"pred_rotmat": rotmat, "pred_shape": betas, "pred_cam": cam}, batch_size=1)
                for k, v in o.items():
                    vout[f"sr_{k}"] = v.numpy()
            np.savez_compressed(OUT / "vpregressor.npz", **vout)

            # --- spin.Regressor (spin.py:210-295), F=3
            reg = m.spin.Regressor()
            missing = reg.load_state_dict(reg_state, strict=False)
            assert not missing.unexpected_keys, missing
            reg.eval()
            x = synthetic.make_features(1, 3, seed=77)[0]
            rout = {"x": x.numpy(), "state_checksum": np.array(checksum({k: v.numpy() for k, v in reg_state.items()})),
                    "data_checksum": np.array(checksum(smpl_data))}
            with torch.no_grad():
                o = reg(x)[-1]
                for k, v in o.items():
                    rout[f"reg_{k}"] = v.numpy()
                o = reg(x, n_iter=1)[-1]
                rout["reg_iter1_theta"] = o["theta"].numpy()
                o = reg(x, J_regressor=jh36m)[-1]
                rout["reg_h36m_kp_3d"] = o["kp_3d"].numpy()
                rout["reg_h36m_kp_2d"] = o["kp_2d"].numpy()
            np.savez_compressed(OUT / "regressor.npz", **rout)
        finally:
            os.chdir(cwd)
    print("smpl_tables.npz, smpl_wrapper.npz, vpregressor.npz, regressor.npz written")


if __name__ == "__main__":
    assert REF.exists(), "run in the container that mounts /root/reference"
    gen_geometry()
    gen_kp_utils()
    gen_models()
