#!/usr/bin/env python
"""Generate tests/golden/postproc.npz by running the REFERENCE's own post-processing code in this container:

  * lib/utils/one_euro_filter.py            (imports cleanly)                      -> One-Euro filter trajectories
  * lib/utils/smooth_pose.py::smooth_pose                           -> filtered poses, vertices, joints
  * lib/utils/demo_utils.py::convert_crop_cam_to_orig_img / convert_crop_coords_to_orig_img   

The modules run unmodified; third-party imports that are missing here and unused by these functions (matplotlib,
pytube, skimage) are replaced by empty stub modules, and `smplx` by the oracle's restatement exactly as in
make_golden.py (so the smplx arithmetic underneath stays "parity unpinned").  Run here:  python tests/golden/make_golden_post.py
"""
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402  (stubs for smplx / yacs / turtle, synthetic data)

from gaitb200 import synthetic  # noqa: E402


class _Stub(types.ModuleType):
    """Module whose every attribute is a dummy object / sub-stub (for imports the tested functions never use)."""
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return type(k, (), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("matplotlib", "pytube", "skimage")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def main():
    assert mg.REF.exists(), "run in the container that mounts /root/reference"
    sys.meta_path.insert(0, _StubFinder())
    out = {}
    rng = np.random.default_rng(21)

    # ---- One-Euro filter exactly as smooth_pose drives it (one_euro_filter.py:14-46, smooth_pose.py:51-56,84-88)
    from lib.utils.one_euro_filter import OneEuroFilter
    T = 48
    base = np.cumsum(rng.standard_normal((T, 24, 3)).astype(np.float32) * 0.05, axis=0)
    pose = (base + rng.standard_normal((T, 24, 3)).astype(np.float32) * 0.02).astype(np.float32)
    for tag, (mc, beta) in {"default": (0.004, 0.7), "stiff": (1.0, 0.0), "fast": (0.05, 3.0)}.items():
        f = OneEuroFilter(np.zeros_like(pose[0]), pose[0], min_cutoff=mc, beta=beta)
        hat = np.zeros_like(pose)
        hat[0] = pose[0]
        for i in range(1, T):
            hat[i] = f(np.ones_like(pose[i]) * i, pose[i])
        out[f"oef_{tag}"] = hat
        out[f"oef_{tag}_params"] = np.array([mc, beta])
    out["oef_in"] = pose

    # ---- crop -> image conversions (demo_utils.py:176-209)
    import lib.utils.demo_utils as du
    N = 9
    cam = np.stack([0.6 + 0.4 * rng.random(N), 0.2 * rng.standard_normal(N), 0.2 * rng.standard_normal(N)], 1).astype(np.float32)
    bbox = np.stack([300 + 200 * rng.random(N), 250 + 150 * rng.random(N), 180 + 120 * rng.random(N),
                     180 + 120 * rng.random(N)], 1)                                   # float64 (cx, cy, h, h) like the db
    kp = (rng.random((N, 29, 2)) * 2 - 1).astype(np.float32)
    out.update(cc_cam=cam, cc_bbox=bbox, cc_kp=kp,
               crop_cam_1280x720=du.convert_crop_cam_to_orig_img(cam, bbox, 1280, 720),
               crop_cam_f32=du.convert_crop_cam_to_orig_img(cam, bbox.astype(np.float32), 640, 480),
               crop_coords_224=du.convert_crop_coords_to_orig_img(bbox, kp.copy(), 224),
               crop_coords_f32=du.convert_crop_coords_to_orig_img(bbox.astype(np.float32), kp.copy(), 224))

    # ---- smooth_pose (smooth_pose.py:28-116) on the synthetic SMPL model, axis-angle and quaternion inputs
    smpl_data = synthetic.make_smpl_data(seed=0, variant="sparse")
    with tempfile.TemporaryDirectory() as td:
        d = Path(td) / "data/smpl_data"
        d.mkdir(parents=True)
        np.savez(d / "SMPL_NEUTRAL_synthetic.npz", **smpl_data)
        np.save(d / "J_regressor_extra.npy", smpl_data["J_regressor_extra"])
        np.savez(d / "smpl_mean_params.npz", **synthetic.make_mean_params())
        cwd = os.getcwd()
        os.chdir(td)
        try:
            mg._install_stubs(d)
            import lib.utils.smooth_pose as sp
            import lib.utils.geometry as rg
            Ts = 12
            aa = (0.3 * np.cumsum(rng.standard_normal((Ts, 72)) * 0.1, axis=0) + 0.2 * rng.standard_normal((1, 72))).astype(np.float32)
            betas = rng.standard_normal((Ts, 10)).astype(np.float32)
            for kin in (False, True):
                v, p, j = sp.smooth_pose(aa.copy(), betas, min_cutoff=0.004, beta=0.7, device="cpu", kinectv2=kin)
                tag = "kin" if kin else "spin"
                out[f"sp_{tag}_verts"], out[f"sp_{tag}_pose"], out[f"sp_{tag}_joints"] = v[:, ::53].copy(), p, j
            quat = rg.axisang2quater(aa.reshape(-1, 3)).reshape(Ts, 96).astype(np.float32)
            v, p, j = sp.smooth_pose(quat.copy(), betas, min_cutoff=0.01, beta=0.5, device="cpu", kinectv2=True)
            out.update(sp_aa=aa, sp_betas=betas, sp_quat=quat, sp_quat_verts=v[:, ::53].copy(), sp_quat_pose=p, sp_quat_joints=j)
        finally:
            os.chdir(cwd)
    out["data_checksum"] = np.array(mg.checksum(smpl_data))
    np.savez_compressed(HERE / "postproc.npz", **out)
    print("postproc.npz written:", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
