#!/usr/bin/env python
"""SMPL_NEUTRAL.pkl (what the reference unpickles through smplx + chumpy, lib/models/smpl.py:102) -> the .npz this package
loads (gaitb200.smpl.load_smpl_data), and max-grnet.pth.tar style checkpoints -> a plain state dict.

  python scripts/convert_smpl_pkl.py smpl  data/smpl_data/SMPL_NEUTRAL.pkl  data/smpl_data/SMPL_NEUTRAL.npz
  python scripts/convert_smpl_pkl.py ckpt  checkpoint/max-grnet.pth.tar     checkpoint/max-grnet_state.pt

No chumpy / smplx needed: chumpy arrays are unpickled through a stand-in class that keeps only their value.  The conversions
are the ones smplx 0.1.26 body_models.SMPL.__init__ applies: shapedirs[:, :, :10]; posedirs (6890,3,207) -> (207, 20670);
J_regressor densified; parents = kintree_table[0] with the root set to -1; weights -> lbs_weights; f -> faces; landmark vertex
ids of smplx/vertex_ids.py['smplh'] in VertexJointSelector order (face, feet, finger tips).
"""
import pickle
import sys
from pathlib import Path

import numpy as np

LANDMARK_VERTS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                  2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]


class _Ch:
    """Stand-in for chumpy.ch.Ch and friends: keeps the pickled state, exposes the array as .r"""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    @property
    def r(self):
        return np.asarray(self.__dict__.get("x", self.__dict__.get("_x")))


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "chumpy":
            return _Ch
        return super().find_class(module, name)


def _arr(v):
    if isinstance(v, _Ch):
        return v.r
    if hasattr(v, "toarray"):                  # scipy sparse J_regressor
        return np.asarray(v.toarray())
    return np.asarray(v)


def smpl_pkl_to_dict(path) -> dict:
    with open(path, "rb") as f:
        raw = _Unpickler(f, encoding="latin1").load()
    shapedirs = _arr(raw["shapedirs"]).astype(np.float32)[:, :, :10]
    posedirs = _arr(raw["posedirs"]).astype(np.float32)
    V = shapedirs.shape[0]
    posedirs = posedirs.reshape(V * 3, -1).T.copy()                       # (207, 3V), smplx: reshape(-1, P).T
    parents = _arr(raw["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    return {
        "v_template": _arr(raw["v_template"]).astype(np.float32),
        "shapedirs": shapedirs, "posedirs": posedirs,
        "J_regressor": _arr(raw["J_regressor"]).astype(np.float32),
        "lbs_weights": _arr(raw["weights"]).astype(np.float32),
        "faces": _arr(raw["f"]).astype(np.int64),
        "parents": parents,
        "landmark_verts": np.asarray(LANDMARK_VERTS, dtype=np.int64),
    }


def checkpoint_to_state_dict(path):
    """batch_generation.py:210-219: ckpt['gen_state_dict'] when present, else the file itself; keys lose a leading 'module.'"""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = ck.get("gen_state_dict", ck.get("state_dict", ck)) if isinstance(ck, dict) else ck
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def main():
    kind, src, dst = sys.argv[1:4]
    if kind == "smpl":
        d = smpl_pkl_to_dict(src)
        extra = Path(src).parent / "J_regressor_extra.npy"
        if extra.exists():
            d["J_regressor_extra"] = np.load(extra).astype(np.float32)
        np.savez(dst, **d)
        print({k: v.shape for k, v in d.items()})
    elif kind == "ckpt":
        import torch
        sd = checkpoint_to_state_dict(src)
        torch.save(sd, dst)
        print(len(sd), "tensors; regressor keys:", [k for k in sd if k.startswith("regressor.")][:5], "...")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
