"""Seeded synthetic stand-ins for the data files the reference downloads.

The reference needs ``data/smpl_data/{SMPL_NEUTRAL.pkl, J_regressor_extra.npy,
smpl_mean_params.npz}`` (lib/models/smpl.py:90-92, lib/core/config.py:23) and
trained checkpoints; none are available offline.  These factories produce arrays
of identical shape/dtype, conditioned like the real data (row-stochastic
regressors and skin weights, small blend-shape magnitudes, near-mean-pose
regressor), following SURVEY.md section 8(d).  Host-side numpy/torch only; no
arithmetic of the hot path happens here.
"""
from __future__ import annotations

import numpy as np
import torch

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_BASIS = 207
NUM_FACES = 13776

# SMPL kinematic tree (kintree_table[0] with the root set to -1).
SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.int64,
)

# Landmark vertices smplx's VertexJointSelector appends for the SMPL body
# (face 5, feet 6, finger tips 2x5), in that order -> joints 24..44.
SMPL_LANDMARK_VERTS = np.array(
    [332, 6260, 2800, 4071, 583,
     3216, 3226, 3387, 6617, 6624, 6787,
     2746, 2319, 2445, 2556, 2673,
     6191, 5782, 5905, 6016, 6133],
    dtype=np.int64,
)


def _row_stochastic(rng, rows, cols, nnz=None):
    """Non-negative rows that sum to one; `nnz` entries per row, or dense."""
    if nnz is None:
        m = rng.random((rows, cols)) + 1e-3
    else:
        m = np.zeros((rows, cols))
        for r in range(rows):
            idx = rng.choice(cols, size=nnz, replace=False)
            m[r, idx] = rng.random(nnz) + 0.05
    m /= m.sum(axis=1, keepdims=True)
    return m.astype(np.float32)


def make_smpl_data(seed: int = 0, variant: str = "sparse") -> dict:
    """SMPL-shaped model arrays.

    variant "sparse": <=4 non-zero skin weights per vertex and ~30 non-zeros per
    joint-regressor row, like the real SMPL_NEUTRAL model; "dense": every entry
    non-zero (the worst case smplx's dense arithmetic is written for).
    """
    if variant not in ("sparse", "dense"):
        raise ValueError(f"unknown variant {variant!r}")
    rng = np.random.default_rng(seed)
    v_template = (rng.standard_normal((NUM_VERTS, 3)) * 0.3 * np.array([0.25, 0.9, 0.15])).astype(np.float32)
    shapedirs = (rng.standard_normal((NUM_VERTS, 3, NUM_BETAS)) * 0.01).astype(np.float32)
    # smplx keeps posedirs as (207, 6890*3)
    posedirs = (rng.standard_normal((NUM_POSE_BASIS, NUM_VERTS * 3)) * 1e-3).astype(np.float32)
    if variant == "sparse":
        j_reg = _row_stochastic(rng, NUM_JOINTS, NUM_VERTS, nnz=30)
        j_extra = _row_stochastic(rng, 9, NUM_VERTS, nnz=30)
        j_h36m = _row_stochastic(rng, 17, NUM_VERTS, nnz=60)
        n_inf = rng.integers(1, 5, size=NUM_VERTS)
        w = np.zeros((NUM_VERTS, NUM_JOINTS))
        for v in range(NUM_VERTS):
            idx = rng.choice(NUM_JOINTS, size=n_inf[v], replace=False)
            w[v, idx] = rng.random(n_inf[v]) + 0.05
        w /= w.sum(axis=1, keepdims=True)
        lbs_weights = w.astype(np.float32)
    else:
        j_reg = _row_stochastic(rng, NUM_JOINTS, NUM_VERTS)
        j_extra = _row_stochastic(rng, 9, NUM_VERTS)
        j_h36m = _row_stochastic(rng, 17, NUM_VERTS)
        lbs_weights = _row_stochastic(rng, NUM_VERTS, NUM_JOINTS)
    faces = rng.integers(0, NUM_VERTS, size=(NUM_FACES, 3)).astype(np.int64)
    return {
        "v_template": v_template,
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": j_reg,
        "parents": SMPL_PARENTS.copy(),
        "lbs_weights": lbs_weights,
        "faces": faces,
        "J_regressor_extra": j_extra,
        "J_regressor_h36m": j_h36m,
        "landmark_verts": SMPL_LANDMARK_VERTS.copy(),
    }


def make_mean_params() -> dict:
    """Stand-in for smpl_mean_params.npz: identity pose in 6-D, zero shape, cam [0.9,0,0]."""
    pose = np.tile(np.array([1, 0, 0, 1, 0, 0], dtype=np.float32), NUM_JOINTS)
    return {
        "pose": pose,
        "shape": np.zeros(NUM_BETAS, dtype=np.float32),
        "cam": np.array([0.9, 0.0, 0.0], dtype=np.float32),
    }


def make_regressor_state(seed: int = 0, decoder_gain: float = 0.01) -> dict:
    """state_dict of the HMR regressor MLP (lib/models/spin.py:216-240): nn.Linear
    default init, decoders xavier_uniform(gain=0.01)."""
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / (in_f ** 0.5)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return w, b

    def xavier(out_f, in_f):
        a = decoder_gain * (6.0 / (in_f + out_f)) ** 0.5
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a

    sd = {}
    sd["fc1.weight"], sd["fc1.bias"] = linear(1024, 2048 + 144 + 13)
    sd["fc2.weight"], sd["fc2.bias"] = linear(1024, 1024)
    for name, n in (("decpose", 144), ("decshape", 10), ("deccam", 3)):
        _, b = linear(n, 1024)
        sd[f"{name}.weight"], sd[f"{name}.bias"] = xavier(n, 1024), b
    mp = make_mean_params()
    sd["init_pose"] = torch.from_numpy(mp["pose"]).unsqueeze(0)
    sd["init_shape"] = torch.from_numpy(mp["shape"]).unsqueeze(0)
    sd["init_cam"] = torch.from_numpy(mp["cam"]).unsqueeze(0)
    return sd


def make_gru_state(seed: int = 0, input_size: int = 2048, hidden_size: int = 2048,
                   num_layers: int = 1, bidirectional: bool = False) -> dict:
    """torch.nn.GRU-keyed state (weight_ih_l0, weight_hh_l0, bias_*; `_reverse`
    suffix for the backward direction), U(-1/sqrt(H), 1/sqrt(H)) like torch."""
    g = torch.Generator().manual_seed(seed + 7919)
    k = 1.0 / (hidden_size ** 0.5)
    sd = {}
    dirs = 2 if bidirectional else 1
    for layer in range(num_layers):
        in_f = input_size if layer == 0 else hidden_size * dirs
        for d in range(dirs):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            sd["weight_ih" + sfx] = (torch.rand(3 * hidden_size, in_f, generator=g) * 2 - 1) * k
            sd["weight_hh" + sfx] = (torch.rand(3 * hidden_size, hidden_size, generator=g) * 2 - 1) * k
            sd["bias_ih" + sfx] = (torch.rand(3 * hidden_size, generator=g) * 2 - 1) * k
            sd["bias_hh" + sfx] = (torch.rand(3 * hidden_size, generator=g) * 2 - 1) * k
    return sd


def make_features(num_seqs: int, seq_len: int, seed: int = 1234, dim: int = 2048) -> torch.Tensor:
    """Backbone-feature stand-in: |N(0,1)|*0.5, shape (S, T, dim) fp32."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(num_seqs, seq_len, dim, generator=g).abs_().mul_(0.5)


def make_pose_inputs(num_frames: int, seed: int = 0, noise: float = 0.3):
    """(rot6d (F,144), betas (F,10), cam (F,3)) for SMPL-stage tests/microbenches."""
    g = torch.Generator().manual_seed(seed + 104729)
    ident = torch.tensor([1.0, 0, 0, 1, 0, 0]).repeat(NUM_JOINTS)
    rot6d = ident + noise * torch.randn(num_frames, 144, generator=g)
    betas = torch.randn(num_frames, NUM_BETAS, generator=g)
    cam = torch.stack([
        0.6 + 0.6 * torch.rand(num_frames, generator=g),
        0.2 * torch.randn(num_frames, generator=g),
        0.2 * torch.randn(num_frames, generator=g)], dim=1)
    return rot6d, betas, cam


def seeded_state(shapes: dict, seed: int = 0, scale: float = 0.05) -> dict:
    """Deterministic stand-in for a checkpoint: one U(-scale, scale) tensor per (key, shape), generated key by key in
    sorted order from its own seeded generator - so a golden-vector script and a test can build the same weights from
    nothing but the state_dict keys and shapes (used for BidirectionalModel, whose 30 MB of GRU weights are not committed)."""
    out = {}
    for i, k in enumerate(sorted(shapes)):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        out[k] = (torch.rand(tuple(shapes[k]), generator=g) * 2 - 1) * scale
    return out
