"""Second, independently written statement of SMPL posing, used ONLY to cross-check oracle/smplx_lbs.py (test infrastructure).

oracle/smplx_lbs.py follows smplx 0.1.26 line by line (relative-joint 4x4 chain, A = G - pad(G.j), W.A as one (V,24)x(24,16)
product).  This file does NOT share that formulation or any code with it: plain Python / numpy FP64 loops over joints and
vertices, the textbook form of linear blend skinning

    world rotation  Rw_j = Rw_parent(j) . R_j             posed joint  Jw_j = Jw_parent(j) + Rw_parent(j) . (J_j - J_parent(j))
    v'              = sum_j w_vj * ( Rw_j . (v_posed - J_j) + Jw_j )

with the rest joints from the shaped template and the pose blend shapes from vec(R_j - I), j = 1..23.  Agreement of the two to
FP32 rounding on random inputs is the evidence that the restatement of smplx (which cannot be imported here)
has no transcription error in the chain / skinning algebra; it is not a substitute for a run of smplx itself.
"""
import numpy as np


def pose_vertices_fp64(data: dict, betas, rotmats, vertex_ids=None):
    """betas (10,), rotmats (24,3,3) -> (posed vertices (n,3) for `vertex_ids` (default all), posed joints (24,3)); float64."""
    f8 = lambda a: np.asarray(a, dtype=np.float64)
    v_template, shapedirs = f8(data["v_template"]), f8(data["shapedirs"])[:, :, :10]
    posedirs, Jreg, W = f8(data["posedirs"]), f8(data["J_regressor"]), f8(data["lbs_weights"])
    parents = [int(p) for p in data["parents"]]
    betas, R = f8(betas), f8(rotmats)
    nj = len(parents)
    V = v_template.shape[0]
    # shaped template and rest joints
    v_shaped = v_template.copy()
    for l in range(10):
        v_shaped += betas[l] * shapedirs[:, :, l]
    J = np.zeros((nj, 3))
    for j in range(nj):
        for k in range(3):
            J[j, k] = float(np.dot(Jreg[j], v_shaped[:, k]))
    # pose feature: row-major vec of (R_j - I) for the 23 body joints
    feat = []
    for j in range(1, nj):
        for a in range(3):
            for b in range(3):
                feat.append(R[j, a, b] - (1.0 if a == b else 0.0))
    feat = np.array(feat)
    # world rotations and posed joint positions by explicit recursion over the tree
    Rw = [None] * nj
    Jw = [None] * nj
    for j in range(nj):
        p = parents[j]
        if p < 0:
            Rw[j] = R[j].copy()
            Jw[j] = J[j].copy()
        else:
            Rw[j] = Rw[p] @ R[j]
            Jw[j] = Jw[p] + Rw[p] @ (J[j] - J[p])
    ids = range(V) if vertex_ids is None else [int(i) for i in vertex_ids]
    out = np.zeros((len(ids), 3))
    for n, v in enumerate(ids):
        vp = v_shaped[v].copy()
        for k in range(3):
            vp[k] += float(np.dot(feat, posedirs[:, 3 * v + k]))
        acc = np.zeros(3)
        for j in range(nj):
            w = W[v, j]
            if w != 0.0:
                acc += w * (Rw[j] @ (vp - J[j]) + Jw[j])
        out[n] = acc
    return out, np.stack(Jw)
