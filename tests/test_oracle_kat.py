"""Analytic known-answer tests.  These are the only pins the smplx
arithmetic has (the reference ships no tests; smplx is absent) - CPU only."""
import numpy as np
import torch

from gaitb200 import synthetic
from oracle import geometry as OG
from oracle import smpl as OS
from oracle import smplx_lbs as OL
from oracle import regressor as OR
from oracle.temporal import TemporalEncoder


def test_rot6d_identity_and_orthonormal():
    eye6 = torch.tensor([[1., 0, 0, 1, 0, 0]])
    assert torch.equal(OG.rot6d_to_rotmat(eye6)[0], torch.eye(3))
    R = OG.rot6d_to_rotmat(torch.randn(100, 6))
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3).expand(100, 3, 3), atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(100), atol=1e-5)


def test_rest_pose_returns_template(smpl_data):
    smpl = OL.SMPLX_SMPL(smpl_data)
    R = torch.eye(3).expand(2, 24, 3, 3)
    so = smpl(betas=torch.zeros(2, 10), body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    vt = torch.from_numpy(smpl_data["v_template"])
    assert torch.allclose(so.vertices, vt.expand(2, -1, -1), atol=1e-6)
    J = torch.from_numpy(smpl_data["J_regressor"]) @ vt
    assert torch.allclose(so.joints[:, :24], J.expand(2, -1, -1), atol=1e-6)
    lm = torch.from_numpy(smpl_data["landmark_verts"])
    assert torch.allclose(so.joints[:, 24:], vt[lm].expand(2, -1, -1), atol=1e-6)


def test_root_only_rotation_is_rigid(smpl_data):
    smpl = OL.SMPLX_SMPL(smpl_data)
    R = torch.eye(3).repeat(1, 24, 1, 1)
    R0 = OG.rot6d_to_rotmat(torch.randn(1, 6))[0]
    R[0, 0] = R0
    betas = torch.randn(1, 10)
    so = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    v_shaped = torch.from_numpy(smpl_data["v_template"]) + torch.einsum(
        'l,mkl->mk', betas[0], torch.from_numpy(smpl_data["shapedirs"]))
    J0 = (torch.from_numpy(smpl_data["J_regressor"]) @ v_shaped)[0]
    # rows of lbs_weights sum to 1 and pose_feature(body)=0  =>  verts = R0 (v - J0) + J0
    expect = (v_shaped - J0) @ R0.T + J0
    assert torch.allclose(so.vertices[0], expect, atol=2e-6)


def test_axis_angle_round_trip():
    a = torch.randn(200, 3)
    a = a / a.norm(dim=1, keepdim=True) * (torch.rand(200, 1) * 3.0 + 0.01)
    R = OL.batch_rodrigues(a)
    back = OG.rotation_matrix_to_angle_axis(R)
    assert torch.allclose(back, a, atol=2e-4)
    R2 = OG.batch_rodrigues(a).view(-1, 3, 3)       # the reference's quaternion variant agrees with smplx's
    assert torch.allclose(R, R2, atol=1e-5)


def test_projection_formula():
    X = torch.randn(3, 5, 3) * 0.3
    cam = torch.tensor([[0.9, 0.1, -0.2], [1.1, 0.0, 0.0], [0.7, -0.3, 0.2]])
    t = torch.stack([cam[:, 1], cam[:, 2], 2 * 5000. / (224. * cam[:, 0] + 1e-9)], -1)
    P = X + t[:, None]
    expect = 5000. * P[..., :2] / P[..., 2:] / 112.
    assert torch.allclose(OG.projection(X, cam), expect, atol=1e-5)


def test_regressor_zero_decoders_returns_init(smpl_data):
    reg = OR.Regressor(smpl_data, synthetic.make_mean_params()).eval()
    for m in (reg.decpose, reg.decshape, reg.deccam):
        torch.nn.init.zeros_(m.weight); torch.nn.init.zeros_(m.bias)
    x = torch.randn(4, 2048)
    for n in (1, 3, 5):
        p, s, c = reg.iterate(x, n_iter=n)
        assert torch.equal(p, reg.init_pose.expand(4, -1)) and torch.equal(c, reg.init_cam.expand(4, -1))


def test_gru_cell_equations():
    """nn.GRU gate order [r,z,n], h' = (1-z) n + z h (what the CUDA cell must reproduce)."""
    enc = TemporalEncoder(hidden_size=32, input_size=32).eval()
    x = torch.randn(2, 5, 32)
    with torch.no_grad():
        y = enc(x)
        g = enc.gru
        h = torch.zeros(2, 32)
        outs = []
        for t in range(5):
            gi = x[:, t] @ g.weight_ih_l0.T + g.bias_ih_l0
            gh = h @ g.weight_hh_l0.T + g.bias_hh_l0
            r = torch.sigmoid(gi[:, :32] + gh[:, :32])
            z = torch.sigmoid(gi[:, 32:64] + gh[:, 32:64])
            n = torch.tanh(gi[:, 64:] + r * gh[:, 64:])
            h = (1 - z) * n + z * h
            outs.append(h + x[:, t])
    assert torch.allclose(y, torch.stack(outs, 1), atol=1e-5)


def test_kinect_joint_selection(smpl_data):
    """spin2 joints 24..28 = smplx joints 35,37,40,42 (landmark verts) + J_regressor_extra row 5."""
    smpl = OS.SMPL(smpl_data)
    rot6d, betas, _ = synthetic.make_pose_inputs(2, seed=1)
    R = OG.rot6d_to_rotmat(rot6d).view(2, 24, 3, 3)
    so = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    lm = smpl_data["landmark_verts"]
    picks = [lm[35 - 24], lm[37 - 24], lm[40 - 24], lm[42 - 24]]
    assert picks == [2746, 2445, 6191, 5905]
    assert torch.equal(so.joints[:, 24:28], so.vertices[:, picks])
    thorax = torch.from_numpy(smpl_data["J_regressor_extra"][5]) @ so.vertices
    assert torch.allclose(so.joints[:, 28], thorax, atol=1e-6)


# ---- hand-computed three-joint chain (root -> joint 1 -> joint 4 of the SMPL tree) with two skinned vertices.
# Rotations are about z: R0 = Rz(90), R1 = Rz(90), R4 = Rz(90); rest joints J0 = (0,0,0), J1 = (0,1,0), J4 = (0,2,0).
#   posed joints: J0' = 0;  J1' = R0 J1 = (-1,0,0);  J4' = J1' + R0 R1 (J4 - J1) = (-1,0,0) + Rz(180)(0,1,0) = (-1,-1,0)
#   vertex a = (0.5,1.5,0), weights 1/2 on joint 1, 1/2 on joint 4:
#       by joint 1: Rz(180)(a - J1) + J1' = (-0.5,-0.5,0) + (-1,0,0)  = (-1.5,-0.5,0)
#       by joint 4: Rz(270)(a - J4) + J4' = (-0.5,-0.5,0) + (-1,-1,0) = (-1.5,-1.5,0)        -> a' = (-1.5,-1.0,0)
#   vertex b = (0.25,0.5,0.1), weights 3/4 on the root, 1/4 on joint 1:
#       by root:    Rz(90) b              = (-0.5,0.25,0.1)
#       by joint 1: Rz(180)(b - J1) + J1' = (-0.25,0.5,0.1) + (-1,0,0) = (-1.25,0.5,0.1)     -> b' = (-0.6875,0.3125,0.1)
def chain_kat():
    import math
    parents = torch.tensor(synthetic.SMPL_PARENTS)
    Rz = lambda d: torch.tensor([[math.cos(math.radians(d)), -math.sin(math.radians(d)), 0.],
                                 [math.sin(math.radians(d)), math.cos(math.radians(d)), 0.], [0., 0., 1.]], dtype=torch.float64).float()
    R = torch.eye(3).repeat(1, 24, 1, 1)
    R[0, 0], R[0, 1], R[0, 4] = Rz(90), Rz(90), Rz(90)
    J = torch.zeros(1, 24, 3)
    J[0, 1] = torch.tensor([0., 1., 0.])
    J[0, 4] = torch.tensor([0., 2., 0.])
    for j in range(24):                      # park the other joints away from the chain; their weights are zero
        if j not in (0, 1, 4):
            J[0, j] = torch.tensor([3. + j, -2., 1.])
    verts = torch.tensor([[[0.5, 1.5, 0.], [0.25, 0.5, 0.1]]])
    W = torch.zeros(2, 24)
    W[0, 1], W[0, 4] = 0.5, 0.5
    W[1, 0], W[1, 1] = 0.75, 0.25
    expect_v = torch.tensor([[[-1.5, -1.0, 0.], [-0.6875, 0.3125, 0.1]]])
    expect_j = {0: [0., 0., 0.], 1: [-1., 0., 0.], 4: [-1., -1., 0.]}
    return parents, R, J, verts, W, expect_v, expect_j


def test_three_joint_chain_hand_computed():
    parents, R, J, verts, W, expect_v, expect_j = chain_kat()
    posed, A = OL.batch_rigid_transform(R, J, parents)
    for j, e in expect_j.items():
        assert torch.allclose(posed[0, j], torch.tensor(e), atol=1e-6), (j, posed[0, j])
    Tm = torch.matmul(W.unsqueeze(0), A.view(1, 24, 16)).view(1, 2, 4, 4)
    out = torch.matmul(Tm, torch.cat([verts, torch.ones(1, 2, 1)], 2).unsqueeze(-1))[:, :, :3, 0]
    assert torch.allclose(out, expect_v, atol=1e-6), out


def test_oracle_matches_independent_fp64_loops(smpl_data):
    """oracle/smplx_lbs.py (smplx's tensor formulation) against oracle/independent_lbs.py (textbook per-vertex FP64 loops that
    share no code with it): 2 random poses, 300 random vertices + all landmark vertices, every posed joint."""
    from oracle.independent_lbs import pose_vertices_fp64
    smpl = OL.SMPLX_SMPL(smpl_data)
    rot6d, betas, _ = synthetic.make_pose_inputs(2, seed=11, noise=0.6)
    R = OG.rot6d_to_rotmat(rot6d).view(2, 24, 3, 3)
    so = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    rng = np.random.default_rng(3)
    ids = np.unique(np.concatenate([rng.choice(6890, 300, replace=False), np.asarray(smpl_data["landmark_verts"])]))
    for f in range(2):
        v64, j64 = pose_vertices_fp64(smpl_data, betas[f].numpy(), R[f].numpy(), ids)
        assert np.abs(so.vertices[f, ids].numpy() - v64).max() <= 5e-6
        assert np.abs(so.joints[f, :24].numpy() - j64).max() <= 5e-6
