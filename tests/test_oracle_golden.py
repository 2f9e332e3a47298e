"""Pin the oracle against vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from gaitb200 import synthetic
from oracle import geometry as OG
from oracle import kp_utils as OK
from oracle import smpl as OS
from oracle import regressor as OR

T = torch.from_numpy


def close(a, b, tol=0.0):
    a = a.numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=0, atol=tol, equal_nan=True)


def test_geometry_functions_match_reference(golden):
    g = golden("geometry")
    # identical op sequences on the same torch build: expect bit-exact, allow 1 ulp-ish slack
    close(OG.rot6d_to_rotmat(T(g["rot6d"])), g["rot6d_to_rotmat"], 1e-7)
    close(OG.rot6d_to_rotmat_spin(T(g["rot6d"][:60])), g["rot6d_to_rotmat_spin"], 1e-7)
    close(OG.rotmat_to_rot6d(T(g["rot6d_to_rotmat"])), g["rotmat_to_rot6d"])
    close(OG.batch_rodrigues(T(g["aa"])), g["batch_rodrigues"], 1e-7)
    close(OG.rotation_matrix_to_quaternion(T(g["Rall"])), g["rotation_matrix_to_quaternion"], 1e-7)
    close(OG.rotation_matrix_to_angle_axis(T(g["Rall"])), g["rotation_matrix_to_angle_axis"], 1e-6)
    close(OG.quaternion_to_angle_axis(T(g["qin"])), g["quaternion_to_angle_axis"], 1e-6)
    close(OG.quat2mat(T(g["qin"])), g["quat2mat"], 1e-7)
    close(OG.projection(T(g["pts"]), T(g["cam"])), g["projection"], 1e-6)
    close(OG.convert_weak_perspective_to_perspective(T(g["cam"])), g["convert_weak_perspective_to_perspective"])
    close(OG.convert_weak_perspective_to_perspective(T(g["cam"]), 1000., 256), g["cwp_1000_256"])
    close(OG.perspective_projection(T(g["pts"]), T(g["pp_rot"]), T(g["pp_trans"]), 1234.5, T(g["pp_center"])),
          g["perspective_projection"], 1e-4)


def test_kp_utils_matches_reference(golden):
    g = golden("kp_utils")
    out = OK.convert_kps(g["joints"], "spin2", "kinectv2")
    assert out.dtype == np.float64 and out.shape == (7, 25, 3)
    close(out, g["spin2_to_kinectv2"])
    assert OK.gather_indices("spin2", "kinectv2") == g["gather"].tolist()
    assert OK.SPIN2_NAMES == g["spin2_names"].tolist()
    assert OK.KINECTV2_NAMES == g["kinectv2_names"].tolist()
    # SURVEY section 4 KAT
    assert g["gather"].tolist() == [0, 6, 12, 15, 16, 18, 20, 22, 17, 19, 21, 23, 1, 4, 7, 10,
                                    2, 5, 8, 11, 28, 25, 24, 27, 26]


def test_joint_tables_match_reference(golden):
    g = golden("smpl_tables")
    assert OS.JOINT_NAMES == g["joint_names"].tolist()
    assert dict(zip(g["joint_map_keys"].tolist(), g["joint_map_vals"].tolist())) == OS.JOINT_MAP
    assert OS.H36M_TO_J17 == g["h36m_to_j17"].tolist()
    assert OS.H36M_TO_J14 == g["h36m_to_j14"].tolist()


def _checksum(arrs):
    return float(sum(np.abs(np.asarray(v, dtype=np.float64)).sum() for v in arrs.values()))


def test_smpl_wrapper_matches_reference(golden, smpl_data):
    g = golden("smpl_wrapper")
    assert abs(_checksum(smpl_data) - float(g["data_checksum"])) < 1e-6, "synthetic SMPL generator drifted"
    smpl = OS.SMPL(smpl_data)
    rot, betas = T(g["rotmat"]), T(g["betas"])
    with torch.no_grad():
        for kin, tag in ((True, "kin"), (False, "spin")):
            smpl.kinectv2 = kin
            so = smpl(betas=betas[:2], body_pose=rot[:2, 1:], global_orient=rot[:2, 0:1], pose2rot=False)
            close(so.vertices, g[f"smpl_{tag}_vertices"], 1e-6)
            close(so.joints, g[f"smpl_{tag}_joints"], 1e-6)
            assert so.joints.shape[1] == (29 if kin else 49)
        smpl.kinectv2 = True
        aa = T(g["smpl_aa"])
        so = smpl(betas=betas[:2], body_pose=aa[:, 3:], global_orient=aa[:, :3], pose2rot=True)
        close(so.vertices, g["smpl_aa_vertices"], 1e-6)
        close(so.joints, g["smpl_aa_joints"], 1e-6)
        head = OS.SMPLHead(smpl_data)
        ho = head(rot[:2], betas[:2], cam=T(g["cam"][:2]), normalize_joints2d=True)
        close(ho["smpl_joints2d"], g["head_joints2d_norm"], 1e-5)
        ho = head(rot[:2], betas[:2], cam=T(g["cam"][:2]), normalize_joints2d=False)
        close(ho["smpl_joints2d"], g["head_joints2d"], 1e-3)   # pixels, ~1e2 magnitude
        close(ho["smpl_joints3d"], g["head_joints3d"], 1e-6)


def test_vpregressor_matches_reference(golden, smpl_data):
    g = golden("vpregressor")
    patt = {"pred_pose": T(g["rotmat"]), "pred_shape": T(g["betas"]), "pred_cam": T(g["cam"])}
    vp = OR.VPRegressor(smpl_data)
    jh = T(smpl_data["J_regressor_h36m"])
    with torch.no_grad():
        o = vp(dict(patt), batch_size=2)[-1]
        assert o["theta"].shape == (2, 2, 85) and o["verts"].shape == (2, 2, 6890, 3)
        assert o["kp_3d"].shape == (2, 2, 29, 3) and o["kp_2d"].shape == (2, 2, 29, 2)
        for k in ("theta", "verts", "kp_3d", "rotmat"):
            close(o[k], g[f"vp_{k}"], 2e-6)
        close(o["kp_2d"], g["vp_kp_2d"], 1e-5)
        o = vp(dict(patt), batch_size=2, J_regressor=jh)[-1]
        assert o["kp_3d"].shape == (2, 2, 14, 3)
        close(o["kp_3d"], g["vp_h36m_kp_3d"], 2e-6)
        sr = OR.SMPLRegressor(smpl_data)
        o = sr({"pred_rotmat": patt["pred_pose"], "pred_shape": patt["pred_shape"], "pred_cam": patt["pred_cam"]}, batch_size=1)
        for k in ("kp_3d", "rotmat", "verts"):
            close(o[k], g[f"sr_{k}"], 2e-6)


def test_regressor_matches_reference(golden, smpl_data):
    g = golden("regressor")
    state = synthetic.make_regressor_state(seed=0, decoder_gain=0.3)
    assert abs(_checksum({k: v.numpy() for k, v in state.items()}) - float(g["state_checksum"])) < 1e-6
    reg = OR.Regressor(smpl_data, synthetic.make_mean_params())
    reg.load_state_dict(state, strict=False)
    reg.eval()
    x = T(g["x"])
    with torch.no_grad():
        o = reg(x)[-1]
        assert o["theta"].shape == (3, 85) and o["kp_3d"].shape == (3, 29, 3)
        for k in ("theta", "verts", "kp_3d", "rotmat"):
            close(o[k], g[f"reg_{k}"], 2e-6)
        close(o["kp_2d"], g["reg_kp_2d"], 1e-5)
        close(reg(x, n_iter=1)[-1]["theta"], g["reg_iter1_theta"], 2e-6)
        o = reg(x, J_regressor=T(smpl_data["J_regressor_h36m"]))[-1]
        close(o["kp_3d"], g["reg_h36m_kp_3d"], 2e-6)
        close(o["kp_2d"], g["reg_h36m_kp_2d"], 1e-5)


def test_postproc_oracle_vs_reference_golden(golden, smpl_data):
    """oracle/postproc.py against the reference's own one_euro_filter.py / demo_utils.py / smooth_pose.py outputs."""
    from oracle import postproc as P
    g = golden("postproc")
    for tag in ("default", "stiff", "fast"):
        mc, b = g[f"oef_{tag}_params"]
        assert np.array_equal(P.one_euro_filter(g["oef_in"], min_cutoff=mc, beta=b), g[f"oef_{tag}"])
    assert np.array_equal(P.convert_crop_cam_to_orig_img(g["cc_cam"], g["cc_bbox"], 1280, 720), g["crop_cam_1280x720"])
    assert np.array_equal(P.convert_crop_cam_to_orig_img(g["cc_cam"], g["cc_bbox"].astype(np.float32), 640, 480), g["crop_cam_f32"])
    assert np.array_equal(P.convert_crop_coords_to_orig_img(g["cc_bbox"], g["cc_kp"].copy(), 224), g["crop_coords_224"])
    for tag, kin in (("spin", False), ("kin", True)):
        v, p, j = P.smooth_pose(smpl_data, g["sp_aa"].copy(), g["sp_betas"], kinectv2=kin)
        assert np.array_equal(p, g[f"sp_{tag}_pose"]) and j.dtype == g[f"sp_{tag}_joints"].dtype
        assert np.abs(v[:, ::53] - g[f"sp_{tag}_verts"]).max() <= 1e-6 and np.abs(j - g[f"sp_{tag}_joints"]).max() <= 1e-6
    v, p, j = P.smooth_pose(smpl_data, g["sp_quat"].copy(), g["sp_betas"], min_cutoff=0.01, beta=0.5, kinectv2=True)
    assert np.array_equal(p, g["sp_quat_pose"]) and np.abs(j - g["sp_quat_joints"]).max() <= 1e-6


def test_heads_oracle_vs_reference_golden(golden):
    """oracle/heads.py against the reference's own LocallyConnected2d / KeypointAttention / PareHead / BidirectionalModel."""
    from oracle import heads as H
    g = golden("heads")
    t = lambda k: torch.from_numpy(g[k])
    assert (H.locally_connected(t("lc_pose_x"), t("lc_pose_w")) - t("lc_pose_y")).abs().max() <= 1e-6
    assert (H.locally_connected(t("lc_cp_x"), t("lc_cp_w"), t("lc_cp_b")) - t("lc_cp_y")).abs().max() <= 1e-6
    assert (H.keypoint_attention(t("ka_feat"), t("ka_heat")) - t("ka_out")).abs().max() <= 1e-6
    assert (H.keypoint_attention(t("ka_feat"), t("ka_heat"), True) - t("ka_out_scaled")).abs().max() <= 1e-6
    sd = {k[len("ph_sd_"):]: t(k) for k in g if k.startswith("ph_sd_")}
    o = H.pare_final(sd, t("ph_smpl_feats"), t("ph_part_attn"))
    for k, gk in (("point_local_feat", "ph_point_local_feat"), ("cam_shape_feats", "ph_cam_shape_feats"), ("pred_rotmat", "ph_pred_rotmat"),
                  ("pred_cam", "ph_pred_cam"), ("pred_shape", "ph_pred_shape"), ("pred_rot6d", "ph_pred_rot6d")):
        assert (o[k] - t(gk)).abs().max() <= 1e-5, k
    o2 = H.pare_final(sd, t("ph_smpl_feats"), t("ph_part_attn"), inits={"pred_rot6d": t("ph_pred_rot6d"), "pred_shape": t("ph_pred_shape"),
                                                                        "pred_cam": t("ph_pred_cam")}, iterative=True)
    assert (o2["pred_rot6d"] - t("ph_it_rot6d")).abs().max() <= 1e-5 and (o2["pred_cam"] - t("ph_it_cam")).abs().max() <= 1e-5
    shapes = {k: tuple(int(x) for x in s.split(",")) for k, s in zip(g["bm_state_keys"], g["bm_state_shapes"])}
    bsd = synthetic.seeded_state(shapes, seed=5)
    y, p, xc = H.bidirectional_model(bsd, t("bm_x"), t("bm_cparams"))
    assert (y - t("bm_y")).abs().max() <= 1e-5 and (p - t("bm_p")).abs().max() <= 1e-5 and (xc[:, :, ::7] - t("bm_xc")).abs().max() <= 1e-6
