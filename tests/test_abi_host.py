"""CPU-side checks: the C-ABI library loads and exports every symbol include/gaitb200.h declares,
argument validation works without a GPU (no compute is enqueued), the host mirror of the
reference API has the reference's tables / state_dict keys, and the product refuses CPU input."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from gaitb200 import _lib, kp_utils, synthetic
from gaitb200 import smpl as PS
from oracle import kp_utils as OK
from oracle import regressor as OR
from oracle import smpl as OS

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        subprocess.run(["make", "-C", str(ROOT), "-j", "8"], check=True)
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "gaitb200.h").read_text()
    declared = sorted(set(re.findall(r"\b(gait_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gaitb200.h but not exported"
    assert set(declared) == set(_lib.EXPORTS), set(declared) ^ set(_lib.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_kernels_use_the_blackwell_instructions_design_md_names(lib):
    """SASS of the built library: the hot kernels really are tcgen05 / tensor-memory / TMA code with per-role register
    budgets and programmatic dependent launch (mnemonics as in /opt/skills/guides/B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    kernels = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = set()
        elif cur is not None:
            for mn in ("UTCHMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UBLKCP", "USETMAXREG", "ACQBULK"):
                if mn in line:
                    kernels[cur].add(mn)

    def has(name_part, *mns):
        hits = [k for k in kernels if name_part in k]
        assert hits, name_part
        for k in hits:
            assert set(mns) <= kernels[k], (k, sorted(kernels[k]))

    has("gemm_tf32x3_kernel", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "USETMAXREG", "ACQBULK")      # tcgen05.mma, TMEM ld/st, TMA, setmaxnreg, griddepcontrol.wait
    has("gru_recurrent_kernel", "UTCHMMA", "UTCCP", "LDTM", "UTMALDG", "USETMAXREG")               # + tcgen05.cp
    has("smpl_lbs_tc_kernel", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "USETMAXREG", "ACQBULK")
    has("joint_regress_stream_kernel", "UBLKCP")


def test_version_and_error_strings(lib):
    assert lib.gait_abi_version() == 1
    assert lib.gait_error_string(0) == b"ok"
    assert b"invalid" in lib.gait_error_string(-1)
    assert lib.gait_launch_count() >= 0


def test_argument_validation_without_gpu(lib):
    # empty problems succeed without touching the device
    assert lib.gait_rot6d_to_rotmat(None, 1, 6, None, 0, 1e-6, None) == 0
    assert lib.gait_smpl_lbs(None, 0, None, None, None, 0, 6890, None) == 0
    assert lib.gait_smpl_lbs_tc(None, 0, None, None, None, None, None, 0, 6890, None) == 0
    assert lib.gait_linear(None, 0, None, 0, None, None, 0, None, 0, 0, 16, 16, None) == 0
    assert lib.gait_gru_layer(None, 0, None, None, None, None, None, None, 0, None, 0, None, 0, None,
                              0, 16, 8, 8, 0, None, 0, None) == 0
    # bad arguments are rejected before any launch
    assert lib.gait_rot6d_to_rotmat(None, 1, 6, None, 4, 1e-6, None) == -1
    assert b"null" in lib.gait_last_error()
    assert lib.gait_rot6d_to_rotmat(None, 1, 6, None, -1, 1e-6, None) == -1
    assert lib.gait_rotmat_to_quaternion(C.c_void_p(16), 5, C.c_void_p(16), 1, 1e-6, None) == -1
    assert lib.gait_batch_rodrigues(C.c_void_p(16), C.c_void_p(16), 1, 7, None) == -1
    assert lib.gait_smpl_lbs(C.c_void_p(16), 3 * 6892, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, 6891, None) == -1
    assert lib.gait_smpl_lbs(C.c_void_p(20), 20670, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, 6890, None) == -1
    # tensor-core LBS: v_posed rows must be padded to whole 128-vertex tiles
    assert lib.gait_smpl_lbs_tc(C.c_void_p(16), 20670, C.c_void_p(16), C.c_void_p(16), None, C.c_void_p(16), None, 1, 6890, None) == -1
    assert lib.gait_smpl_lbs_pack_bytes(6890) == 54 * 24576 and lib.gait_smpl_lbs_aop_bytes(1024) == 128 * 18432
    assert lib.gait_linear(C.c_void_p(16), 4, C.c_void_p(16), 8, None, None, 0, C.c_void_p(16), 8, 2, 8, 8, None) == -1
    assert lib.gait_gru_workspace_bytes(64, 16, 2048) == (64 * 16 * 6144 + 4 * 64 * 6144) * 4 + 128 * (2048 // 16 + 2) + 2 * 64 * 2048 * 4   # + step flags, h_lo scratch
    assert lib.gait_hmr_workspace_bytes(1024, 1024) == (3 * 1024 * 1024 + 6 * 1024 * 160 + 2 * 1024) * 4
    assert lib.gait_hmr_folded_workspace_bytes(1024) == 6 * 1024 * 160 * 4 and lib.gait_hmr_folded_workspace_bytes(0) == 0
    assert lib.gait_hmr_regressor_folded(None, 2048, None, None, None, 4, 2048, None, 0, None) == -1   # null pointers
    n0 = lib.gait_launch_count()
    assert lib.gait_gru_layer(C.c_void_p(16), 8, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), None,
                              C.c_void_p(16), 8, None, 0, None, 0, None, 2, 3, 8, 8, 0, C.c_void_p(16), 8, None) == -4
    assert lib.gait_launch_count() == n0


def test_new_entry_points_validate_without_gpu(lib):
    """joints-only skinning, post-processing, heads and prepared weights: empty problems succeed, bad arguments fail, no launch."""
    n0 = lib.gait_launch_count()
    assert lib.gait_smpl_lbs_tc_joints(None, 0, None, None, None, None, None, 21, None, 0, 6890, None) == 0
    assert lib.gait_smpl_lbs_tc_joints(C.c_void_p(16), 20736, C.c_void_p(16), C.c_void_p(16), None, None, None, 0, None, 4, 6890, None) == -1
    assert lib.gait_one_euro_filter(None, None, 0, 72, 0.004, 0.7, 1.0, None) == 0
    assert lib.gait_one_euro_filter(None, None, 4, 72, 0.004, 0.7, 1.0, None) == -1
    assert lib.gait_crop_cam_to_orig_img(None, None, 1, 4, 1280.0, 720.0, None, 0, None) == 0
    assert lib.gait_crop_cam_to_orig_img(C.c_void_p(16), C.c_void_p(16), 1, 2, 1280.0, 720.0, C.c_void_p(16), 3, None) == -1
    assert lib.gait_crop_coords_to_orig_img(C.c_void_p(16), 1, 4, C.c_void_p(16), C.c_void_p(16), 3, 25, 1, 224.0, None) == -1
    assert lib.gait_keypoint_attention(None, None, 1.0, None, 0, 128, 24, 3136, 1, 1, 1, None) == 0
    assert lib.gait_keypoint_attention(C.c_void_p(16), C.c_void_p(16), 1.0, C.c_void_p(16), 2, 128, 24, 1 << 20, 1, 1, 1, None) == -1
    assert lib.gait_locally_connected(C.c_void_p(16), 1, 1, 1, C.c_void_p(16), 1, 1, 1, None, 0, 0, C.c_void_p(16), 1, 1, 1,
                                      C.c_void_p(16), None, 2, 3, 4, 24, None) == -1          # resid without out2
    assert lib.gait_activation(C.c_void_p(16), C.c_void_p(16), 8, 7, 0.0, None) == -1
    assert lib.gait_prepare_weight(None, None, 0, None) == 0 and lib.gait_prepare_weight(None, None, 8, None) == -1
    assert lib.gait_release_weight(C.c_void_p(16)) == 0
    assert lib.gait_launch_count() == n0


def test_postproc_and_heads_have_no_cpu_path():
    """Without a CUDA device the new host modules raise instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    from gaitb200 import postproc as PP
    from gaitb200.layers import KeypointAttention, LocallyConnected2d
    with pytest.raises(_lib.GaitLibraryError):
        PP.one_euro_filter(np.zeros((4, 72), np.float32))
    with pytest.raises(_lib.GaitLibraryError):
        PP.convert_crop_cam_to_orig_img(np.zeros((2, 3), np.float32), np.ones((2, 4)), 640, 480)
    with pytest.raises(TypeError):
        PP.one_euro_filter(np.zeros((4, 72), np.float64))
    with pytest.raises(_lib.GaitLibraryError):
        KeypointAttention()(torch.zeros(1, 8, 4, 4), torch.zeros(1, 24, 4, 4))
    with pytest.raises(_lib.GaitLibraryError):
        LocallyConnected2d(3, 4, [24, 1], 1, 1)(torch.zeros(1, 3, 24, 1))


def test_product_has_no_cpu_path(smpl_data):
    from gaitb200 import geometry as G
    with pytest.raises(_lib.GaitLibraryError):
        G.rot6d_to_rotmat(torch.zeros(2, 6))
    with pytest.raises(TypeError):
        G.rot6d_to_rotmat(np.zeros((2, 6), np.float32))
    with pytest.raises(TypeError):
        G.rotation_matrix_to_quaternion([1, 2, 3])
    with pytest.raises(ValueError):
        G.quaternion_to_angle_axis(torch.zeros(2, 3))
    smpl = PS.SMPL(smpl_data)
    with pytest.raises(_lib.GaitLibraryError):
        smpl(betas=torch.zeros(1, 10), body_pose=torch.eye(3).expand(1, 23, 3, 3),
             global_orient=torch.eye(3).expand(1, 1, 3, 3), pose2rot=False)


def test_product_sources_do_not_import_oracle():
    pkg = ROOT / "video-based-gait-analysis-for-dementia_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f


def test_tables_match_oracle_and_golden(golden):
    g = golden("kp_utils")
    assert kp_utils.get_spin2_joint_names() == g["spin2_names"].tolist() == OK.SPIN2_NAMES
    assert kp_utils.get_kinectv2_joint_names() == g["kinectv2_names"].tolist()
    assert kp_utils.SPIN2_TO_KINECTV2 == g["gather"].tolist()
    assert kp_utils.get_kinectv2_skeleton().tolist() == g["kinectv2_skeleton"].tolist()
    t = golden("smpl_tables")
    assert PS.JOINT_NAMES == t["joint_names"].tolist()
    assert PS.JOINT_MAP == dict(zip(t["joint_map_keys"].tolist(), t["joint_map_vals"].tolist()))
    assert PS.H36M_TO_J14 == t["h36m_to_j14"].tolist() and PS.H36M_TO_J17 == t["h36m_to_j17"].tolist()
    with pytest.raises(NameError):
        kp_utils.gather_indices("spin2", "nope")


def test_state_dict_keys_match_reference_layout(smpl_data):
    from gaitb200.regressor import Regressor, VPRegressor
    from gaitb200.temporal import TemporalEncoder
    mean = synthetic.make_mean_params()
    mine = Regressor(mean, smpl_data)
    keys = set(mine.state_dict().keys())
    for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "decpose.weight", "decpose.bias",
              "decshape.weight", "decshape.bias", "deccam.weight", "deccam.bias", "init_pose", "init_shape",
              "init_cam", "smpl.v_template", "smpl.shapedirs", "smpl.posedirs", "smpl.J_regressor",
              "smpl.lbs_weights", "smpl.parents", "smpl.faces_tensor", "smpl.J_regressor_extra",
              "smpl.betas", "smpl.global_orient", "smpl.body_pose", "smpl.vertex_joint_selector.extra_joints_idxs"):
        assert k in keys, k
    assert mine.init_pose.shape == (1, 144) and mine.init_shape.shape == (1, 10) and mine.init_cam.shape == (1, 3)
    # loads the synthetic checkpoint strictly for the MLP part
    state = synthetic.make_regressor_state(seed=0)
    res = mine.load_state_dict(state, strict=False)
    assert not res.unexpected_keys
    # oracle regressor (restating spin.py) exposes the same MLP / init keys
    ork = set(OR.Regressor(smpl_data, mean).state_dict().keys())
    assert {k for k in ork if not k.startswith("smpl.")} <= keys
    vp = VPRegressor(smpl_model_dir=smpl_data)
    assert "smpl.smpl.J_regressor_extra" in vp.state_dict() and "smpl.smpl.lbs_weights" in vp.state_dict()
    enc = TemporalEncoder(n_layers=2, hidden_size=32, bidirectional=True, input_size=48)
    assert {"gru.weight_ih_l0", "gru.weight_hh_l1_reverse", "linear.weight"} <= set(enc.state_dict().keys())
    assert tuple(mine.smpl.joint_map.tolist()) == tuple(OS.SMPL(smpl_data).joint_map.tolist())
    assert PS.SMPL.extra is True and PS.SMPL.kinectv2 is True


def test_kinect_db_writer_format(tmp_path):
    """the joblib database of batch_generation.py:226-283 / doc/batch_generation.md - keys, dtypes, shapes,
    per-frame video names, shard naming and the reference's rule for cutting shards."""
    import joblib
    import numpy as np
    from gaitb200.postproc import KinectDbWriter
    rng = np.random.default_rng(0)
    vids = [(f"clip{i}.mp4", rng.standard_normal((3 + i, 25, 3)), rng.standard_normal((3 + i, 4))) for i in range(5)]
    w = KinectDbWriter(str(tmp_path / "db.json"), max_videos=2, min_tail=1)
    for i, (name, j, b) in enumerate(vids):
        w.add(name, torch.from_numpy(j) if i % 2 else j.reshape(-1, 75), b, videos_left=len(vids) - 1 - i)
    files = w.close()
    assert [f.rsplit("/", 1)[1] for f in files] == ["db_0.json", "db_1.json"]     # cut before clip2; before clip4 only 1 video remains
    dbs = [joblib.load(f) for f in files]
    assert all(set(d) == {"vid_name", "bbox", "joints3D"} for d in dbs)
    assert dbs[0]["joints3D"].shape == (3 + 4, 25, 3) and dbs[1]["joints3D"].shape == (5 + 6 + 7, 25, 3)
    assert dbs[0]["joints3D"].dtype == np.float32 and dbs[0]["bbox"].dtype == np.float32 and dbs[0]["bbox"].shape == (7, 4)
    assert list(dbs[0]["vid_name"]) == ["clip0"] * 3 + ["clip1"] * 4
    allj = np.concatenate([v[1] for v in vids]).astype(np.float32)
    assert np.array_equal(np.concatenate([d["joints3D"] for d in dbs]), allj)
    with pytest.raises(ValueError):
        KinectDbWriter(str(tmp_path / "db.pkl"))
    # skipped videos (no annotations, batch_generation.py:246-248) advance the index the shard rule uses: with
    # MAX_VID = 2 and the list [a, SKIP, b, c, SKIP, d] the reference cuts before idx 2 and idx 4 (6 - 4 = 2 > 1)
    w = KinectDbWriter(str(tmp_path / "sk.json"), max_videos=2, min_tail=1, total=6)
    seq = ["a", None, "b", "c", None, "d"]
    for name in seq:
        if name is None:
            w.skip()
        else:
            w.add(name + ".avi", rng.standard_normal((2, 25, 3)), rng.standard_normal((2, 4)))
    dbs = [joblib.load(f) for f in w.close()]
    assert [sorted(set(d["vid_name"])) for d in dbs] == [["a"], ["b", "c"], ["d"]]


def test_strict_load_of_reference_key_set(smpl_data):
    """batch_generation.py:210-219 loads the checkpoint with strict=True.  tests/golden/state_dict_keys.json holds key -> shape of
    the state dicts of the reference's own spin.Regressor / pare.VPRegressor / pare.SMPLRegressor / smpl.SMPLHead (unmodified
    source behind import stubs, tests/golden/make_golden_keys.py).  A state dict with exactly that key set must load strictly
    into the drop-in modules.  One key is renamed: the stub that stands in for smplx registers `extra_joints_idxs` on the body
    model itself, real smplx 0.1.26 keeps it in the `vertex_joint_selector` sub-module - the product follows smplx."""
    import json
    from pathlib import Path
    from gaitb200.regressor import Regressor, SMPLRegressor, VPRegressor
    from gaitb200.smpl import SMPLHead
    ref = json.loads((Path(__file__).parent / "golden" / "state_dict_keys.json").read_text())
    smpl_data = {k: np.array(v, copy=True) for k, v in smpl_data.items()}     # module buffers may alias the arrays they are built from
    mean = synthetic.make_mean_params()
    mods = {"spin.Regressor": Regressor(mean, smpl_data), "pare.VPRegressor": VPRegressor(smpl_model_dir=smpl_data),
            "pare.SMPLRegressor": SMPLRegressor(smpl_model_dir=smpl_data), "smpl.SMPLHead": SMPLHead(smpl_model_dir=smpl_data)}
    g = torch.Generator().manual_seed(0)
    for name, mod in mods.items():
        want = {k.replace("extra_joints_idxs", "vertex_joint_selector.extra_joints_idxs"): tuple(s) for k, s in ref[name].items()}
        have = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        assert have == want, (name, set(have) ^ set(want), {k: (have[k], want[k]) for k in have if k in want and have[k] != want[k]})
        # a "checkpoint" with the reference's keys, shapes and dtypes loads strictly and lands in the module
        ckpt = {k: (torch.randn(*s, generator=g) if v.dtype.is_floating_point else torch.zeros(*s, dtype=v.dtype))
                for (k, s), v in zip(want.items(), [mod.state_dict()[k] for k in want])}
        res = mod.load_state_dict(ckpt, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        k0 = next(k for k in want if k.endswith("lbs_weights"))
        assert torch.equal(mod.state_dict()[k0], ckpt[k0])


def test_smpl_pkl_loader_without_chumpy(tmp_path, smpl_data):
    """The reference loads data/smpl_data/SMPL_NEUTRAL.pkl through smplx + chumpy (lib/models/smpl.py:102).  A pickle with
    the official file's structure - chumpy arrays, a scipy-sparse J_regressor, (6890,3,207) posedirs, 300 shape components,
    kintree_table with 2^32-1 as the root's parent - loads through the stand-in unpickler into the arrays SMPL() is built from."""
    import pickle
    import sys
    import types
    import scipy.sparse as sp
    from gaitb200.smpl import load_smpl_data
    mod = types.ModuleType("chumpy"); ch = types.ModuleType("chumpy.ch")

    class Ch:                                             # minimal stand-in for the class the official pickle refers to
        def __init__(self, x):
            self.x = np.asarray(x)

        def __getstate__(self):
            return {"x": self.x, "_dirty_vars": set()}
    Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
    ch.Ch = Ch; mod.ch = ch
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = mod, ch
    try:
        V = 6890
        rng = np.random.default_rng(0)
        shapedirs300 = np.concatenate([smpl_data["shapedirs"], rng.standard_normal((V, 3, 290)).astype(np.float32)], axis=2)
        posedirs_v3p = smpl_data["posedirs"].T.reshape(V, 3, 207)
        kintree = np.stack([smpl_data["parents"].astype(np.int64) % (2 ** 32), np.arange(24)])
        raw = {"v_template": Ch(smpl_data["v_template"]), "shapedirs": Ch(shapedirs300), "posedirs": Ch(posedirs_v3p),
               "J_regressor": sp.csc_matrix(smpl_data["J_regressor"]), "weights": Ch(smpl_data["lbs_weights"]),
               "f": smpl_data["faces"].astype(np.uint32), "kintree_table": kintree.astype(np.uint32), "J": Ch(np.zeros((24, 3)))}
        d = tmp_path / "smpl_data"; d.mkdir()
        with open(d / "SMPL_NEUTRAL.pkl", "wb") as f:
            pickle.dump(raw, f, protocol=2)
        np.save(d / "J_regressor_extra.npy", smpl_data["J_regressor_extra"])
    finally:
        del sys.modules["chumpy"], sys.modules["chumpy.ch"]          # loading must not need chumpy
    got = load_smpl_data(d)
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "J_regressor_extra"):
        assert np.array_equal(np.asarray(got[k], dtype=np.float32), np.asarray(smpl_data[k], dtype=np.float32)), k
    assert np.array_equal(got["parents"], smpl_data["parents"]) and np.array_equal(got["faces"], smpl_data["faces"])
    assert np.array_equal(got["landmark_verts"], smpl_data["landmark_verts"])
