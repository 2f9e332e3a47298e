#!/usr/bin/env python
"""Generate tests/golden/heads.npz by running the REFERENCE's own modules for SURVEY.md 8(f) f1 and f4:

  * lib/models/layers/locallyconnected2d.py::LocallyConnected2d, keypoint_attention.py::KeypointAttention
  * lib/models/pare.py::PareHead._get_local_feats / forward / _pare_get_final_preds  (final-prediction MLPs, f1)
  * lib/models/layers/gait_feat_encoder.py::BidirectionalModel (use_pareFeat=True, the only branch that runs as
    shipped: without it `xc` is undefined at :102,104)                                                         (f4)

Modules run unmodified; missing third-party imports are stubbed as in make_golden.py / make_golden_post.py (`timm` too,
which feature_correction.py imports but BidirectionalModel never uses).  Run here:  python tests/golden/make_golden_heads.py
"""
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402
import make_golden_post as mgp  # noqa: E402
from gaitb200 import synthetic  # noqa: E402


def main():
    assert mg.REF.exists()
    mgp._StubFinder.ROOTS = mgp._StubFinder.ROOTS + ("timm", "einops_exts")
    sys.meta_path.insert(0, mgp._StubFinder())
    torch.manual_seed(7)
    out = {}
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
            m = mg._install_stubs(d)
            LC = sys.modules["lib.models.layers.locallyconnected2d"].LocallyConnected2d
            KA = sys.modules["lib.models.layers.keypoint_attention"].KeypointAttention
            with torch.no_grad():
                # ---- LocallyConnected2d, both uses: pose MLP (128 -> 6) and cparam MLP (3 -> 128, expanded input)
                lc = LC(in_channels=128, out_channels=6, output_size=[24, 1], kernel_size=1, stride=1)
                x = torch.randn(5, 128, 24, 1)
                out.update(lc_pose_w=lc.weight.numpy(), lc_pose_x=x.numpy(), lc_pose_y=lc(x).numpy())
                lcb = LC(in_channels=3, out_channels=128, output_size=[24, 1], kernel_size=1, stride=1, bias=True)
                xb = torch.randn(7, 3, 1, 1).expand(7, 3, 24, 1)
                out.update(lc_cp_w=lcb.weight.numpy(), lc_cp_b=lcb.bias.numpy(), lc_cp_x=xb.contiguous().numpy(), lc_cp_y=lcb(xb).numpy())
                # ---- KeypointAttention (softmax, with and without scale)
                feat = torch.relu(torch.randn(3, 128, 14, 14))
                heat = torch.randn(3, 24, 14, 14) * 2
                out.update(ka_feat=feat.numpy(), ka_heat=heat.numpy(), ka_out=KA(act='softmax')(feat, heat).numpy(),
                           ka_out_scaled=KA(act='softmax', use_scale=True)(feat, heat).numpy())
                # ---- PareHead: local features + final predictions (pare.py:261-289, 318-375)
                ph = m.pare.PareHead(num_joints=24, num_input_features=480)
                ph.eval()
                smpl_feats = torch.relu(torch.randn(4, 128, 14, 14))
                part_attn = torch.randn(4, 24, 14, 14)
                plf, csf = ph._get_local_feats(smpl_feats, part_attn)
                o = ph(plf, csf, {})
                out.update(ph_smpl_feats=smpl_feats.numpy(), ph_part_attn=part_attn.numpy(),
                           ph_cam_shape_map=ph.smpl_final_layer(smpl_feats).numpy(),
                           ph_point_local_feat=plf.numpy(), ph_cam_shape_feats=csf.numpy(),
                           ph_pred_rotmat=o['pred_rotmat'].numpy(), ph_pred_cam=o['pred_cam'].numpy(),
                           ph_pred_shape=o['pred_shape'].numpy(), ph_pred_rot6d=o['pred_rot6d'].numpy())
                for k, v in ph.state_dict().items():
                    if k.split(".")[0] in ("pose_mlp", "shape_mlp", "cam_mlp", "smpl_final_layer", "init_pose", "init_shape", "init_cam"):
                        out["ph_sd_" + k] = v.numpy()
                # iterative branch with inits (pare.py:270-276, 347-358)
                ph.iterative_regression = True
                inits = {"pred_rot6d": o['pred_rot6d'], "pred_shape": o['pred_shape'], "pred_cam": o['pred_cam']}
                o2 = ph(plf, csf, {}, inits=inits)
                out.update(ph_it_rot6d=o2['pred_rot6d'].numpy(), ph_it_cam=o2['pred_cam'].numpy(), ph_it_shape=o2['pred_shape'].numpy())
                # ---- BidirectionalModel (gait_feat_encoder.py:10-104)
                __import__("importlib").import_module("lib.models.layers.feature_correction")   # import order of lib/models/layers/__init__.py (circular import)
                gfe = sys.modules["lib.models.layers.gait_feat_encoder"]
                bm = gfe.BidirectionalModel(seqlen=16, use_pareFeat=True)
                bm.eval()
                # weights: gaitb200.synthetic.seeded_state over the reference's own keys / shapes (30 MB, not committed)
                bm.load_state_dict(synthetic.seeded_state({k: v.shape for k, v in bm.state_dict().items()}, seed=5))
                xs = torch.randn(3, 16, 3072) * 0.3
                cp = torch.randn(3, 16, 3)
                y, p, xc = bm(xs, cp)
                out.update(bm_x=xs.numpy(), bm_cparams=cp.numpy(), bm_y=y.numpy(), bm_p=p.numpy(), bm_xc=xc[:, :, ::7].numpy())
                out["bm_state_keys"] = np.array(list(bm.state_dict().keys()))
                out["bm_state_shapes"] = np.array([",".join(map(str, v.shape)) for v in bm.state_dict().values()])
        finally:
            os.chdir(cwd)
    np.savez_compressed(HERE / "heads.npz", **out)
    print("heads.npz written:", {k: v.shape for k, v in out.items() if not k.startswith(("bm_sd_", "ph_sd_"))})
    print("state keys:", [k for k in out if k.startswith(("bm_sd_", "ph_sd_"))])


if __name__ == "__main__":
    main()
