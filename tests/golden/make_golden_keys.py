#!/usr/bin/env python
"""tests/golden/state_dict_keys.json: key -> shape of the state dicts of the REFERENCE's own modules on the path
(spin.Regressor spin.py:210-240, pare.VPRegressor pare.py:24-36, pare.SMPLRegressor pare.py:95-106, smpl.SMPLHead
smpl.py:137-147), instantiated from their unmodified source files behind the import stubs of make_golden.py.
batch_generation.py:210-219 loads the checkpoint with strict=True, so a drop-in module must expose exactly these keys.
The `smpl.*` entries below the reference's wrapper are smplx's buffers; smplx is absent here, so those names are the ones
restated in oracle/smplx_lbs.py from smplx 0.1.26 body_models.SMPL (parity unpinned).

Run here (CPU container, /root/reference mounted):  python tests/golden/make_golden_keys.py"""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402  (stubs + synthetic data)


def main():
    smpl_data = mg.synthetic.make_smpl_data(seed=0, variant="sparse")
    mean = mg.synthetic.make_mean_params()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        d = Path(td) / "data/smpl_data"
        d.mkdir(parents=True)
        np.savez(d / "SMPL_NEUTRAL_synthetic.npz", **smpl_data)
        np.save(d / "J_regressor_extra.npy", smpl_data["J_regressor_extra"])
        np.savez(d / "smpl_mean_params.npz", **mean)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            m = mg._install_stubs(d)
            mods = {"spin.Regressor": m.spin.Regressor(), "pare.VPRegressor": m.pare.VPRegressor(),
                    "pare.SMPLRegressor": m.pare.SMPLRegressor(),
                    "smpl.SMPLHead": m.smpl.SMPLHead(focal_length=5000., img_res=224, smpl_model_dir=str(d))}
            for name, mod in mods.items():
                out[name] = {k: list(v.shape) for k, v in mod.state_dict().items()}
        finally:
            os.chdir(cwd)
    (HERE / "state_dict_keys.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    for k, v in out.items():
        print(k, len(v), "keys")


if __name__ == "__main__":
    assert mg.REF.exists(), "run in the container that mounts /root/reference"
    main()
