"""N-rank == 1-rank: the sequence-sharded run with the final gather onto the root (sharding.RootGather) reproduces the
single-process output.  One-GPU boxes run two processes on cuda:0 (gloo for the control plane, CUDA IPC for the data:
the peer-store / peer-copy paths are exactly the multi-GPU ones); with >= 2 GPUs the NCCL variants run as well."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _run(tmp_path, nproc, port, *extra):
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "_gather_worker.py"), "--out", str(out), *extra]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(out.read_text())


def _check(res, mesh=True):
    # sequences are independent and every rank runs the same kernels on its rows: the gathered result must agree with the
    # one-process run to FP32 rounding of different GEMM row tilings (tolerances of north_star / 10)
    assert res["kinect25_max_abs_diff"] <= 1e-5, res
    if mesh:
        assert res["verts_max_abs_diff"] <= 1e-5, res


@pytest.mark.parametrize("mode", ["peer-copy", "peer-store"])
def test_two_ranks_one_gpu_ipc_gather_matches_single_process(tmp_path, mode):
    res = _run(tmp_path, 2, 29631 + (mode == "peer-store"), "--backend", "gloo", "--same-gpu", "--mode", mode, "--chunks", "2",
               "--seqs", "10", "--frames", "8")
    assert res["world"] == 2 and res["chunks"] == 2
    _check(res)


def test_two_ranks_one_gpu_smpl_chunks(tmp_path):
    """One encoder + regressor pass per rank, the SMPL part and the gather in 3 pieces (odd frame count, ragged pieces)."""
    res = _run(tmp_path, 2, 29634, "--backend", "gloo", "--same-gpu", "--mode", "peer-store", "--chunks", "1", "--smpl-chunks", "3",
               "--seqs", "9", "--frames", "5")
    assert res["pieces"] == 3
    _check(res)


def test_two_ranks_one_gpu_joints_only_gather(tmp_path):
    res = _run(tmp_path, 2, 29633, "--backend", "gloo", "--same-gpu", "--mode", "peer-copy", "--chunks", "1", "--seqs", "7",
               "--frames", "16", "--joints-only")
    _check(res, mesh=False)


@pytest.mark.parametrize("mode", ["nccl", "peer-copy", "peer-store"])
def test_nccl_ranks_gather_matches_single_process(tmp_path, mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    res = _run(tmp_path, 2, 29640 + ["nccl", "peer-copy", "peer-store"].index(mode), "--backend", "nccl", "--mode", mode, "--chunks", "2",
               "--smpl-chunks", "2", "--seqs", "12", "--frames", "16")
    _check(res)
