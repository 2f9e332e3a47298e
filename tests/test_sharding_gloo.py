"""N>1 host logic on CPU: sequence sharding and the final gather over a world-size-2 gloo group
(the GPU path uses the same code over NCCL; sequences are independent so there is no data-path
collective to test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaitb200.sharding import gather_sequences, max_over_ranks, shard_bounds, shard_counts


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            bounds = [shard_bounds(n, w, r) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))       # contiguous, no overlap
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard_counts(n, w)
    assert shard_bounds(1024, 8, 3) == (384, 512)                                    # C3: 128 sequences per GPU
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_seqs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(num_seqs, world, rank)
        full = torch.arange(num_seqs * 4 * 25 * 3, dtype=torch.float32).reshape(num_seqs, 4, 25, 3)   # (S,T,25,3) joints
        got = gather_sequences(full[lo:hi].clone(), num_seqs)
        ok = torch.equal(got, full)
        slow = max_over_ranks(float(rank + 1), device="cpu")
        q.put((rank, ok, slow))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_seqs", [8, 7])        # even shards (all_gather_into_tensor) and ragged shards (padded)
def test_gather_over_gloo_world2(num_seqs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_seqs, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, slow in results:
        assert ok, f"rank {rank}: gathered tensor differs"
        assert slow == 2.0


def test_gather_single_process_checks_counts():
    x = torch.zeros(4, 2, 25, 3)
    assert gather_sequences(x, 4) is x
    with pytest.raises(ValueError):
        gather_sequences(x, 5)
    assert max_over_ranks(3.5) == 3.5


def test_root_gather_piece_bounds_cover_every_shard():
    """RootGather cuts a rank's shard into sequence chunks x SMPL pieces; the root derives every OTHER rank's piece bounds from
    the same rule (that is what its receive slices are built from).  The pieces must tile each shard exactly, for even and
    ragged shards."""
    from gaitb200.sharding import RootGather
    for S_total, world, chunks, smpl in ((1024, 8, 2, 4), (1024, 2, 1, 8), (10, 2, 2, 1), (9, 2, 1, 3), (7, 3, 2, 2)):
        chunks_eff = max(1, min(chunks, min(shard_counts(S_total, world))))
        for r in range(world):
            lo, hi = shard_bounds(S_total, world, r)
            fake = type("F", (), {})()                           # only the two attributes _piece_bounds reads
            fake.cb = [shard_bounds(hi - lo, chunks_eff, c) for c in range(chunks_eff)]
            fake.smpl_chunks = smpl
            pieces = RootGather._piece_bounds(fake, hi - lo)
            assert pieces[0][0] == 0 and pieces[-1][1] == hi - lo
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:])) and all(b > a for a, b in pieces)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the arm the driver times beside ours) runs on CPU only and prints one JSON line with the
    contract's keys; at N > 1 only rank 0 prints and the workload is all of BASELINE configs[2] (here shrunk by --seqs-per-gpu)."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    env = dict(os.environ, WORLD_SIZE="2", RANK="0", LOCAL_RANK="0", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--seqs-per-gpu", "2", "--frames", "4"], capture_output=True, text=True, timeout=300, env=env, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["gpu_launches"] == 0
    assert line["config"]["global_frames_per_step"] == 2 * 2 * 4 and "configs[2]" in line["config"]["workload"]
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--seqs-per-gpu", "2", "--frames", "4"], capture_output=True, text=True, timeout=300, env=env, cwd=str(root))
    assert r.returncode == 0 and r.stdout.strip() == ""
