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
