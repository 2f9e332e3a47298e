"""Worker of tests/test_gpu_gather.py (launched with torch.distributed.run): every rank runs its shard of S_total
sequences through sharding.RootGather; the root also runs ALL sequences in one process and compares."""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="nccl")
ap.add_argument("--mode", default="peer-copy")
ap.add_argument("--chunks", type=int, default=2)
ap.add_argument("--smpl-chunks", type=int, default=1)
ap.add_argument("--seqs", type=int, default=10)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--same-gpu", action="store_true", help="all ranks on cuda:0 (gloo + CUDA IPC on a one-GPU box)")
ap.add_argument("--joints-only", action="store_true")
ap.add_argument("--out", default="")
args = ap.parse_args()

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if args.backend == "nccl":
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo", rank=rank, world_size=world)

from gaitb200 import synthetic  # noqa: E402
from gaitb200.head import GaitHead  # noqa: E402
from gaitb200.sharding import RootGather, shard_bounds  # noqa: E402


def make_head():
    return GaitHead(synthetic.make_smpl_data(seed=0, variant="sparse"), synthetic.make_mean_params(),
                    synthetic.make_regressor_state(seed=0, decoder_gain=0.3), synthetic.make_gru_state(seed=0),
                    write_mesh=not args.joints_only).cuda()


S, T = args.seqs, args.frames
feats = synthetic.make_features(S, T, seed=77)                  # identical on every rank; each takes its block
lo, hi = shard_bounds(S, world, rank)
head = make_head()
rg = RootGather(head, S, T, mode=args.mode, chunks=args.chunks, smpl_chunks=args.smpl_chunks)
rg.load_features(feats[lo:hi])
for _ in range(2):                                               # twice: buffers are reused across steps
    rg.run()
torch.cuda.synchronize()
res = None
if rank == 0:
    got = {k: v.clone() for k, v in rg.gathered().items()}
    ref = make_head()(feats.cuda())
    res = {"mode": args.mode, "backend": args.backend, "world": world, "chunks": len(rg.cb), "pieces": len(rg.pieces),
           "root_ingest_bytes": rg.root_ingest_bytes}
    for k, v in got.items():
        res[k + "_max_abs_diff"] = float((v - ref[k]).abs().max())
        res[k + "_bitwise"] = bool(torch.equal(v, ref[k]))
    print("GATHER_RESULT " + json.dumps(res), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(res))
dist.barrier()
rg.close()
dist.destroy_process_group()
