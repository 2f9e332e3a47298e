"""Multi-GPU plumbing: one process per GPU, sequences sharded contiguously, weights replicated.

Sequences never interact anywhere on the path, so there is no data-path
collective: each rank runs the whole head on its own block of sequences.  The only exchange the
north_star names is the final gather of joints / meshes, done here with NCCL
(torch.distributed); it is optional and off the compute path.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(num_seqs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of sequences for `rank`; the first `num_seqs % world_size`
    ranks take one extra."""
    if world_size < 1 or not (0 <= rank < world_size) or num_seqs < 0:
        raise ValueError(f"bad shard request: num_seqs={num_seqs} world_size={world_size} rank={rank}")
    base, rem = divmod(num_seqs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(num_seqs: int, world_size: int) -> list[int]:
    return [shard_bounds(num_seqs, world_size, r)[1] - shard_bounds(num_seqs, world_size, r)[0]
            for r in range(world_size)]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the process
    group (NCCL when CUDA is available, else gloo) when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def gather_sequences(local: torch.Tensor, num_seqs: int, group=None) -> torch.Tensor:
    """All-gather per-rank blocks (S_r, ...) of a sequence-sharded tensor into (num_seqs, ...)
    on every rank, in sequence order.  Even shards use one all_gather_into_tensor; uneven
    shards are padded to the largest block."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != num_seqs:
            raise ValueError(f"single process holds {local.shape[0]} of {num_seqs} sequences")
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(num_seqs, world)
    rank = dist.get_rank(group)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} sequences, expected {counts[rank]}")
    local = local.contiguous()
    if len(set(counts)) == 1:
        out = torch.empty((num_seqs,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    big = max(counts)
    pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host float over all ranks (used for the bench's max-over-ranks step time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------------
# Final gather onto one rank (north_star: "NCCL used only for the final gather of joints and meshes"; the reference's
# boundary it replaces is the per-process .cpu().numpy() at batch_generation.py:316-323 / demo.py:183-188).
class PeerUnavailable(RuntimeError):
    """CUDA IPC mapping of the root's buffer failed on some rank (every rank raises together)."""


def _all_ok(ok: bool, group=None) -> bool:
    """Logical AND of a host flag over the group (works on NCCL and gloo)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


class _RawCuda:
    """__cuda_array_interface__ holder: lets torch view device memory this library allocated (float32, flat)."""

    def __init__(self, addr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(addr), False),
                                         "version": 3, "strides": None}


class PeerBuffer:
    """One device buffer owned by `root` and mapped into the address space of every other rank through CUDA IPC
    (gait_peer_alloc / _export / _open).  `.addr` is the buffer's address in THIS process: local memory on the root, a
    peer mapping elsewhere (kernels store to it / copy engines write to it over NVLink).  Collective: every rank of the
    group constructs it together; raises PeerUnavailable on all ranks if any mapping fails."""

    def __init__(self, nbytes: int, root: int = 0, group=None):
        import ctypes as C
        from . import _lib as L
        self.nbytes, self.root, self.group = int(nbytes), root, group
        self.is_root = dist.get_rank(group) == root
        self.addr, self._owned, self._mapped, self._holder = None, None, None, None
        ok, handle = True, [None]
        if self.is_root:
            try:
                p = C.c_void_p()
                L.call("gait_peer_alloc", C.byref(p), self.nbytes)
                self._owned = p.value
                h = C.create_string_buffer(64)
                L.call("gait_peer_export", self._owned, h)
                handle = [bytes(h.raw)]
            except L.GaitLibraryError as e:
                ok, handle = False, [repr(e)]
        dist.broadcast_object_list(handle, src=root, group=group)
        if self.is_root:
            self.addr = self._owned
        elif isinstance(handle[0], bytes):
            try:
                p = C.c_void_p()
                L.call("gait_peer_open", handle[0], C.byref(p))
                self._mapped = self.addr = p.value
            except L.GaitLibraryError as e:
                ok, self._err = False, repr(e)
        else:
            ok = False
        if not _all_ok(ok, group):
            self.close()
            raise PeerUnavailable(getattr(self, "_err", "CUDA IPC mapping failed on some rank"))

    def tensor(self, offset_bytes: int, shape) -> torch.Tensor:
        """float32 view of the root's own memory (root only)."""
        if not self.is_root:
            raise RuntimeError("PeerBuffer.tensor: only the root holds the buffer as local memory")
        n = 1
        for d in shape:
            n *= int(d)
        self._holder = self._holder or []
        h = _RawCuda(self.addr + offset_bytes, n)
        self._holder.append(h)
        return torch.as_tensor(h, device=torch.device("cuda", torch.cuda.current_device())).view(*shape)

    def close(self):
        from . import _lib as L
        lib = L.load()
        if self._mapped:
            lib.gait_peer_close(self._mapped)
            self._mapped = None
        if self._owned:
            torch.cuda.synchronize()
            lib.gait_peer_free(self._owned)
            self._owned = None
        self.addr = None


GATHER_MODES = ("nccl", "peer-copy", "peer-store")


class RootGather:
    """Sequence-sharded run of a GaitHead with the final gather of meshes + Kinect-25 joints onto `root` INSIDE the step.

    Every rank holds S_local = shard of S_total sequences and runs them as `chunks` consecutive sub-batches; the outputs
    of chunk c travel to the root while chunk c+1 computes:
      'peer-store' : the skinning kernel's epilogue stores go straight into the root's buffer (verts = peer address, the
                     collective IS the kernel's own coalesced stores over NVLink); joints by a small peer copy
      'peer-copy'  : the chunk's mesh is written locally, then put into the root's buffer by an asynchronous peer copy on a
                     second stream (copy engines over NVLink; no SM touches the link)
      'nccl'       : grouped ncclSend/ncclRecv (torch.distributed.batch_isend_irecv) per chunk, waited at the end
    smpl_chunks > 1 splits only the part AFTER the regressor (chain, blend, skinning, joints) of every sequence chunk: the
    encoder + regressor run once over the whole chunk (their GEMMs are more efficient on many rows), the meshes are produced
    and leave in smpl_chunks pieces.
    The root computes its own shard directly into its slice of the gathered buffers in every mode.  A step ends with a
    completion signal (a 1-element all-reduce on the launching stream; host barrier on gloo), after which the root may read
    `gathered()`: {'verts': (S_total,T,V,3), 'kinect25': (S_total,T,25,3)} (joints-only heads: 'kinect25' only)."""

    def __init__(self, head, S_total: int, T: int, mode: str = "peer-copy", chunks: int = 1, root: int = 0, group=None,
                 use_graphs: bool = True, smpl_chunks: int = 1):
        from . import _lib as L
        if mode not in GATHER_MODES:
            raise ValueError(f"mode must be one of {GATHER_MODES}")
        if not dist.is_initialized():
            raise RuntimeError("RootGather needs an initialised process group")
        self.head, self.mode, self.root, self.group, self.T = head, mode, root, group, T
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.smpl_chunks = max(1, smpl_chunks)
        self.is_root = self.rank == root
        self.S_total = S_total
        self.lo, self.hi = shard_bounds(S_total, self.world, self.rank)
        self.S_local = self.hi - self.lo
        self.mesh = bool(head.write_mesh)
        dev = head.regressor.fc1.weight.device
        self.dev = dev
        V = head.regressor.smpl.v_template.shape[0]
        self.V = V
        self.backend = dist.get_backend(group)
        chunks = max(1, min(chunks, min(shard_counts(S_total, self.world))))              # same count on every rank, none empty
        if self.S_local < 1:
            raise ValueError(f"rank {self.rank} got no sequence ({S_total} sequences over {self.world} ranks)")
        self.cb = [shard_bounds(self.S_local, chunks, c) for c in range(chunks)]        # chunk bounds inside the shard
        per_seq_mesh, per_seq_kin = T * V * 3, T * 25 * 3
        mesh_bytes = S_total * per_seq_mesh * 4 if self.mesh else 0
        kin_off = (mesh_bytes + 255) // 256 * 256
        total = kin_off + S_total * per_seq_kin * 4
        self.peer = None
        if mode == "nccl":
            if self.is_root:
                self._store = torch.empty(total // 4, device=dev, dtype=torch.float32)
                base = self._store.data_ptr()
                view = lambda off, shape: self._store[off // 4: off // 4 + _numel(shape)].view(*shape)
        else:
            self.peer = PeerBuffer(total, root, group)
            base = self.peer.addr
            if self.is_root:
                view = self.peer.tensor
        if self.is_root:
            self.verts_all = view(0, (S_total, T, V, 3)) if self.mesh else None
            self.kinect_all = view(kin_off, (S_total, T, 25, 3))
        self._mesh_addr = (lambda s: base + s * per_seq_mesh * 4) if (mode != "nccl" or self.is_root) else None
        self._kin_addr = (lambda s: base + kin_off + s * per_seq_kin * 4) if (mode != "nccl" or self.is_root) else None
        # one unit per sequence chunk: a full plan, or a front plan (encoder + regressor) with smpl_chunks SMPL sub-plans.
        # where the mesh of a piece is written: root -> its slice of the gathered buffer; peer-store -> the root's slice through
        # the mapping; else -> the plan's own buffer (then copied).
        def dest(seq):
            ok = self.mesh and (self.is_root or mode == "peer-store") and (self._mesh_addr(self.lo + seq) % 8 == 0)
            return self._mesh_addr(self.lo + seq) if ok else None

        self.units, self.pieces = [], []              # pieces: (plan, graph, first local sequence) in sequence order
        smpl_chunks = max(1, smpl_chunks)
        launches, first = 0, True
        for (a, b) in self.cb:
            k = min(smpl_chunks, b - a)
            if k <= 1:
                p = head._make_plan(b - a, T, verts_addr=dest(a))
                g = head.capture_plan(p, warm=first) if use_graphs else None
                launches += head.launches_per_step if use_graphs else 0
                unit = {"front": None, "fgraph": None, "x": p["x"], "a": a, "b": b, "pieces": [(p, g, a)]}
            else:
                front = head._make_plan(b - a, T, front_only=True)
                fg = head.capture_plan(front, warm=first, part="front") if use_graphs else None
                launches += head.launches_per_step if use_graphs else 0
                unit = {"front": front, "fgraph": fg, "x": front["x"], "a": a, "b": b, "pieces": []}
                for i in range(k):
                    sa, sb = shard_bounds(b - a, k, i)
                    sp = head._make_plan(sb - sa, T, verts_addr=dest(a + sa), parent=front, seq_off=sa)
                    if first and use_graphs:
                        head._launch(front, "front")       # the sub-plan's warm-up reads the regressor state
                    sg = head.capture_plan(sp, warm=first, part="smpl") if use_graphs else None
                    launches += head.launches_per_step if use_graphs else 0
                    unit["pieces"].append((sp, sg, a + sa))
            first = False
            self.units.append(unit)
            self.pieces += unit["pieces"]
        self.plans = [pc[0] for pc in self.pieces]
        self.launches_per_step = launches if use_graphs else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._flag = torch.zeros(1, device=dev, dtype=torch.float32)
        self._L = L
        # bytes this rank sends per step / the root receives per step
        self.sent_bytes = 0 if self.is_root else self.S_local * (per_seq_kin + (per_seq_mesh if self.mesh else 0)) * 4
        self.root_ingest_bytes = (S_total - shard_counts(S_total, self.world)[root]) * (per_seq_kin + (per_seq_mesh if self.mesh else 0)) * 4

    # ------------------------------------------------------------------
    def load_features(self, x_local: torch.Tensor):
        """Copy this rank's (S_local,T,2048) features (host or device) into the chunk input buffers."""
        if x_local.shape[0] != self.S_local:
            raise ValueError(f"rank {self.rank} expects {self.S_local} sequences, got {x_local.shape[0]}")
        for u in self.units:
            u["x"].copy_(x_local[u["a"]:u["b"]], non_blocking=True)

    def _piece_bounds(self, n_local):
        """[(a, b)] of every piece of a shard of n_local sequences, in order (same rule on every rank)."""
        out = []
        for c in range(len(self.cb)):
            a, b = shard_bounds(n_local, len(self.cb), c)
            k = min(self.smpl_chunks, b - a)
            for i in range(max(k, 1)):
                sa, sb = shard_bounds(b - a, max(k, 1), i)
                out.append((a + sa, a + sb))
        return out

    def _nccl_ops(self, i):
        """send / recv operations of piece i (the i-th piece of every rank's shard)."""
        p = self.plans[i]
        if self.is_root:
            ops = []
            for r in range(self.world):
                if r == self.root:
                    continue
                rlo, rhi = shard_bounds(self.S_total, self.world, r)
                ca, cb_ = self._piece_bounds(rhi - rlo)[i]
                ops.append(dist.P2POp(dist.irecv, self.kinect_all[rlo + ca: rlo + cb_], r, self.group))
                if self.mesh:
                    ops.append(dist.P2POp(dist.irecv, self.verts_all[rlo + ca: rlo + cb_], r, self.group))
            return ops
        ops = [dist.P2POp(dist.isend, p["kinect"], self.root, self.group)]
        if self.mesh:
            ops.append(dist.P2POp(dist.isend, p["verts"], self.root, self.group))
        return ops

    @torch.no_grad()
    def run(self):
        """One step: all chunks of this rank's shard + the gather + the completion signal, enqueued on the current stream
        (plus the copy stream / NCCL's stream); returns without host synchronisation on NCCL."""
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        works = []
        i = 0
        for u in self.units:
            if u["front"] is not None:
                if u["fgraph"] is not None:
                    u["fgraph"].replay()
                else:
                    self.head._launch(u["front"], "front")
            for (p, g, a) in u["pieces"]:
                if g is not None:
                    g.replay()
                else:
                    self.head._launch(p, "smpl" if u["front"] is not None else "all")
                if self.is_root:
                    # own Kinect-25 joints into the gathered buffer (the mesh was written in place unless its slice is misaligned)
                    self._L.call("gait_peer_copy", self._kin_addr(self.lo + a), p["kinect"].data_ptr(), p["kinect"].numel() * 4, cur.cuda_stream)
                    if self.mesh and p["verts"] is not None:
                        self._L.call("gait_peer_copy", self._mesh_addr(self.lo + a), p["verts"].data_ptr(), p["verts"].numel() * 4, cur.cuda_stream)
                elif self.mode != "nccl":
                    ev = cur.record_event()
                    cs.wait_event(ev)
                    self._L.call("gait_peer_copy", self._kin_addr(self.lo + a), p["kinect"].data_ptr(), p["kinect"].numel() * 4, cs.cuda_stream)
                    if self.mesh and p["verts"] is not None:       # peer-copy, or a peer-store piece whose slice is misaligned
                        self._L.call("gait_peer_copy", self._mesh_addr(self.lo + a), p["verts"].data_ptr(), p["verts"].numel() * 4, cs.cuda_stream)
                if self.mode == "nccl" and self.world > 1:
                    ops = self._nccl_ops(i)
                    if ops:
                        works += dist.batch_isend_irecv(ops)
                i += 1
        for w in works:
            w.wait()
        cur.wait_stream(cs)
        self._signal()

    def _signal(self):
        if self.backend == "nccl":
            dist.all_reduce(self._flag, group=self.group)          # stream-ordered after this rank's stores / copies
        else:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def gathered(self):
        if not self.is_root:
            return None
        out = {"kinect25": self.kinect_all}
        if self.mesh:
            out["verts"] = self.verts_all
        return out

    def local_outputs(self):
        """This rank's small per-frame outputs, chunk by chunk concatenated: {'kinect25','kp_3d','rotmat','theta','kp_2d'}."""
        keys = {"kinect25": "kinect", "kp_3d": "joints", "rotmat": "rotmat", "theta": "theta", "kp_2d": "kp2d"}
        return {k: torch.cat([p[v].view(p["S"], p["T"], *p[v].shape[1:]) for p in self.plans], 0) for k, v in keys.items()}

    def close(self):
        torch.cuda.synchronize()
        self.units, self.pieces, self.plans = [], [], []
        if self.peer is not None:
            self.peer.close()


def _numel(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n
