"""GaitHead - the whole hot path as one object: backbone features -> TemporalEncoder ->
Regressor (MLP loop, rot6d->R, SMPL, projection, axis-angle) -> Kinect-25 joints.

It owns a TemporalEncoder and a Regressor (the reference-API modules; their parameters and
buffers are the single source of truth) and runs them for a FIXED (S, T) through the C-ABI
with every intermediate pre-allocated, so a step is a fixed sequence of kernel launches on one
stream - which is what gets captured into a CUDA graph and replayed.  Output keys/shapes follow
spin.Regressor's dict (spin.py:288-295) reshaped to (S,T,...), plus 'kinect25'
(convert_kps(kp_3d,'spin2','kinectv2'), batch_generation.py:323).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .kp_utils import SPIN2_TO_KINECTV2
from .regressor import Regressor
from .temporal import TemporalEncoder

_STATE_LD = 160


# small per-frame outputs packed into one buffer: (plan key, output key, per-frame shape); segments start 16-byte aligned
_SMALL = (("rotmat", "rotmat", (24, 3, 3)), ("joints", "kp_3d", (29, 3)), ("kp2d", "kp_2d", (29, 2)), ("kinect", "kinect25", (25, 3)),
          ("theta", "theta", (85,)))


def _small_layout(F: int):
    """[(plan key, output key, shape, offset, numel)], total floats."""
    out, off = [], 0
    for key, name, shape in _SMALL:
        n = F
        for d in shape:
            n *= d
        out.append((key, name, shape, off, n))
        off += (n + 3) // 4 * 4
    return out, max(off, 4)


class HostOutputs(dict):
    """dict of pinned host tensors from GaitHead.alloc_host_outputs(); `.packed` is the single buffer behind all of them."""
    packed = None


class GaitHead(nn.Module):
    def __init__(self, smpl_data, mean_params, regressor_state=None, gru_state=None, write_mesh=True,
                 n_iter=3, fold_regressor=False, joints_mode="reduced", **encoder_kw):
        """fold_regressor: run the regressor loop as its folded affine map (Regressor.fold; opt-in, same outputs up to FP32
        rounding, the iteration's GEMMs are gone - not the default and not what bench.py's headline measures).
        joints_mode (write_mesh=False only): "reduced" - the Kinect-25 set needs 21 landmark vertices and one regressor row
        (thorax); skinning is linear in v_posed, so that row is folded through the skinning weights once per model and NO
        vertex other than the landmarks is ever formed (SMPL._prepare_reduced, gait_smpl_reduced_joints); "skin" - every
        vertex is blended and skinned on chip and only the landmarks / the thorax partials are written."""
        super().__init__()
        if joints_mode not in ("reduced", "skin"):
            raise ValueError("joints_mode must be 'reduced' or 'skin'")
        self.joints_mode = joints_mode
        self.fold_regressor = bool(fold_regressor)
        self.encoder = TemporalEncoder(**encoder_kw)
        self.regressor = Regressor(mean_params, smpl_data)
        if gru_state is not None:
            self.encoder.gru.load_state_dict(gru_state)
        if regressor_state is not None:
            self.regressor.load_state_dict(regressor_state, strict=False)
        g = self.encoder.gru
        if g.num_layers != 1 or g.bidirectional or self.encoder.linear is not None or g.hidden_size != g.input_size:
            raise L.GaitLibraryError("GaitHead fuses the 1-layer unidirectional residual TemporalEncoder; "
                                     "use TemporalEncoder + Regressor separately for other encoders")
        self.n_iter = n_iter
        self.write_mesh = write_mesh
        self.eval()
        self._plan = None
        self._graph = None
        self._slots, self._graphs = [], []

    # ------------------------------------------------------------------ buffers for one (S,T)
    def _make_plan(self, S: int, T: int, verts_addr: int | None = None, parent=None, seq_off: int = 0, front_only: bool = False):
        """Buffers of one (S,T) step.  verts_addr: raw device address that receives the mesh (F,V,3) instead of the plan's own
        output buffer - a slice of a gathered buffer, possibly PEER memory of the root rank (sharding.RootGather): the
        skinning kernel then also writes the landmark vertices to a local buffer, so that nothing reads the remote mesh.
        parent / seq_off: a SUB-plan for sequences [seq_off, seq_off + S) of `parent`: it runs only the SMPL stages
        (_stages(..., part="smpl")) on the parent's regressor state, so one encoder + regressor pass over a whole shard can be
        followed by several SMPL passes whose outputs leave one after the other (overlapping the gather with compute).
        front_only: the parent of such sub-plans - encoder + regressor buffers only (no mesh-sized allocations)."""
        dev = self.regressor.fc1.weight.device
        if dev.type != "cuda":
            raise L.GaitLibraryError("GaitHead is on %s; move it to a CUDA device (no CPU path)" % dev)
        L.require_device()
        lib = L.load()
        F = S * T
        H = self.encoder.gru.hidden_size
        V = self.regressor.smpl.v_template.shape[0]
        e = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        gru_bytes = lib.gait_gru_workspace_bytes(S, T, H)
        hmr_bytes = lib.gait_hmr_workspace_bytes(F, self.regressor.fc1.out_features)
        sub = parent is not None
        F_all, F = F, (0 if front_only else F)          # SMPL-stage buffers are empty in a front-only plan
        p = {
            "S": S, "T": T, "F": F_all, "V": V, "dev": dev,
            "x": None if sub else e(S, T, H), "y_raw": None if sub else e(S, T, H), "enc": None if sub else e(S, T, H),
            "ws": None if sub else e(max(gru_bytes, hmr_bytes, 4) // 4), "gru_bytes": gru_bytes, "hmr_bytes": hmr_bytes,
            "state": parent["state"][seq_off * T:(seq_off + S) * T] if sub else e(F_all, _STATE_LD), "Jp": e(F, 24, 3),
            "aop": e(lib.gait_smpl_lbs_aop_bytes(F) // 4),
            "coef": e(F, 224), "v_posed": e(F, 384 * ((V + 127) // 128)),
            # full mesh (F,V,3), or in joints-only mode just the landmark vertices the joint sets read (config 5)
            "verts": None, "verts_addr": None,
            "extra": e(4 * ((V + 127) // 128), F, 1, 3),      # gait_smpl_lbs_jx_parts(V) partial sums of the thorax row
            "gather": torch.tensor(SPIN2_TO_KINECTV2, dtype=torch.int32, device=dev),
        }
        if verts_addr is not None and not self.write_mesh:
            raise L.GaitLibraryError("verts_addr given to a joints-only head")
        self_reduced = (not self.write_mesh) and self.joints_mode == "reduced"
        if self_reduced:
            rk = self.regressor.smpl._prepare_reduced()
            p["A12"] = e(F, 24, 12)                       # skinning transforms as plain 3x4 rows (the reduced path has no tensor-core operand)
            p["red_u"] = e(F, rk["red_ld"])
            p["v_posed"] = e(0)                           # never formed
            p["extra"] = e(1, F, 1, 3)
        external = verts_addr is not None
        local_lm = external or not self.write_mesh
        n_lm = self.regressor.smpl._prepare()["n_landmarks"]
        p["lm_verts"] = e(F, n_lm, 3) if local_lm else None
        p["lm_iota"] = torch.arange(n_lm, dtype=torch.int32, device=dev) if local_lm else None
        # every output of a step lives in ONE buffer [mesh | small per-frame outputs], so the host copy is one transfer
        nm = 0 if external else self._mesh_floats(F, V)
        p["outbuf"] = e(nm + _small_layout(F)[1])
        if external:
            if verts_addr % 8:
                raise L.GaitLibraryError("verts_addr must be 8-byte aligned")
            p["verts_addr"] = int(verts_addr)
        elif self.write_mesh:
            p["verts"] = p["outbuf"][:F * V * 3].view(F, V, 3)
            p["verts_addr"] = p["verts"].data_ptr()
        p["small"] = p["outbuf"][nm:]
        for key, _, shape, off, n in _small_layout(F)[0]:
            p[key] = p["small"][off:off + n].view(F, *shape)
        return p

    def _mesh_floats(self, F, V):
        """floats reserved for the mesh at the front of the output buffer (16-byte multiple; 0 in joints-only mode)"""
        return (F * V * 3 + 3) // 4 * 4 if self.write_mesh else 0

    def alloc_host_outputs(self):
        """Pinned host buffers for run_host_batches: a dict with the keys/shapes of outputs(); the small outputs are views of
        one pinned buffer laid out like the device-side one, so they arrive with a single D2H transfer."""
        p = self._plan
        S, T, F = p["S"], p["T"], p["F"]
        layout, total = _small_layout(F)
        nm = self._mesh_floats(F, p["V"])
        buf = torch.empty(nm + total).pin_memory()
        out = HostOutputs()
        out.packed = buf
        small = buf[nm:]
        for _, name, shape, off, n in layout:
            out[name] = small[off:off + n].view(S, T, *shape)
        if self.write_mesh:
            out["verts"] = buf[:F * p["V"] * 3].view(S, T, p["V"], 3)
        return out

    def plan(self, S: int, T: int, slots: int = 1):
        """Allocate `slots` independent buffer sets for (S,T) (slot 0 is the default one; a second
        slot lets the D2H copy of one batch overlap the kernels of the next, see run_host_batches)."""
        self._slots = [self._make_plan(S, T) for _ in range(max(1, slots))]
        self._graphs = [None] * len(self._slots)
        self._plan = self._slots[0]
        self._graph = None
        return self._plan

    _FRONT = ("gru", "regressor")

    def _stages(self, p, part: str = "all"):
        """The step as an ordered list of (name, thunk); each thunk enqueues one stage on the
        current stream through the C-ABI (no allocation, no torch op).  part: "all", "front" (encoder + regressor) or
        "smpl" (everything after the regressor state: chain, blend, skinning, joints)."""
        if part != "all":
            return [(n, f) for n, f in self._stages(p) if (n in self._FRONT) == (part == "front")]
        reg, smpl, gru = self.regressor, self.regressor.smpl, self.encoder.gru
        rk, sk = reg._prepare(), smpl._prepare()
        if not (smpl.extra and smpl.kinectv2):
            raise L.GaitLibraryError("GaitHead emits the Kinect-25 set; SMPL.extra and SMPL.kinectv2 must be True")
        S, T, F, V = p["S"], p["T"], p["F"], p["V"]
        H = gru.hidden_size
        ptr, call, st = L.ptr, L.call, L.stream_ptr
        state = p["state"].data_ptr()
        betas, cam = state + 4 * 144, state + 4 * 154
        L.prepare_weight(gru.weight_ih_l0)
        if T > 1 and L.load().gait_gru_plan(S, T, H) == 0:
            L.prepare_weight(gru.weight_hh_l0)        # per-step recurrence (many sequences): T-1 GEMMs over W_hh per step
        fk = reg.fold(self.n_iter) if self.fold_regressor else None
        if "red_u" in p:
            # joints-only without the mesh: chain (plain transforms), one small GEMM, landmark skinning + thorax, assembly
            rd = smpl._prepare_reduced()
            front = self._front_stages(p, fk, rk, H, S, T, F)
            return front + [
                ("pose_chain", lambda: call(
                    "gait_smpl_pose_chain_rot6d", state, _STATE_LD, 1e-6, betas, _STATE_LD, cam, _STATE_LD, ptr(sk["J_template"]),
                    ptr(sk["J_shapedirs"]), ptr(sk["parents"]), ptr(p["rotmat"]), ptr(p["A12"]), ptr(p["Jp"]), ptr(p["coef"]), None,
                    ptr(p["theta"]), F, st())),
                ("blend", lambda: call(
                    "gait_linear", ptr(p["coef"]), 224, ptr(rd["red_basis"]), 224, None, None, 0, ptr(p["red_u"]), rd["red_ld"],
                    F, rd["red_ld"], 224, st())),
                ("lbs", lambda: call(
                    "gait_smpl_reduced_joints", ptr(p["A12"]), ptr(p["red_u"]), rd["red_ld"], ptr(rd["lm_weights"]), ptr(rd["red_s"]),
                    ptr(p["lm_verts"]), ptr(p["extra"]), F, sk["n_landmarks"], st())),
                ("joints", lambda: call(
                    "gait_joints_assemble", ptr(p["Jp"]), ptr(p["lm_verts"]), sk["n_landmarks"], ptr(p["lm_iota"]), sk["n_landmarks"],
                    ptr(p["extra"]), 1, 1, F * 3, ptr(sk["map_kinect"]), 29, ptr(p["joints"]), cam, _STATE_LD,
                    5000., 224., 112.,
                    ptr(p["kp2d"]), ptr(p["gather"]), 25, ptr(p["kinect"]), F, st())),
            ]
        return self._front_stages(p, fk, rk, H, S, T, F) + [
            # rot6d -> R, kinematic chain, blend coefficients, skinning operand and theta in ONE launch
            ("pose_chain", lambda: call(
                "gait_smpl_pose_chain_rot6d", state, _STATE_LD, 1e-6, betas, _STATE_LD, cam, _STATE_LD, ptr(sk["J_template"]),
                ptr(sk["J_shapedirs"]), ptr(sk["parents"]), ptr(p["rotmat"]), None, ptr(p["Jp"]), ptr(p["coef"]), ptr(p["aop"]),
                ptr(p["theta"]), F, st())),
            ("blend", lambda: call(
                "gait_smpl_blend", ptr(p["coef"]), ptr(sk["basis_t"]), ptr(p["v_posed"]), sk["ldv"], F, 3 * V, st())),
            # mesh to p["verts_addr"] (own buffer, or an external / peer address together with local landmark vertices),
            # or landmark vertices only (joints-only mode)
            ("lbs", lambda: call(
                "gait_smpl_lbs_tc_ex", ptr(p["v_posed"]), sk["ldv"], ptr(p["aop"]), ptr(sk["lbs_wpack"]),
                ptr(sk["extra_thorax"]), p["verts_addr"], ptr(p["extra"]),
                ptr(sk["landmarks"]) if p["lm_verts"] is not None else None, sk["n_landmarks"] if p["lm_verts"] is not None else 0,
                ptr(p["lm_verts"]), F, V, st())),
            ("joints", (lambda: call(
                "gait_joints_assemble", ptr(p["Jp"]), p["verts_addr"], V, ptr(sk["landmarks"]), sk["n_landmarks"],
                ptr(p["extra"]), 1, sk["vtiles"], F * 3, ptr(sk["map_kinect"]), 29, ptr(p["joints"]), cam, _STATE_LD,
                5000., 224., 112.,
                ptr(p["kp2d"]), ptr(p["gather"]), 25, ptr(p["kinect"]), F, st())) if p["lm_verts"] is None else (lambda: call(
                "gait_joints_assemble", ptr(p["Jp"]), ptr(p["lm_verts"]), sk["n_landmarks"], ptr(p["lm_iota"]), sk["n_landmarks"],
                ptr(p["extra"]), 1, sk["vtiles"], F * 3, ptr(sk["map_kinect"]), 29, ptr(p["joints"]), cam, _STATE_LD,
                5000., 224., 112.,
                ptr(p["kp2d"]), ptr(p["gather"]), 25, ptr(p["kinect"]), F, st()))),
        ]

    def _front_stages(self, p, fk, rk, H, S, T, F):
        """encoder + regressor stages (shared by the full-mesh / skinned and the reduced joints-only step)"""
        gru = self.encoder.gru
        ptr, call, st = L.ptr, L.call, L.stream_ptr
        return [
            ("gru", lambda: call(
                "gait_gru_layer", ptr(p["x"]), H, ptr(gru.weight_ih_l0), ptr(gru.weight_hh_l0), ptr(gru.bias_ih_l0),
                ptr(gru.bias_hh_l0), None, ptr(p["y_raw"]), H, ptr(p["x"]), H, ptr(p["enc"]), H, None, S, T, H, H, 0,
                ptr(p["ws"]), p["gru_bytes"], st())),
            ("regressor", (lambda: call(
                "gait_hmr_regressor_folded", ptr(p["enc"]), H, ptr(fk["Wf"]), ptr(fk["bf"]), ptr(p["state"]), F, rk["din"],
                ptr(p["ws"]), p["hmr_bytes"], st())) if self.fold_regressor else lambda: call(
                "gait_hmr_regressor", ptr(p["enc"]), H, ptr(rk["W1x"]), ptr(rk["W1s"]), ptr(rk["b1"]), ptr(rk["W2"]),
                ptr(rk["b2"]), ptr(rk["Wd"]), ptr(rk["bd"]), ptr(rk["init"]), 1, self.n_iter, ptr(p["state"]), F,
                rk["din"], rk["dh"], ptr(p["ws"]), p["hmr_bytes"], st())),
        ]

    def _launch(self, p, part: str = "all"):
        """Enqueue one step (or its "front" / "smpl" part) on the current stream."""
        for _, fn in self._stages(p, part):
            fn()

    @torch.no_grad()
    def profile_stages(self, iters: int = 20, flush=None):
        """Per-stage device time (ms, mean over `iters`) with CUDA events on the launching stream and
        the number of kernel launches each stage makes.  `flush()` (e.g. an L2 flush) runs before
        every timed stage, outside the event pair."""
        p = self._plan
        stages = self._stages(p)
        self._launch(p)
        torch.cuda.synchronize()
        res = {}
        for name, fn in stages:
            tot, n0 = 0.0, L.launch_count()
            for _ in range(iters):
                if flush is not None:
                    flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                tot += a.elapsed_time(b)
            res[name] = {"ms": tot / iters, "launches": (L.launch_count() - n0) // iters}
        return res

    @torch.no_grad()
    def time_stage_back_to_back(self, name: str, launches: int = 8, repeats: int = 5):
        """Average duration (ms) of one launch of stage `name`, timed as `launches` consecutive launches between
        ONE pair of CUDA events, alternating between the planned buffer slots (with two slots the LBS stage
        touches 2 x 170 MB > the 126 MB L2, so no launch finds its operands cached).  A single launch bracketed
        by its own event pair also pays ~5 us of event / launch gap, which is not kernel time."""
        if len(self._slots) < 2:
            raise L.GaitLibraryError("time_stage_back_to_back needs plan(..., slots=2)")
        for p in self._slots:
            self._launch(p)
        fns = [dict(self._stages(p))[name] for p in self._slots]
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(repeats):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(launches):
                fns[i % len(fns)]()
            b.record()
            b.synchronize()
            best = min(best, a.elapsed_time(b) / launches)
        return best

    def capture_plan(self, p, warm: bool = True, part: str = "all"):
        """Capture one step (or one part of it) over the buffers of plan `p` into a CUDA graph (returns it; sets
        launches_per_step to the launches of what was captured)."""
        if warm:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._launch(p, part)            # warm-up outside capture (lazy module loading)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        n0 = L.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch(p, part)
        self.launches_per_step = L.launch_count() - n0
        return g

    def capture(self, S: int, T: int, slots: int = 1):
        """Plan buffers for (S,T) and capture one step per slot into a CUDA graph."""
        self.plan(S, T, slots)
        for i, p in enumerate(self._slots):
            self._graphs[i] = self.capture_plan(p, warm=(i == 0))
        self._graph = self._graphs[0]
        return self._graph

    @torch.no_grad()
    def run_host_batches(self, inputs, outputs, sync: bool = True):
        """End-to-end path for HOST data: for every batch, pinned features (S,T,2048) -> H2D -> one step ->
        D2H of every output into the matching dict of pinned host tensors.  With two planned slots the
        three phases of consecutive batches overlap on separate streams (copy-in, compute, copy-out);
        with sync=True (default) the host waits for the last D2H of every slot before returning, so the pinned outputs may
        be read right away; sync=False only orders the current stream after the copies (the caller synchronises).
        `outputs[i]` needs the keys of self.outputs()."""
        if self._plan is None:
            raise L.GaitLibraryError("call plan()/capture() first")
        ns = len(self._slots)
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_io_streams"):
            self._io_streams = (torch.cuda.Stream(), torch.cuda.Stream())
        s_in, s_out = self._io_streams
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        ev_in = [None] * ns          # H2D of the slot's input finished
        ev_cmp = [None] * ns         # the slot's kernels finished
        ev_out = [None] * ns         # D2H of the slot's outputs finished
        for i, (x_host, out_host) in enumerate(zip(inputs, outputs)):
            k = i % ns
            p = self._slots[k]
            with torch.cuda.stream(s_in):
                if ev_cmp[k] is not None:
                    s_in.wait_event(ev_cmp[k])               # previous batch in this slot consumed its input
                p["x"].copy_(x_host, non_blocking=True)
                ev_in[k] = s_in.record_event()
            cur.wait_event(ev_in[k])
            if ev_out[k] is not None:
                cur.wait_event(ev_out[k])                    # previous outputs of this slot are on the host
            if self._graphs[k] is not None:
                self._graphs[k].replay()
            else:
                self._launch(p)
            ev_cmp[k] = cur.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                if isinstance(out_host, HostOutputs):
                    out_host.packed.copy_(p["outbuf"], non_blocking=True)
                else:
                    for key, v in self.outputs(k).items():
                        out_host[key].copy_(v, non_blocking=True)
                ev_out[k] = s_out.record_event()
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)
        if sync:
            for ev in ev_out:
                if ev is not None:
                    ev.synchronize()             # the host may read outputs[...] as soon as this returns

    def outputs(self, slot: int = 0):
        p = self._slots[slot]
        S, T = p["S"], p["T"]
        v = lambda t: t.view(S, T, *t.shape[1:])
        out = {"theta": v(p["theta"]), "kp_2d": v(p["kp2d"]), "kp_3d": v(p["joints"]), "rotmat": v(p["rotmat"]),
               "kinect25": v(p["kinect"])}
        if p["verts"] is not None:               # absent in joints-only mode and when the mesh goes to an external address
            out["verts"] = v(p["verts"])
        return out

    @torch.no_grad()
    def step(self):
        """Run one step on the planned input buffer `self.input` (graph replay if captured)."""
        if self._graph is not None:
            self._graph.replay()
        else:
            self._launch(self._plan)

    @property
    def input(self) -> torch.Tensor:
        return self._plan["x"]

    @torch.no_grad()
    def forward(self, features: torch.Tensor):
        """features (S,T,2048) FP32 CUDA -> dict of (S,T,...) tensors (views of the planned buffers:
        valid until the next call)."""
        features = L.f32(features, "features")
        if features.dim() != 3:
            raise ValueError(f"features must be (S,T,{self.encoder.input_size}), got {tuple(features.shape)}")
        S, T, _ = features.shape
        if self._plan is None or (self._plan["S"], self._plan["T"]) != (S, T):
            self.plan(S, T)
        self._plan["x"].copy_(features)
        self.step()
        return self.outputs()
