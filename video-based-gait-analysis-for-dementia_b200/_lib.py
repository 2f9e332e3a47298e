"""ctypes binding of the C-ABI CUDA library (include/gaitb200.h).

The library is built in-tree (``make`` / ``__graft_entry__.build()``) as
``lib/libgaitb200.so``.  There is no other implementation behind this module: if the
library is missing, or a tensor is not an FP32 CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("GAITB200_LIB", _HERE / "lib" / "libgaitb200.so")).resolve()

P, I32, I64, F32, F64, SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGS = {
    "gait_abi_version": [],
    "gait_error_string": [I32],
    "gait_last_error": [],
    "gait_device_info": [C.POINTER(I32), C.POINTER(I32), C.POINTER(I32)],
    "gait_launch_count": [],
    "gait_rot6d_to_rotmat": [P, I32, I64, P, I64, F32, P],
    "gait_rotmat_to_rot6d": [P, P, I64, P],
    "gait_rotmat_to_quaternion": [P, I32, P, I64, F32, P],
    "gait_quaternion_to_axis_angle": [P, P, I64, P],
    "gait_rotmat_to_axis_angle": [P, I32, P, I64, I32, I64, I32, P],
    "gait_quat2mat": [P, P, I64, P],
    "gait_batch_rodrigues": [P, P, I64, I32, P],
    "gait_weak_perspective_to_translation": [P, P, I64, F32, F32, P],
    "gait_perspective_projection": [P, P, P, P, F32, F32, P, I64, I32, P],
    "gait_linear": [P, I64, P, I64, P, P, I64, P, I64, I64, I64, I64, P],
    "gait_prepare_weight": [P, P, I64, P],
    "gait_release_weight": [P],
    "gait_split_weight": [P, P, I64, P],
    "gait_linear_prepared": [P, I64, P, I64, P, I64, P, P, P, I64, P, I64, I64, I64, I64, P],
    "gait_debug_linear_trace": [P],
    "gait_debug_gru_trace": [P],
    "gait_debug_pdl_mask": [I32],
    "gait_gru_workspace_bytes": [I64, I64, I64],
    "gait_gru_plan": [I64, I64, I64],
    "gait_gru_layer": [P, I64, P, P, P, P, P, P, I64, P, I64, P, I64, P, I64, I64, I64, I64, I32, P, SZ, P],
    "gait_relu": [P, P, I64, P],
    "gait_hmr_workspace_bytes": [I64, I64],
    "gait_hmr_regressor": [P, I64, P, P, P, P, P, P, P, P, I64, I32, P, I64, I64, I64, P, SZ, P],
    "gait_hmr_folded_workspace_bytes": [I64],
    "gait_hmr_regressor_folded": [P, I64, P, P, P, I64, I64, P, SZ, P],
    "gait_smpl_pose_chain": [P, P, I64, P, P, P, P, P, P, P, I64, P],
    "gait_smpl_pose_chain_rot6d": [P, I64, F32, P, I64, P, I64, P, P, P, P, P, P, P, P, P, I64, P],
    "gait_smpl_blend": [P, P, P, I64, I64, I64, P],
    "gait_smpl_lbs": [P, I64, P, P, P, I64, I64, P],
    "gait_smpl_lbs_jx_parts": [I64],
    "gait_smpl_lbs_pack_bytes": [I64],
    "gait_smpl_lbs_pack": [P, P, I64, P],
    "gait_smpl_lbs_aop_bytes": [I64],
    "gait_smpl_lbs_tc": [P, I64, P, P, P, P, P, I64, I64, P],
    "gait_smpl_lbs_tc_joints": [P, I64, P, P, P, P, P, I32, P, I64, I64, P],
    "gait_smpl_lbs_tc_ex": [P, I64, P, P, P, P, P, P, I32, P, I64, I64, P],
    "gait_peer_alloc": [C.POINTER(P), SZ],
    "gait_peer_free": [P],
    "gait_peer_export": [P, C.c_char_p],
    "gait_peer_open": [C.c_char_p, C.POINTER(P)],
    "gait_peer_close": [P],
    "gait_peer_copy": [P, P, SZ, P],
    "gait_smpl_reduced_joints": [P, P, I64, P, P, P, P, I64, I32, P],
    "gait_joint_regress": [P, P, P, I64, I64, I32, P],
    "gait_joint_regress_pack_bytes": [I64, I32],
    "gait_joint_regress_pack": [P, P, I64, I32, P],
    "gait_joint_regress_packed": [P, P, P, I64, I64, I32, P],
    "gait_joints_assemble": [P, P, I64, P, I32, P, I32, I32, I64, P, I32, P, P, I64, F32, F32, F32, P, P, I32, P, I64, P],
    "gait_gather_joints": [P, I32, P, I32, P, I64, P],
    "gait_one_euro_filter": [P, P, I64, I64, F64, F64, F64, P],
    "gait_crop_cam_to_orig_img": [P, P, I32, I64, F64, F64, P, I64, P],
    "gait_crop_coords_to_orig_img": [P, I32, I64, P, P, I64, I32, I32, F64, P],
    "gait_keypoint_attention": [P, P, F32, P, I64, I32, I32, I32, I64, I64, I64, P],
    "gait_locally_connected": [P, I64, I64, I64, P, I64, I64, I64, P, I64, I64, P, I64, I64, I64, P, P, I64, I32, I32, I32, P],
    "gait_activation": [P, P, I64, I32, F32, P],
    "gait_pack_theta": [P, P, I64, P, I64, P, I64, P],
}
_RESTYPES = {
    "gait_error_string": C.c_char_p,
    "gait_last_error": C.c_char_p,
    "gait_launch_count": I64,
    "gait_gru_workspace_bytes": SZ,
    "gait_hmr_workspace_bytes": SZ,
    "gait_hmr_folded_workspace_bytes": SZ,
    "gait_smpl_lbs_jx_parts": I64,
    "gait_smpl_lbs_pack_bytes": SZ,
    "gait_smpl_lbs_aop_bytes": SZ,
    "gait_joint_regress_pack_bytes": SZ,
}
EXPORTS = tuple(_SIGS)

_lib = None


class GaitLibraryError(RuntimeError):
    pass


def load():
    """Load lib/libgaitb200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GaitLibraryError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "gaitb200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, I32)
    if lib.gait_abi_version() != 1:
        raise GaitLibraryError(f"ABI version mismatch: library reports {lib.gait_abi_version()}")
    _lib = lib
    return lib


def call(name: str, *args):
    """Call an int-returning entry point; raise on a negative return code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        detail = lib.gait_last_error().decode(errors="replace")
        kind = lib.gait_error_string(rc).decode()
        raise GaitLibraryError(f"{name} failed ({rc}: {kind}): {detail}")


_prepared = {}      # data_ptr of a registered weight -> (weakref to its owner, [hi|lo] tensor, version)
_owner_key = {}     # id(owner tensor) -> data_ptr it was registered under (an owner whose storage moved drops its old entry)


def _drop_prepared(key):
    ent = _prepared.pop(key, None)
    if ent is not None:
        for oid, k in list(_owner_key.items()):
            if k == key:
                _owner_key.pop(oid, None)
        if _lib is not None:
            try:
                _lib.gait_release_weight(key)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass


def purge_prepared():
    """Drop every registered weight whose owner is gone or no longer lives at the registered address (a Parameter keeps its
    identity through module.to(device) / `.data = ...` / load_state_dict(assign=True) while its storage changes, and the
    caching allocator may hand the old range to another tensor)."""
    for key, (ref, _lo, _ver) in list(_prepared.items()):
        w = ref()
        if w is None or w.data_ptr() != key:
            _drop_prepared(key)


def prepare_weight(w: torch.Tensor) -> torch.Tensor:
    """Register the constant FP32 CUDA weight `w` (contiguous; a Parameter, buffer or packed tensor that its owner keeps
    alive) with the library: its TF32 hi/lo split is computed once and the GEMMs that use `w` (or a view into it) load it by
    TMA instead of recomputing it per call.  Idempotent; re-prepared when `w` was modified in place; released when `w` is
    garbage-collected, when it is registered again after its storage moved, or by purge_prepared() (modules call it from
    _apply, i.e. on .to() / .cuda() / .float())."""
    import weakref
    if not (torch.is_tensor(w) and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
        raise GaitLibraryError("prepare_weight: expected a contiguous float32 CUDA tensor")
    key = w.data_ptr()
    old = _owner_key.get(id(w))
    if old is not None and old != key:
        _drop_prepared(old)                       # same tensor object, new storage: the old range is no longer ours
    hit = _prepared.get(key)
    if hit is not None and hit[0]() is w and hit[2] == w._version:
        return w
    if hit is not None:
        _drop_prepared(key)                       # another (dead or moved) owner had this address, or w changed in place
    if w.numel() % 4:
        return w                                  # not TMA-addressable anyway
    with torch.cuda.device(w.device):
        lo = torch.empty((2,) + tuple(w.shape), device=w.device, dtype=torch.float32)    # [hi | lo]
        call("gait_prepare_weight", key, ptr(lo), w.numel(), stream_ptr())
    oid = id(w)
    _prepared[key] = (weakref.ref(w, lambda _r, k=key, o=oid: (_owner_key.pop(o, None), _drop_prepared(k))), lo, w._version)
    _owner_key[oid] = key
    return w


def release_weight(w: torch.Tensor):
    _drop_prepared(w.data_ptr())


_packed_jreg = {}      # id(tensor) -> (weakref to the regressor tensor, its _version, data_ptr, packed tensor); at most 8 entries


def joint_regress(verts: torch.Tensor, Jr: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """out (F,Rj,3) = Jr (Rj,V) . verts (F,V,3)  (pare.py:70-76, spin.py:279-282, smpl.py:113).  Batches of >= 8 frames take the
    streaming cluster kernel (jreg.cu) with the regressor packed once per tensor (re-packed when it was modified in place or
    replaced); tiny batches (model preparation) take the generic kernel."""
    import weakref
    F, V, Rj = verts.shape[0], verts.shape[1], Jr.shape[0]
    if out is None:
        out = torch.empty(F, Rj, 3, device=verts.device, dtype=torch.float32)
    if F < 8 or V % 2 or verts.data_ptr() % 8:
        call("gait_joint_regress", ptr(verts), ptr(Jr), ptr(out), F, V, Rj, stream_ptr())
        return out
    key = id(Jr)
    hit = _packed_jreg.get(key)
    if hit is None or hit[0]() is not Jr or hit[1] != Jr._version or hit[2] != Jr.data_ptr():
        packed = torch.empty(load().gait_joint_regress_pack_bytes(V, Rj) // 4, device=Jr.device, dtype=torch.float32)
        call("gait_joint_regress_pack", ptr(Jr), ptr(packed), V, Rj, stream_ptr())
        if len(_packed_jreg) >= 8:
            _packed_jreg.pop(next(iter(_packed_jreg)))
        _packed_jreg[key] = hit = (weakref.ref(Jr, lambda _r, k=key: _packed_jreg.pop(k, None)), Jr._version, Jr.data_ptr(), packed)
    call("gait_joint_regress_packed", ptr(verts), ptr(hit[3]), ptr(out), F, V, Rj, stream_ptr())
    return out


def launch_count() -> int:
    return int(load().gait_launch_count())


def require_device():
    """Fail loudly unless a CUDA sm_100 device is current."""
    if not torch.cuda.is_available():
        raise GaitLibraryError("gaitb200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    sm, major, minor = I32(), I32(), I32()
    call("gait_device_info", C.byref(sm), C.byref(major), C.byref(minor))
    return sm.value, major.value, minor.value


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def f32(t, name: str = "tensor") -> torch.Tensor:
    """Validate an input: FP32 CUDA tensor, made contiguous."""
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise GaitLibraryError(f"{name}: expected a CUDA tensor (gaitb200 has no CPU path), got device {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
