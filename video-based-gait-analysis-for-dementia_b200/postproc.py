"""Post-processing that follows / re-uses the regression head, on the GPU:

  * one_euro_filter / smooth_pose          lib/utils/one_euro_filter.py:5-46, lib/utils/smooth_pose.py:28-116
  * convert_crop_cam_to_orig_img           lib/utils/demo_utils.py:176-193
  * convert_crop_coords_to_orig_img        lib/utils/demo_utils.py:196-209

Same names, arguments and return conventions as the reference (numpy in -> numpy out; CUDA tensors are accepted too and
then returned as tensors).  The arithmetic runs in the sm_100a kernels of csrc/postproc.cu, which reproduce numpy's
float32 / float64 operation sequence exactly; there is no CPU implementation here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import geometry as G
from .kp_utils import convert_kps


def _to_cuda(a, name, dtypes=(torch.float32,)):
    was_np = isinstance(a, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(a)) if was_np else a
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected numpy.ndarray or torch.Tensor, got {type(a)}")
    if t.dtype not in dtypes:
        raise TypeError(f"{name}: expected dtype in {dtypes}, got {t.dtype}")
    L.require_device()
    return t.cuda().contiguous(), was_np


def one_euro_filter(x, min_cutoff=1.0, beta=0.0, d_cutoff=1.0):
    """OneEuroFilter(t0=zeros, x0=x[0], min_cutoff, beta, d_cutoff) applied to x[1], x[2], ... at t = 1, 2, ...
    (smooth_pose.py:51-56,84-88).  x (T, ...) float32 -> filtered (T, ...), x_hat[0] = x[0]."""
    xd, was_np = _to_cuda(x, "x")
    T = xd.shape[0]
    out = torch.empty_like(xd)
    L.call("gait_one_euro_filter", L.ptr(xd), L.ptr(out), T, xd.numel() // max(T, 1), float(min_cutoff), float(beta),
           float(d_cutoff), L.stream_ptr())
    return out.cpu().numpy() if was_np else out


def smooth_pose(pred_pose, pred_betas, min_cutoff=0.004, beta=0.7, device='cuda', kinectv2=False, smpl=None):
    """lib/utils/smooth_pose.py:28-116.  pred_pose (T,72) axis-angle or (T,96) quaternions, pred_betas (T,10) ->
    (verts (T,6890,3), pred_pose_hat, joints3d): One-Euro filtered poses, then SMPL(pose2rot=True) on every frame with
    the first frame's betas - here ONE filter launch and ONE batched SMPL pass instead of T batch-1 passes.
    `smpl`: a gaitb200.smpl.SMPL to use (default: SMPL(SMPL_MODEL_DIR), as the reference constructs it)."""
    from .smpl import SMPL, SMPL_MODEL_DIR
    pose, _ = _to_cuda(pred_pose, "pred_pose")
    betas, _ = _to_cuda(pred_betas, "pred_betas")
    T = betas.shape[0]
    if pose.shape[-1] == 72:
        q = 3
    elif pose.shape[-1] == 96:
        q = 4
    else:
        raise ValueError(f"Invalid pred_pose format: {tuple(pose.shape)}")
    pshape = pose.shape
    hat = one_euro_filter(pose.reshape(T, 24 * q), min_cutoff=min_cutoff, beta=beta)
    if smpl is None:
        smpl = SMPL(model_path=SMPL_MODEL_DIR)
    smpl = smpl.to(pose.device)
    smpl.kinectv2 = kinectv2
    aa = hat.reshape(T, 24, 3) if q == 3 else G.quaternion_to_angle_axis(hat.reshape(T * 24, 4)).reshape(T, 24, 3)
    out = smpl(betas=betas[0:1].expand(T, -1).contiguous(), body_pose=aa[:, 1:].reshape(T, 69), global_orient=aa[:, 0],
               pose2rot=True)
    joints = out.joints
    if kinectv2:
        joints = convert_kps(joints, 'spin2', 'kinectv2').cpu().numpy().astype(np.float64)   # kp_utils.py:30 returns float64
    else:
        joints = joints.cpu().numpy()
    return out.vertices.cpu().numpy(), hat.reshape(pshape).cpu().numpy(), joints


def _bbox(bbox):
    b, _ = _to_cuda(bbox, "bbox", (torch.float32, torch.float64))
    if b.dim() != 2 or b.shape[1] < 3:
        raise ValueError(f"bbox must be (N, >=3) [c_x, c_y, h, ...], got {tuple(b.shape)}")
    return b


def convert_crop_cam_to_orig_img(cam, bbox, img_width, img_height):
    """demo_utils.py:176-193.  cam (N,3) float32, bbox (N,>=3) float32/float64 -> (N,4) [sx, sy, tx, ty] in bbox's dtype."""
    c, was_np = _to_cuda(cam, "cam")
    b = _bbox(bbox)
    out = torch.empty(c.shape[0], 4, device=c.device, dtype=b.dtype)
    L.call("gait_crop_cam_to_orig_img", L.ptr(c), L.ptr(b), int(b.dtype == torch.float64), b.stride(0), float(img_width),
           float(img_height), L.ptr(out), c.shape[0], L.stream_ptr())
    return out.cpu().numpy() if was_np else out


def convert_crop_coords_to_orig_img(bbox, keypoints, crop_size):
    """demo_utils.py:196-209.  keypoints (N,J,D>=2) float32 in [-1,1] crop units -> original-image pixels (float32)."""
    k, was_np = _to_cuda(keypoints, "keypoints")
    b = _bbox(bbox)
    if k.dim() != 3 or k.shape[2] < 2 or k.shape[0] != b.shape[0]:
        raise ValueError(f"keypoints must be (N,J,>=2) with N = {b.shape[0]}, got {tuple(k.shape)}")
    out = torch.empty_like(k)
    L.call("gait_crop_coords_to_orig_img", L.ptr(b), int(b.dtype == torch.float64), b.stride(0), L.ptr(k), L.ptr(out),
           k.shape[0], k.shape[1], k.shape[2], float(crop_size), L.stream_ptr())
    return out.cpu().numpy() if was_np else out


class KinectDbWriter:
    """The Kinect-25 database batch_generation.py writes (batch_generation.py:226-238,262-283; doc/batch_generation.md:6-10):
    a joblib dump of {'vid_name': (N,) str, 'bbox': (N,4) float32, 'joints3D': (N,25,3) float32}, one entry per frame in
    temporal order, split into `<stem>_<k>.json` shards every `max_videos` entries of the video list (the reference's
    MAX_VID; skipped videos count, `skip()`) with the remainder in the last shard.  `add` accepts the head's `kinect25` output as a CUDA tensor (one device-to-host copy per
    video) or a numpy array; this is host-side data-format code, no arithmetic."""

    def __init__(self, outpath: str, max_videos: int = 300, min_tail: int = 10, total: int | None = None):
        """total: len(vidnames) of the reference's loop, INCLUDING videos that will be skipped; the reference cuts a shard at
        the top of iteration idx when idx % MAX_VID == 0, idx > 0 and len(vidnames) - idx > 10 (batch_generation.py:226)."""
        if not outpath.endswith(".json"):
            raise ValueError("outpath must end with .json (batch_generation.py:235)")
        self.outpath, self.max_videos, self.min_tail = outpath, int(max_videos), int(min_tail)
        self.total = None if total is None else int(total)
        self._db = {"vid_name": [], "bbox": [], "joints3D": []}
        self._idx = 0                                 # the reference's enumerate index: counts skipped videos too
        self.files = []

    def _top_of_iteration(self, videos_left):
        """The reference's shard cut, evaluated before video self._idx is processed or skipped."""
        idx = self._idx
        if videos_left is not None:
            remaining = videos_left + 1               # len(vidnames) - idx
        elif self.total is not None:
            remaining = self.total - idx
        else:
            remaining = None                          # unknown length: always cut
        if idx > 0 and idx % self.max_videos == 0 and (remaining is None or remaining > self.min_tail):
            self._flush()
        self._idx += 1

    def skip(self, videos_left: int | None = None):
        """A video without annotations (batch_generation.py:246-248 `continue`): nothing is stored, but it advances the
        enumerate index the shard rule is evaluated on, exactly as in the reference."""
        self._top_of_iteration(videos_left)

    def add(self, vid_name: str, joints3d, bbox, videos_left: int | None = None):
        """One video: joints3d (frames,25,3) [any shape with frames*75 elements], bbox (frames,4).  `videos_left`: how many
        entries of the video list follow this one (alternative to `total`; the reference only cuts a shard when more than
        `min_tail` remain, counting the current one)."""
        if torch.is_tensor(joints3d):
            joints3d = joints3d.detach().to("cpu", torch.float32).numpy()
        bbox = np.asarray(bbox.detach().cpu().numpy() if torch.is_tensor(bbox) else bbox)
        n = bbox.reshape(-1, 4).shape[0]
        j = np.asarray(joints3d).reshape(n, 25, 3)
        self._top_of_iteration(videos_left)
        self._db["vid_name"].extend([vid_name.split(".")[0]] * n)
        self._db["bbox"].append(bbox.reshape(n, 4))
        self._db["joints3D"].append(j)

    def _flush(self):
        import joblib
        if not self._db["vid_name"]:
            return None
        db = {"vid_name": np.array(self._db["vid_name"]),
              "bbox": np.concatenate(self._db["bbox"], axis=0).astype(np.float32),
              "joints3D": np.concatenate(self._db["joints3D"], axis=0).astype(np.float32)}
        path = self.outpath[:-5] + f"_{len(self.files)}.json"
        joblib.dump(db, path)
        self.files.append(path)
        self._db = {"vid_name": [], "bbox": [], "joints3D": []}
        return path

    def close(self):
        """Write the remaining frames; returns the list of files written."""
        self._flush()
        return self.files
