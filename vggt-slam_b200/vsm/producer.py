"""Producer hand-off: VGGT predictions -> Submap without the host round trip.

The reference copies every prediction tensor to host numpy (vggt_slam/solver.py:478-480), prepares the point map,
colours and poses with numpy in ``Solver.add_points`` (solver.py:249-263, 301, 337-340) and stores numpy arrays in the
Submap; fusion then copies them back to the GPU.  Here the prediction dict may hold CUDA tensors (or numpy arrays):
the same preparation runs as libvsm kernels and the Submap keeps device tensors, which the fuse calls use in place.

``FrameStream`` fuses one frame at a time as the producer emits them (BASELINE configs[4]).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _native as N
from .voxel_map import DeviceVoxelMap, _ptr, _stream_ptr, as_device, require_cuda


def closed_form_inverse_se3(extrinsic) -> np.ndarray:
    """(S,3,4) or (S,4,4) world-to-camera -> (S,4,4) float64 camera-to-world; R^T and -R^T t are formed in the input's
    dtype (vggt.utils.geometry.closed_form_inverse_se3, used at solver.py:263).  S x 12 numbers: host arithmetic."""
    e = extrinsic.detach().cpu().numpy() if isinstance(extrinsic, torch.Tensor) else np.asarray(extrinsic)
    R, T = e[:, :3, :3], e[:, :3, 3:]
    Rt = np.transpose(R, (0, 2, 1))
    out = np.tile(np.eye(4), (len(R), 1, 1))
    out[:, :3, :3] = Rt
    out[:, :3, 3:] = -np.matmul(Rt, T)
    return out


def unproject_depth_map_to_point_map(depth_map, extrinsics_cam, intrinsics_cam, out_f64: bool = False,
                                     device: Optional[torch.device] = None) -> torch.Tensor:
    """depth (S,H,W,1) or (S,H,W), extrinsic (S,3,4), intrinsic (S,3,3) -> world points (S,H,W,3) on the device
    (vsm_unproject_depth; float32 by default -- what the fuse calls read -- or the reference's float64)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    d = as_device(depth_map, dev, torch.float32)
    if d.ndim == 4:
        d = d[..., 0].contiguous()
    S, H, W = (int(v) for v in d.shape)
    c2w = torch.from_numpy(np.ascontiguousarray(closed_form_inverse_se3(extrinsics_cam)[:, :3, :], dtype=np.float64)).to(dev)
    K = as_device(intrinsics_cam, dev, torch.float32).reshape(S, 9)
    out = torch.empty((S, H, W, 3), dtype=torch.float64 if out_f64 else torch.float32, device=dev)
    N.check(N.lib.vsm_unproject_depth(_ptr(d), _ptr(c2w), _ptr(K), S, H, W, _ptr(out), int(out_f64), _stream_ptr(dev)))
    return out


def images_to_colors(images, device: Optional[torch.device] = None) -> torch.Tensor:
    """(S,3,H,W) float32 in [0,1] -> (S,H,W,3) uint8 on the device (solver.py:260)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    img = as_device(images, dev, torch.float32)
    S, _, H, W = (int(v) for v in img.shape)
    out = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev)
    N.check(N.lib.vsm_images_to_colors(_ptr(img), S, H, W, _ptr(out), _stream_ptr(dev)))
    return out


def scale_points_(points: torch.Tensor, scale: float) -> torch.Tensor:
    """In-place ``world_points *= scale_factor`` on a float32 CUDA tensor (solver.py:301)."""
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
    N.check(N.lib.vsm_scale_points(_ptr(points), int(points.numel()), float(scale), _stream_ptr(points.device)))
    return points


def add_points_to_submap(submap, pred_dict: dict, conf_threshold_percentile: float, use_point_map: bool = True,
                         scale_factor: Optional[float] = None, H_world_map=None):
    """The data path of ``Solver.add_points`` (solver.py:249-263, 301, 337-340) for one submap: pick the point map
    (``world_points`` / ``world_points_conf``) or unproject ``depth`` / ``depth_conf``, derive the colours and the
    camera-to-world poses, optionally apply the Sim(3) scale factor, and hand everything to the Submap as DEVICE tensors.
    The relative-transform estimation of add_points (RANSAC / scale alignment against the previous submap, pose graph)
    is the caller's: pass its result as ``H_world_map``.  Returns the device point map."""
    extr, intr = pred_dict["extrinsic"], pred_dict["intrinsic"]
    if use_point_map:
        pts, conf = pred_dict["world_points"], pred_dict["world_points_conf"]
        dev = torch.device("cuda", torch.cuda.current_device())
        pts = as_device(pts, dev, torch.float32)
        if scale_factor is not None and pts.data_ptr() == getattr(pred_dict["world_points"], "data_ptr", lambda: 0)():
            pts = pts.clone()  # the reference scales its own host copy, never the caller's tensor
    else:
        pts, conf = unproject_depth_map_to_point_map(pred_dict["depth"], extr, intr), pred_dict["depth_conf"]
    conf = as_device(conf, pts.device, torch.float32)
    colors = images_to_colors(pred_dict["images"], pts.device)
    cam_to_world = closed_form_inverse_se3(extr)
    if scale_factor is not None:
        scale_points_(pts, float(scale_factor))
        cam_to_world[:, 0:3, 3] *= scale_factor
    if H_world_map is not None:
        submap.set_reference_homography(np.asarray(H_world_map, dtype=np.float64))
    submap.add_all_poses(cam_to_world)
    submap.add_all_points(pts, colors, conf, conf_threshold_percentile, intr)
    submap.set_conf_masks(conf)
    return pts


class FrameStream:
    """Fuses a submap frame by frame as the producer emits frames: each call is one vsm_fuse_submap of a single frame
    with the frame's index inside the submap as ``frame_base`` (so contributor masks name the right frame).  The
    confidence threshold of a streamed submap cannot be the percentile over the whole submap (it is not complete yet):
    the caller supplies it (e.g. the previous submap's, solver.py:287)."""

    def __init__(self, device_map: DeviceVoxelMap, submap_id: int, H_world_map, conf_threshold: float, flags: int = 0):
        self.dm, self.submap_id, self.flags = device_map, int(submap_id), int(flags)
        self.H_world_map = np.asarray(H_world_map, dtype=np.float64)
        self.conf_threshold = float(conf_threshold)
        self.frames = 0

    def push(self, points: torch.Tensor, conf: torch.Tensor, emb: torch.Tensor) -> dict:
        """points (H,W,3) f32, conf (H,W) f32, emb (H,W,d) -- CUDA tensors of ONE frame."""
        H, W = int(points.shape[0]), int(points.shape[1])
        p = self.dm.make_params(1, H, W, 1, 1, self.conf_threshold, self.H_world_map, self.submap_id, self.flags,
                                frame_base=self.frames)
        st = self.dm.fuse(points[None].contiguous(), conf[None].contiguous(), emb[None].contiguous(), p)
        self.frames += 1
        return st
