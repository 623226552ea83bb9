"""Occupancy grid of a point cloud with the reference's API (get_occupancy.py:130-210), built on the GPU
(vsm_occupancy_build, csrc/occupancy.cu): the same hash + atomic min/max + sort of the distinct cells as the voxel map.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .voxel_map import _ptr, _stream_ptr, as_device, require_cuda


def build_occupancy_from_pointcloud(points_xyz, voxel_size: float, ceiling_z: float, height_thresh: float,
                                    device: Optional[torch.device] = None
                                    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """centers (M,3) float32 (cube centres at minz + voxel/2), is_blocked (M,) bool (height range > height_thresh),
    cell_keys (M,2) int64, minz (M,) float32 -- cells in np.unique(axis=0) order, like get_occupancy.py:130-179.
    `points_xyz` may be a numpy array or a CUDA tensor."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    pts = as_device(points_xyz, dev, torch.float32).reshape(-1, 3)
    n = int(pts.shape[0])
    n_cells, n_kept = C.c_int64(0), C.c_int64(0)
    args = (float(voxel_size), float(ceiling_z), float(height_thresh))
    N.check(N.lib.vsm_occupancy_build(_ptr(pts), n, *args, 0, None, None, None, None, C.byref(n_cells), C.byref(n_kept),
                                      _stream_ptr(dev)))
    m = int(n_cells.value)
    centers = torch.empty((m, 3), dtype=torch.float32, device=dev)
    blocked = torch.empty((m,), dtype=torch.uint8, device=dev)
    keys = torch.empty((m, 2), dtype=torch.int64, device=dev)
    minz = torch.empty((m,), dtype=torch.float32, device=dev)
    if m:
        N.check(N.lib.vsm_occupancy_build(_ptr(pts), n, *args, m, _ptr(centers), _ptr(blocked), _ptr(keys), _ptr(minz),
                                          C.byref(n_cells), C.byref(n_kept), _stream_ptr(dev)))
    return centers.cpu().numpy(), blocked.cpu().numpy().astype(bool), keys.cpu().numpy(), minz.cpu().numpy()


def segment_is_navigable(p0, p1, voxel_size: float, blocked_cells: Dict[Tuple[int, int], bool],
                         unknown_is_free: bool = True) -> bool:
    """Straight-line navigability in XY by sampling occupancy cells every half voxel (get_occupancy.py:182-210)."""
    p0 = np.asarray(p0, dtype=np.float32).reshape(3)
    p1 = np.asarray(p1, dtype=np.float32).reshape(3)
    d = float(np.linalg.norm(p1[:2] - p0[:2]))
    n = max(2, int(np.ceil(d / (float(voxel_size) * 0.5))) + 1)
    ts = np.linspace(0.0, 1.0, n, dtype=np.float32)
    xs = p0[0] + (p1[0] - p0[0]) * ts
    ys = p0[1] + (p1[1] - p0[1]) * ts
    for x, y in zip(xs, ys):
        key = (int(np.floor(x / voxel_size)), int(np.floor(y / voxel_size)))
        if key not in blocked_cells:
            if unknown_is_free:
                continue
            return False
        if blocked_cells[key]:
            return False
    return True
