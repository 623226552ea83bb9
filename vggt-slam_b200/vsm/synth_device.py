"""Device-side synthetic submaps for the benchmark: the same "box room + trajectory"
geometry, 1 + Gamma(2,2) confidence and bf16-exact N(0,1) embeddings as ``vsm.synth``, generated with
torch on the GPU so that config-2-sized inputs (20 submaps x 5 GB of embeddings) never touch the host.
Data generation only -- nothing here is on the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import synth


@dataclass
class DeviceSubmapData:
    submap_id: int
    points: torch.Tensor  # (S,H,W,3) f32 cuda
    conf: torch.Tensor  # (S,H,W) f32 cuda
    emb: torch.Tensor  # (S,H,W,d) bf16 or f32 cuda
    H_world_map: np.ndarray  # (4,4) f64
    frame_paths: List[str]
    last_non_loop_frame_index: int
    conf_percentile: float = 25.0


def _rot(axis, angle: float) -> np.ndarray:
    return synth._rot(np.asarray(axis, dtype=np.float64), angle)


def box_room_frames_device(gen: torch.Generator, S: int, H: int, W: int, room, noise: float, start: float,
                           device: torch.device, step: float = 0.08):
    room_t = torch.tensor(room, dtype=torch.float64, device=device)
    fx = 0.8 * W
    u = (torch.arange(W, dtype=torch.float64, device=device) - (W - 1) / 2.0) / fx
    v = (torch.arange(H, dtype=torch.float64, device=device) - (H - 1) / 2.0) / fx
    dirs = torch.stack(torch.broadcast_tensors(u[None, :], v[:, None], torch.ones((H, W), dtype=torch.float64,
                                                                               device=device)), dim=-1)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    pts = torch.empty((S, H, W, 3), dtype=torch.float64, device=device)
    cams = []
    room_np = np.asarray(room, dtype=np.float64)
    for s in range(S):
        t = start + step * s
        centre = room_np * (0.5 + 0.22 * np.array([math.cos(t), math.sin(1.3 * t), 0.3 * math.sin(0.7 * t)]))
        R = _rot([0.0, 0.0, 1.0], 0.9 * t + 0.3) @ _rot([1.0, 0.0, 0.0], -math.pi / 2 + 0.15 * math.sin(t))
        cams.append((R, centre))
        Rt = torch.tensor(R, dtype=torch.float64, device=device)
        c = torch.tensor(centre, dtype=torch.float64, device=device)
        d = dirs @ Rt.T
        t_hi = (room_t - c) / d
        t_lo = (0.0 - c) / d
        tt = torch.where(d > 0, t_hi, t_lo)
        tt = torch.where(torch.isfinite(tt) & (tt > 0), tt, torch.full_like(tt, float("inf")))
        depth = tt.min(dim=-1).values
        depth = depth + noise * torch.randn(depth.shape, dtype=torch.float64, device=device, generator=gen)
        pts[s] = c + d * depth[..., None]
    R0, c0 = cams[0]
    local = (pts - torch.tensor(c0, dtype=torch.float64, device=device)) @ torch.tensor(R0, dtype=torch.float64,
                                                                                          device=device)
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R0, c0
    return local.to(torch.float32).contiguous(), M


def make_submap_device(seed: int, submap_id: int, S: int = 32, H: int = 294, W: int = 518, d: int = 512,
                       mode: str = "sl4", room=(12.0, 8.0, 3.0), start: Optional[float] = None, noise: float = 0.005,
                       emb_dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None,
                       first_frame_number: int = 0, with_emb: bool = True) -> DeviceSubmapData:
    """with_emb=False leaves ``emb`` None (callers that attach indexed embeddings: make_indexed_device)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + submap_id)
    if start is None:
        start = 0.08 * S * submap_id
    pts, M_room_local = box_room_frames_device(gen, S, H, W, room, noise, start, device)
    # 1 + Gamma(2, 2): sum of two exponentials with scale 2
    u = torch.rand((2, S, H, W), dtype=torch.float32, device=device, generator=gen).clamp_min_(1e-12)
    conf = (1.0 - 2.0 * (torch.log(u[0]) + torch.log(u[1]))).contiguous()
    del u
    emb = None
    if with_emb:
        emb = torch.empty((S, H, W, d), dtype=emb_dtype, device=device)
        for s in range(S):  # frame by frame: keeps the float32 temporary small
            emb[s] = torch.randn((H, W, d), dtype=torch.float32, device=device, generator=gen).to(torch.bfloat16).to(
                emb_dtype)
    rng = np.random.default_rng([seed, submap_id])
    grng = np.random.default_rng([seed, 987654321])
    G = synth.random_sl4(grng) if mode == "sl4" else synth.random_sim3(grng, scale=None if mode == "se3" else 1.7)
    Hm = G @ M_room_local
    if mode == "sl4":
        Hm = Hm @ (np.eye(4) + 1e-3 * rng.normal(size=(4, 4)))
        Hm = Hm / abs(np.linalg.det(Hm)) ** 0.25
    elif mode == "se3":
        Hm[3, :] = [0.0, 0.0, 0.0, 1.0]
    paths = [f"left_{first_frame_number + i:06d}.png" for i in range(S)]
    return DeviceSubmapData(submap_id, pts, conf, emb, Hm.astype(np.float64), paths, S - 1)


def make_trajectory_submap_device(seed: int, submap_id: int, S: int = 32, H: int = 294, W: int = 518, d: int = 512,
                                   room=(8.0, 6.0, 3.0), step: float = 0.2, noise: float = 0.005,
                                   emb_dtype: torch.dtype = torch.bfloat16,
                                   device: Optional[torch.device] = None, with_emb: bool = True,
                                   emb_from: Optional[torch.Tensor] = None, projective: float = 1e-5) -> DeviceSubmapData:
    """Submap `submap_id` of the LONG-TRAJECTORY workload (BASELINE configs[2]): a corridor of rooms chained
    along x, one room per submap (room i spans x in [i*Lx, (i+1)*Lx]); neighbouring rooms share a wall plane, so
    consecutive submaps overlap there and the voxel count of the map grows linearly with the number of submaps.
    Everything is a function of (seed, submap_id) only: any rank can regenerate any submap, nothing depends on how the
    submaps are sharded.  ``emb_from``: reuse an existing embedding tensor (the benchmark keeps a small pool of 5 GB
    embedding arrays instead of one per submap: the values do not influence which voxels a point falls into).
    ``projective``: scale of the common SL(4) frame's last row -- small, so that w stays within a few percent of 1 along a
    corridor of 200 rooms (1.6 km); the box-room default 2e-3 would put w = 0 inside the corridor."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + 7919 * submap_id + 17)
    pts, M_room_local = box_room_frames_device(gen, S, H, W, room, noise, 0.37 * submap_id, device, step)
    u = torch.rand((2, S, H, W), dtype=torch.float32, device=device, generator=gen).clamp_min_(1e-12)
    conf = (1.0 - 2.0 * (torch.log(u[0]) + torch.log(u[1]))).contiguous()
    del u
    emb = emb_from
    if emb is None and with_emb:
        emb = torch.empty((S, H, W, d), dtype=emb_dtype, device=device)
        for s in range(S):
            emb[s] = torch.randn((H, W, d), dtype=torch.float32, device=device, generator=gen).to(torch.bfloat16).to(emb_dtype)
    T = np.eye(4)
    T[0, 3] = float(room[0]) * submap_id  # the room's place in the corridor
    rng = np.random.default_rng([seed, submap_id, 3])
    G = synth.random_sl4(np.random.default_rng([seed, 987654321]), projective=projective)
    Hm = G @ T @ M_room_local
    Hm = Hm @ (np.eye(4) + 1e-4 * rng.normal(size=(4, 4)))
    Hm = Hm / abs(np.linalg.det(Hm)) ** 0.25
    paths = [f"left_{submap_id * S + i:07d}.png" for i in range(S)]
    return DeviceSubmapData(submap_id, pts, conf, emb, Hm.astype(np.float64), paths, S - 1)


def make_indexed_device(seed: int, submap_id: int, S: int = 32, H: int = 294, W: int = 518, d: int = 512, n_masks: int = 1024,
                        block: int = 16, emb_dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None):
    """SAM-like indexed embeddings: mask ids (S,H,W) int32, piecewise constant over block x block pixels (0 = no mask),
    and a table (n_masks, d) of unit vectors whose row 0 is zero (semantic_embedder.py:324-349 paints exactly this)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 7919 + submap_id)
    coarse = torch.randint(0, n_masks, (S, (H + block - 1) // block, (W + block - 1) // block), device=device, generator=gen,
                           dtype=torch.int32)
    ids = coarse.repeat_interleave(block, dim=1).repeat_interleave(block, dim=2)[:, :H, :W].contiguous()
    table = torch.randn((n_masks, d), dtype=torch.float32, device=device, generator=gen)
    table = table / table.norm(dim=1, keepdim=True)
    table[0] = 0.0
    return ids, table.to(emb_dtype).contiguous()


def to_submap(data: DeviceSubmapData, host: bool = False, pin: bool = True):
    """vsm.Submap over the generated tensors (device-resident, or copied to pinned host memory)."""
    from .submap import Submap

    sm = Submap(data.submap_id)
    if host:
        def h(t):
            out = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
            out.copy_(t)
            return out
        pts, conf, emb = h(data.points).numpy(), h(data.conf).numpy(), h(data.emb)
        if emb.dtype == torch.float32:
            emb = emb.numpy()
        sm.pointclouds, sm.conf = pts, conf
        sm._dev_cache.clear()
        from . import voxel_map as vm

        sm.conf_threshold = vm.conf_threshold(data.conf, data.conf_percentile)
        sm.add_all_semantic_embeddings(emb)
    else:
        sm.add_all_points(data.points, None, data.conf, data.conf_percentile, None)
        sm.add_all_semantic_embeddings(data.emb)
    sm.set_conf_masks(sm.conf)
    sm.set_reference_homography(data.H_world_map)
    sm.set_frame_ids(data.frame_paths)
    sm.set_last_non_loop_frame_index(data.last_non_loop_frame_index)
    return sm
