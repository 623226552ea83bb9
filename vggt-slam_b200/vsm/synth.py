"""Seeded synthetic inputs of the shape the hot path consumes.

Host-side numpy generators used by the golden-vector script, the parity tests
and the benchmark's CPU sample.  The device-scale generator (same geometry
model, counter-based RNG) lives in csrc/synth_kernels.cuh and is reached with
``vsm._native.synth_submap``.

Geometry: "box room + trajectory".  Every pixel of every frame is the first
hit of its pinhole ray with the inside of an axis-aligned room, seen from a
camera that walks along a path; points are expressed in the submap-local frame
(the first camera's frame) with Gaussian range noise, which gives realistic
voxel occupancy instead of uniform noise.  Confidence is 1 + Gamma(2, 2) like
VGGT's ``1 + exp(.)`` head.  Embeddings are N(0,1) rounded to bf16 (so float32
and bf16 pipelines see identical values) or SAM-like piecewise-constant unit
vectors with zero background.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 values to the nearest bf16 (ties to even), returned as
    float32.  NaN/Inf pass through."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    finite = np.isfinite(x)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = np.where(finite, r, x.view(np.uint32))
    return out.view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """bf16 bit patterns (uint16) of values that are already bf16-exact."""
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) >> 16).astype(np.uint16)


def _rot(axis: np.ndarray, angle: float) -> np.ndarray:
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def random_sl4(rng: np.random.Generator, eps: float = 0.05, projective: float = 2e-3) -> np.ndarray:
    """Random 4x4 with det = 1 close to a rigid motion, with a small projective
    last row (SL(4) mode: vggt_slam/graph.py:28-29 optimises 15-dof homographies)."""
    R = _rot(rng.normal(size=3), rng.uniform(-0.5, 0.5))
    H = np.eye(4)
    H[:3, :3] = R @ (np.eye(3) + eps * rng.normal(size=(3, 3)))
    H[:3, 3] = rng.uniform(-1.0, 1.0, size=3)
    H[3, :3] = projective * rng.normal(size=3)
    det = np.linalg.det(H)
    H = H / np.sign(det) / abs(det) ** 0.25
    return H.astype(np.float64)


def random_sim3(rng: np.random.Generator, scale: Optional[float] = None) -> np.ndarray:
    """Rotation + translation (+ optional uniform scale); last row exactly
    [0,0,0,1] (Sim(3)/SE(3) mode: vggt_slam/graph_se3.py, map.py:383-396)."""
    H = np.eye(4)
    s = 1.0 if scale is None else scale
    H[:3, :3] = s * _rot(rng.normal(size=3), rng.uniform(-np.pi, np.pi))
    H[:3, 3] = rng.uniform(-2.0, 2.0, size=3)
    return H.astype(np.float64)


@dataclass
class SynthSubmap:
    submap_id: int
    points: np.ndarray  # (S,H,W,3) f32, submap-local frame
    conf: np.ndarray  # (S,H,W) f32
    colors: np.ndarray  # (S,H,W,3) u8
    emb: np.ndarray  # (S,H,W,d) f32 (bf16-exact values)
    H_world_map: np.ndarray  # (4,4) f64
    frame_paths: List[str]
    last_non_loop_frame_index: int
    conf_percentile: float = 25.0


def box_room_frames(rng: np.random.Generator, S: int, H: int, W: int, room=(6.0, 4.0, 3.0),
                    noise: float = 0.005, start: float = 0.0):
    """(S,H,W,3) float32 points in the first camera's frame, and the 4x4
    local->room pose of that first camera."""
    room = np.asarray(room, dtype=np.float64)
    fx = 0.8 * W
    u = (np.arange(W) - (W - 1) / 2.0) / fx
    v = (np.arange(H) - (H - 1) / 2.0) / fx
    dirs_cam = np.stack(np.broadcast_arrays(u[None, :], v[:, None], np.ones((H, W))), axis=-1)
    dirs_cam /= np.linalg.norm(dirs_cam, axis=-1, keepdims=True)
    pts = np.empty((S, H, W, 3), dtype=np.float64)
    cams = []
    for s in range(S):
        t = start + 0.08 * s
        centre = room * (0.5 + 0.22 * np.array([np.cos(t), np.sin(1.3 * t), 0.3 * np.sin(0.7 * t)]))
        yaw = 0.9 * t + 0.3
        R = _rot(np.array([0.0, 0.0, 1.0]), yaw) @ _rot(np.array([1.0, 0.0, 0.0]), -np.pi / 2 + 0.15 * np.sin(t))
        cams.append((R, centre))
        d = dirs_cam @ R.T
        with np.errstate(divide="ignore", invalid="ignore"):
            t_hi = (room - centre) / d
            t_lo = (0.0 - centre) / d
        tt = np.where(d > 0, t_hi, t_lo)
        tt = np.where(np.isfinite(tt) & (tt > 0), tt, np.inf)
        depth = tt.min(axis=-1)
        depth = depth + noise * rng.normal(size=depth.shape)
        pts[s] = centre + d * depth[..., None]
    R0, c0 = cams[0]
    local = (pts - c0) @ R0  # room -> first-camera frame
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R0, c0
    return local.astype(np.float32), M


def painted_embeddings(rng: np.random.Generator, S: int, H: int, W: int, d: int, n_masks: int = 12,
                       block: int = 7) -> np.ndarray:
    """SAM-like dense map: blocks painted with one of ``n_masks`` unit vectors,
    ~25 % background left at zero (vggt_slam/semantic_embedder.py:324-349)."""
    table = rng.normal(size=(n_masks, d))
    table /= np.linalg.norm(table, axis=1, keepdims=True)
    table = round_to_bf16(table.astype(np.float32))
    table = np.concatenate([np.zeros((1, d), np.float32), table], axis=0)
    hb, wb = -(-H // block), -(-W // block)
    ids = rng.integers(0, n_masks + 1, size=(S, hb, wb))
    ids = np.where(rng.random(size=ids.shape) < 0.25, 0, ids)
    ids = np.repeat(np.repeat(ids, block, axis=1), block, axis=2)[:, :H, :W]
    return table[ids]


def make_submap(seed: int, submap_id: int, S: int = 4, H: int = 28, W: int = 42, d: int = 16, mode: str = "sl4",
                emb_kind: str = "normal", room=(6.0, 4.0, 3.0), n_loop_frames: int = 0, bad_fraction: float = 0.0,
                start: float = 0.0, noise: float = 0.005, name_fmt: str = "left_{:06d}.png",
                first_frame_number: int = 0) -> SynthSubmap:
    """One submap's worth of hot-path inputs.  ``n_loop_frames`` trailing
    frames play the loop-closure frames (zero embeddings, solver.py:459-463).
    ``bad_fraction`` of the points / embedding rows get NaN or Inf injected."""
    rng = np.random.default_rng([seed, submap_id])
    pts, M_room_local = box_room_frames(rng, S, H, W, room=room, noise=noise, start=start)
    conf = (1.0 + rng.gamma(2.0, 2.0, size=(S, H, W))).astype(np.float32)
    colors = rng.integers(0, 256, size=(S, H, W, 3), dtype=np.uint8)
    if emb_kind == "normal":
        emb = round_to_bf16(rng.normal(size=(S, H, W, d)).astype(np.float32))
    elif emb_kind == "painted":
        emb = painted_embeddings(rng, S, H, W, d)
    else:
        raise ValueError(emb_kind)
    if n_loop_frames:
        emb[S - n_loop_frames:] = 0.0
    if bad_fraction > 0:
        n_bad = max(2, int(bad_fraction * S * H * W))
        idx = rng.integers(0, S * H * W, size=n_bad)
        flat_p = pts.reshape(-1, 3)
        flat_e = emb.reshape(-1, d)
        half = n_bad // 2
        flat_p[idx[:half], rng.integers(0, 3, size=half)] = np.where(rng.random(half) < 0.5, np.nan, np.inf)
        flat_e[idx[half:], rng.integers(0, d, size=n_bad - half)] = np.where(rng.random(n_bad - half) < 0.5,
                                                                             np.nan, -np.inf)
    # world <- room transform shared by every submap of a seed (so submaps overlap
    # in the world frame), times this submap's local -> room pose.
    grng = np.random.default_rng([seed, 987654321])
    G = random_sl4(grng) if mode == "sl4" else random_sim3(grng, scale=None if mode == "se3" else 1.7)
    Hm = G @ M_room_local
    if mode == "sl4":  # per-submap 15-dof perturbation, as the pose graph would leave it
        Hm = Hm @ (np.eye(4) + 1e-3 * rng.normal(size=(4, 4)))
        Hm = Hm / abs(np.linalg.det(Hm)) ** 0.25
    elif mode == "se3":
        Hm[3, :] = [0.0, 0.0, 0.0, 1.0]
    n_real = S - n_loop_frames
    paths = [name_fmt.format(first_frame_number + i) for i in range(n_real)]
    return SynthSubmap(submap_id, pts, conf, colors, emb, Hm, paths, n_real - 1)
