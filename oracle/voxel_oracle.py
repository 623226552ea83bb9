"""CPU oracle for the semantic voxel-mapping + text-query hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vggt-slam_b200/`` may import this
module: the product path is CUDA-only and fails loudly without its extension.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

It is a numpy (+ torch-CPU for the query, as the reference uses torch there)
restatement of what the reference computes, written stage by stage so each
stage can be compared against a CUDA kernel.  Every function cites the
reference lines it follows (paths relative to the upstream repo root).

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is the reference itself, executed in the
build container: ``tests/golden/make_golden.py`` imports the upstream package
unmodified, runs its functions on seeded inputs and commits inputs + outputs
as fixtures; ``tests/test_oracle_golden.py`` checks this file against them.
Pinned environment: numpy 2.3.x, torch 2.11 (percentile index arithmetic and
weak-scalar promotion are numpy-2 semantics).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# a1: confidence threshold (vggt_slam/submap.py:38)
# --------------------------------------------------------------------------


def conf_threshold(conf: np.ndarray, percentile: float):
    """``np.percentile`` over the whole (S,H,W) confidence volume, loop frames
    included (vggt_slam/submap.py:34-39)."""
    return np.percentile(conf, percentile)


def percentile_linear_restated(values: np.ndarray, q_percent: float):
    """Explicit restatement of numpy-2 ``np.percentile(values, q)`` (method
    'linear') for a 1-D float32 array, spelling out the arithmetic dtype of
    every step so that a device kernel can mirror it:

      q      = f32(q_percent) / f32(100)                (weak python scalar)
      vidx   = f32(n-1) * q                             (float32 product)
      lo     = floor(vidx), hi = lo+1  (both -> last element if vidx >= n-1)
      gamma  = vidx - lo                                (float32)
      r      = a + (b-a)*gamma          ; if gamma >= 0.5: r = b - (b-a)*(1-gamma)
      any NaN in the data -> NaN.

    Used by the bbox filter of the global builder (vggt_slam/map.py:257-258)
    and by ``conf_threshold``.  Checked against np.percentile in the tests.
    """
    v = np.asarray(values)
    assert v.dtype == np.float32 and v.ndim == 1 and v.size > 0
    n = v.size
    q = np.float32(q_percent) / np.float32(100)
    vidx = np.float32(n - 1) * q
    if np.isnan(v).any():
        return np.float32(np.nan)
    s = np.sort(v)
    if vidx >= n - 1:
        lo = hi = n - 1
    elif vidx < 0:
        lo = hi = 0
    else:
        lo = int(np.floor(vidx))
        hi = lo + 1
    a, b = s[lo], s[hi]
    gamma = np.float32(vidx - np.floor(vidx))
    diff = np.float32(b - a)
    r = np.float32(a + np.float32(diff * gamma))
    if gamma >= np.float32(0.5):
        r = np.float32(b - np.float32(diff * np.float32(np.float32(1) - gamma)))
    return r


# --------------------------------------------------------------------------
# a4/a5: confidence filtering and world-frame points
# --------------------------------------------------------------------------


def filter_by_confidence(data: np.ndarray, conf: np.ndarray, thr, stride: int = 1):
    """Boolean gather in (s,h,w) row-major order (vggt_slam/submap.py:155-164)."""
    if stride == 1:
        return data[conf >= thr]
    return data[:, ::stride, ::stride, :][conf[:, ::stride, ::stride] >= thr]


def homography_apply_f64(points_n3: np.ndarray, H: np.ndarray) -> np.ndarray:
    """(H @ [p;1]) / w in float64 (vggt_slam/submap.py:185-188, 171-174;
    vggt_slam/map.py:232-234).  The product is numpy's matmul, as upstream."""
    pts = np.asarray(points_n3).reshape(-1, 3)
    ones = np.ones((pts.shape[0], 1), dtype=pts.dtype)
    hom = np.concatenate([pts, ones], axis=1)
    out = (np.asarray(H) @ hom.T).T
    return out[:, :3] / out[:, 3:]


def homography_apply_fma_chain(points_n3: np.ndarray, H: np.ndarray) -> np.ndarray:
    """The same transform written as the float64 FMA chain the CUDA kernel
    uses: r_i = fma(H[i,3], 1, fma(H[i,2], z, fma(H[i,1], y, H[i,0]*x))).
    Pure-python loop with an exact-rational FMA (correctly rounded): small inputs only.  Exists to show that the
    BLAS k-order assumption of the kernel holds on this host."""
    from fractions import Fraction

    def fma(a, b, c):  # exact a*b+c, one rounding (int/int true division is correctly rounded)
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    pts = np.asarray(points_n3, dtype=np.float32).reshape(-1, 3)
    H = np.asarray(H, dtype=np.float64)
    out = np.empty((pts.shape[0], 3), dtype=np.float64)
    for n in range(pts.shape[0]):
        x, y, z = (float(pts[n, 0]), float(pts[n, 1]), float(pts[n, 2]))
        r = []
        for i in range(4):
            acc = H[i, 0] * x
            acc = fma(H[i, 1], y, acc)
            acc = fma(H[i, 2], z, acc)
            acc = fma(H[i, 3], 1.0, acc)
            r.append(acc)
        with np.errstate(all="ignore"):
            out[n] = np.array(r[:3]) / np.float64(r[3])
    return out


def points_in_world_frame(points, conf, thr, H, stride: int = 1) -> np.ndarray:
    """Submap.get_points_in_world_frame (vggt_slam/submap.py:182-188): float64."""
    return homography_apply_f64(filter_by_confidence(points, conf, thr, stride), H)


def points_list_in_world_frame(points, conf_masks, thr, H, frame_ids, last_non_loop=None,
                               ignore_loop_closure_frames=False):
    """Submap.get_points_list_in_world_frame (vggt_slam/submap.py:166-180):
    per-frame unfiltered float64 (H,W,3) maps, frame ids and per-frame masks."""
    out_pts, out_ids, out_masks = [], [], []
    for i, frame in enumerate(points):
        out_pts.append(homography_apply_f64(frame, H).reshape(frame.shape))
        out_ids.append(frame_ids[i])
        out_masks.append(conf_masks[i] >= thr)
        if ignore_loop_closure_frames and i == last_non_loop:
            break
    return out_pts, out_ids, out_masks


# --------------------------------------------------------------------------
# voxel keys
# --------------------------------------------------------------------------


def world_points_f32(points_n3: np.ndarray, H: np.ndarray) -> np.ndarray:
    """float64 transform rounded to float32 (vggt_slam/submap.py:279,
    vggt_slam/map.py:234)."""
    with np.errstate(all="ignore"):
        return homography_apply_f64(points_n3, H).astype(np.float32)


def voxel_keys(points_world_f32: np.ndarray, cell: float) -> np.ndarray:
    """floor(p / cell) -> int64, float32 division by the weak python scalar
    (vggt_slam/submap.py:282, vggt_slam/map.py:274, 351)."""
    with np.errstate(all="ignore"):
        return np.floor(points_world_f32 / cell).astype(np.int64)


def voxelise(keys_n3: np.ndarray, feats_nd: np.ndarray, cell: float, exact_order: bool = True):
    """unique rows (lexicographic signed order) + scatter-add + mean
    (vggt_slam/submap.py:283-293, vggt_slam/map.py:351-362).

    exact_order=True follows the reference (``np.add.at`` = sequential float32
    adds in point order).  exact_order=False sums in float64 with a sort-based
    reduction: same mathematics, ~50x faster, for large parity cases whose
    tolerance (1e-3 rel) does not care about float32 summation order."""
    uniq, inverse = np.unique(keys_n3, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    V = uniq.shape[0]
    d = feats_nd.shape[1]
    counts = np.zeros((V,), dtype=np.int64)
    np.add.at(counts, inverse, 1)
    if exact_order:
        sums = np.zeros((V, d), dtype=np.float32)
        np.add.at(sums, inverse, feats_nd.astype(np.float32))
    else:
        order = np.argsort(inverse, kind="stable")
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        sums = np.add.reduceat(feats_nd[order].astype(np.float64), starts, axis=0)
    mean = sums / counts[:, None]  # float64, like upstream
    centers = ((uniq.astype(np.float32) + 0.5) * cell).astype(np.float32)
    return uniq, inverse, counts, sums, mean, centers


# --------------------------------------------------------------------------
# a6: per-submap fusion (vggt_slam/submap.py:221-311)
# --------------------------------------------------------------------------


@dataclass
class OracleSubmap:
    """The fields of upstream ``Submap`` that the hot path reads
    (vggt_slam/submap.py:11-29)."""

    submap_id: int
    points: np.ndarray  # (S,H,W,3) f32
    conf: np.ndarray  # (S,H,W) f32
    conf_threshold: object  # np.float32 scalar from np.percentile
    emb: Optional[np.ndarray]  # (S,H,W,d)
    H_world_map: Optional[np.ndarray]  # (4,4) f64
    frame_ids: Optional[List[float]] = None
    frame_id_to_name: Optional[Dict[str, str]] = None
    last_non_loop_frame_index: Optional[int] = None


@dataclass
class OracleVoxels:
    voxel_size: float
    coords: np.ndarray  # (V,3) int64 true keys (not exposed upstream)
    centers_world: np.ndarray  # (V,3) f32
    features: np.ndarray  # (V,d) f64 (numpy path)
    counts: np.ndarray  # (V,) int64
    contributors: List[List[Tuple[int, str]]]
    frame_name_maps: Dict[str, Dict[str, str]] = field(default_factory=dict)
    n_points: int = 0


def fuse_submap(sm: OracleSubmap, voxel_size: float, ignore_loop_closure_frames: bool = False,
                exact_order: bool = True, with_contributors: bool = True) -> OracleVoxels:
    """Submap.get_semantic_voxel_in_world_frame (vggt_slam/submap.py:221-311).
    No outlier filters; contributors hold one tuple per *point*."""
    if voxel_size <= 0.0:
        raise ValueError("voxel_size must be > 0")
    if sm.points is None or sm.emb is None or sm.H_world_map is None:
        raise RuntimeError("submap is missing points / embeddings / homography")
    end = sm.points.shape[0]
    if ignore_loop_closure_frames and sm.last_non_loop_frame_index is not None:
        end = min(end, sm.last_non_loop_frame_index + 1)
    pts, emb, conf = sm.points[:end], sm.emb[:end], sm.conf[:end]
    mask = conf >= sm.conf_threshold
    p = pts[mask]
    e = emb[mask]
    d = emb.shape[-1]
    if p.shape[0] == 0:
        return OracleVoxels(voxel_size, np.zeros((0, 3), np.int64), np.zeros((0, 3), np.float32),
                            np.zeros((0, d), np.float32), np.zeros((0,), np.int64), [])
    frame_of_point = np.nonzero(mask)[0].astype(np.int32)
    pw = world_points_f32(p, sm.H_world_map)
    keys = voxel_keys(pw, voxel_size)
    uniq, inverse, counts, _sums, mean, centers = voxelise(keys, e, voxel_size, exact_order)
    contributors: List[List[Tuple[int, str]]] = [[] for _ in range(uniq.shape[0])]
    if with_contributors:
        sid = int(sm.submap_id)
        for f, v in zip(frame_of_point.tolist(), inverse.tolist()):
            if sm.frame_ids is not None and f < len(sm.frame_ids):
                fid = str(sm.frame_ids[f])
            else:
                fid = str(int(f))
            contributors[v].append((sid, fid))
    return OracleVoxels(voxel_size, uniq, centers, mean, counts, contributors, n_points=int(p.shape[0]))


# --------------------------------------------------------------------------
# a7: global fusion with the three per-submap filters (vggt_slam/map.py:170-381)
# --------------------------------------------------------------------------


def submap_observations(sm: OracleSubmap, voxel_size: float, stride: int = 1,
                        ignore_loop_closure_frames: bool = True, stages: Optional[dict] = None):
    """One iteration of the per-submap loop (vggt_slam/map.py:196-291): returns
    (pts_world f32 (N,3), feats (N,d), frame_index (N,) int32) after the
    finite / percentile-bbox / coarse-cell filters, or None if the submap is
    skipped.  ``stages`` (optional dict) receives the survivor counts and
    bounds of each filter for stage-wise kernel checks."""
    if sm.emb is None or sm.points is None or sm.conf is None or sm.conf_threshold is None:
        return None
    if sm.H_world_map is None:
        return None
    end = sm.points.shape[0]
    if ignore_loop_closure_frames and sm.last_non_loop_frame_index is not None:
        end = min(end, sm.last_non_loop_frame_index + 1)
    pts, emb, conf = sm.points[:end], sm.emb[:end], sm.conf[:end]
    if stride > 1:
        pts, emb, conf = pts[:, ::stride, ::stride, :], emb[:, ::stride, ::stride, :], conf[:, ::stride, ::stride]
    mask = conf >= sm.conf_threshold
    p, e = pts[mask], emb[mask]
    if p.shape[0] == 0:
        return None
    fidx = np.nonzero(mask)[0].astype(np.int32)
    pw = world_points_f32(p, sm.H_world_map)
    if stages is not None:
        stages["n_conf"] = int(pw.shape[0])
    # filter 1: finite rows (map.py:247-251)
    ok = np.isfinite(pw).all(axis=1) & np.isfinite(e).all(axis=1)
    if not ok.all():
        pw, e, fidx = pw[ok], e[ok], fidx[ok]
    if stages is not None:
        stages["n_finite"] = int(pw.shape[0])
    if pw.shape[0] == 0:
        return None
    # filter 2: inclusive [0.5, 99.5] percentile box (map.py:257-263)
    lo = np.percentile(pw, 0.5, axis=0)
    hi = np.percentile(pw, 99.5, axis=0)
    ok = (pw >= lo).all(axis=1) & (pw <= hi).all(axis=1)
    if not ok.all():
        pw, e, fidx = pw[ok], e[ok], fidx[ok]
    if stages is not None:
        stages["lo"], stages["hi"], stages["n_bbox"] = lo, hi, int(pw.shape[0])
    if pw.shape[0] == 0:
        return None
    # filter 3: coarse occupancy cells of 3*voxel_size with < 10 points (map.py:271-280)
    coarse = float(voxel_size) * 3.0
    if coarse > 0.0:
        ck = voxel_keys(pw, coarse)
        _, inv, cnt = np.unique(ck, axis=0, return_inverse=True, return_counts=True)
        ok = cnt[inv.reshape(-1)] >= 10
        if not ok.all():
            pw, e, fidx = pw[ok], e[ok], fidx[ok]
    if stages is not None:
        stages["n_coarse"] = int(pw.shape[0])
    if pw.shape[0] == 0:
        return None
    return pw, e.astype(np.float32), fidx


def build_global(submaps: Sequence[OracleSubmap], voxel_size: float, stride: int = 1,
                 ignore_loop_closure_frames: bool = True, deduplicate_contributors: bool = True,
                 exact_order: bool = True, with_contributors: bool = True) -> OracleVoxels:
    """GraphMap.build_semantic_voxel_map, numpy branch (vggt_slam/map.py:170-381)."""
    if voxel_size <= 0.0:
        raise ValueError("voxel_size must be > 0")
    if stride < 1:
        raise ValueError("stride must be >= 1")
    P, F, SID, FID = [], [], [], []
    names: Dict[str, Dict[str, str]] = {}
    for sm in sorted(submaps, key=lambda s: s.submap_id):
        obs = submap_observations(sm, voxel_size, stride, ignore_loop_closure_frames)
        if obs is None:
            continue
        pw, e, fidx = obs
        P.append(pw)
        F.append(e)
        SID.append(np.full((pw.shape[0],), int(sm.submap_id), dtype=np.int32))
        FID.append(np.array([str(sm.frame_ids[int(i)]) for i in fidx], dtype=object))
        if sm.frame_id_to_name is not None:
            names[str(int(sm.submap_id))] = dict(sm.frame_id_to_name)
    if not P:
        return OracleVoxels(float(voxel_size), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.float32),
                            np.zeros((0, 0), np.float32), np.zeros((0,), np.int64), [], names)
    pw = np.concatenate(P, axis=0)
    e = np.concatenate(F, axis=0)
    sid = np.concatenate(SID, axis=0)
    fid = np.concatenate(FID, axis=0)
    keys = voxel_keys(pw, voxel_size)
    uniq, inverse, counts, _sums, mean, centers = voxelise(keys, e, voxel_size, exact_order)
    V = uniq.shape[0]
    contributors: List[List[Tuple[int, str]]] = [[] for _ in range(V)]
    if with_contributors:
        if deduplicate_contributors:
            sets = [set() for _ in range(V)]
            for i, v in enumerate(inverse.tolist()):
                sets[v].add((int(sid[i]), str(fid[i])))
            contributors = [sorted(s) for s in sets]
        else:
            for i, v in enumerate(inverse.tolist()):
                contributors[v].append((int(sid[i]), str(fid[i])))
    return OracleVoxels(float(voxel_size), uniq, centers, mean, counts, contributors, names,
                        n_points=int(pw.shape[0]))


# --------------------------------------------------------------------------
# a11-a14: SemanticVoxelMap lookups and query (vggt_slam/semantic_voxel.py)
# --------------------------------------------------------------------------


def coords_from_centers(centers_world: np.ndarray, voxel_size: float) -> np.ndarray:
    """The lossy float32 reconstruction of integer coords
    (vggt_slam/semantic_voxel.py:62-66); differs from the true keys for a
    fraction of voxels."""
    return np.floor(centers_world / voxel_size - 0.5).astype(np.int64)


def position_to_coord(position_world, voxel_size: float) -> Tuple[int, int, int]:
    """vggt_slam/semantic_voxel.py:68-72."""
    p = np.asarray(position_world, dtype=np.float32).reshape(3)
    c = np.floor(p / voxel_size).astype(np.int64)
    return int(c[0]), int(c[1]), int(c[2])


def coord_index(recon_coords: np.ndarray) -> Dict[Tuple[int, int, int], int]:
    """dict built in ascending index order: a later duplicate coordinate
    overwrites an earlier one (vggt_slam/semantic_voxel.py:39-41)."""
    return {(int(c[0]), int(c[1]), int(c[2])): i for i, c in enumerate(recon_coords)}


def query(features: np.ndarray, qe: np.ndarray, top_k: int = 1):
    """SemanticVoxelMap.query_with_embedding for ONE prompt
    (vggt_slam/semantic_voxel.py:97-116): float32 dot products, torch.topk.
    Voxel features are not normalised (Appendix A-1)."""
    import torch

    f = torch.from_numpy(np.ascontiguousarray(features)).float()
    q = torch.from_numpy(np.asarray(qe)).float()
    if q.ndim == 1:
        q = q[None, :]
    sims = torch.matmul(f, q.T).squeeze(-1)
    idx = torch.topk(sims, top_k).indices.tolist()
    return idx, sims[idx].tolist(), sims.numpy()


def query_normalised(features: np.ndarray, qe: np.ndarray, top_k: int = 1, eps: float = 1e-12):
    """Opt-in cosine variant named by BASELINE.json's north_star ("normalises
    voxel embeddings"): f / max(||f||, eps) in float32, then the same scoring.
    Not a reference behaviour (upstream never normalises voxel features)."""
    f = np.asarray(features, dtype=np.float32)
    nrm = np.sqrt((f.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    fn = f / np.maximum(nrm, np.float32(eps))[:, None]
    return query(fn, qe, top_k)


def latest_contributor(contribs: List[Tuple[int, str]]) -> Tuple[int, str]:
    """max over (submap_id:int, frame_id:str) with *string* ordering of frame
    ids (vggt_slam/semantic_voxel.py:118-126, Appendix A-6)."""
    return sorted(contribs, key=lambda x: (x[0], x[1]), reverse=True)[0]


def frame_id_from_name(path: str) -> float:
    """First integer/decimal in the basename as float (vggt_slam/submap.py:109-131)."""
    import os
    import re

    name = os.path.basename(path)
    m = re.search(r"\d+(?:\.\d+)?", name)
    if not m:
        raise ValueError(f"No number found in image name: {name}")
    return float(m.group())
