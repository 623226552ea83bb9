"""CPU restatements (TEST INFRASTRUCTURE ONLY -- never imported by the product path) of the three "next" rows of
SURVEY.md 8f that round 2 builds:

  score_hypotheses        vggt_slam/h_solve.py:16-41 (apply_homography_batch), :150-160 (errors, inlier counts, argmax)
  build_occupancy         get_occupancy.py:130-179 (build_occupancy_from_pointcloud)
  unproject_depth         vggt.utils.geometry.unproject_depth_map_to_point_map as called at vggt_slam/solver.py:254-256;
                          the dependency (facebookresearch/vggt, installed from git, unpinned in requirements.txt) is
                          NOT in /root/reference, so this follows its published algorithm: PARITY UNPINNED
  images_to_colors        vggt_slam/solver.py:260

Pinned against the reference itself by tests/golden/case_f_ransac.npz and case_g_occupancy.npz
(tests/test_oracle_golden_r2.py); `unproject_depth` has no reference output to pin against.
"""
from __future__ import annotations

import numpy as np


def score_hypotheses(H: np.ndarray, X1: np.ndarray, X2: np.ndarray, threshold: float):
    """(B,4,4) float32 hypotheses, (N,3) float32 point pairs -> (errors (B,N) float32, inlier counts (B,) int64, argmax).
    float32 throughout, like the torch code: bmm(H, [X;1]^T), divide by the w row, 2-norm of the difference."""
    H = np.asarray(H, dtype=np.float32)
    X1 = np.asarray(X1, dtype=np.float32)
    X2 = np.asarray(X2, dtype=np.float32)
    Xh = np.concatenate([X1, np.ones((X1.shape[0], 1), dtype=np.float32)], axis=1)  # h_solve.py:31-32
    Xt = np.matmul(H, Xh.T[None, :, :])                                              # :35-36  (B,4,N)
    with np.errstate(all="ignore"):
        pred = (Xt[:, :3, :] / Xt[:, 3:4, :]).transpose(0, 2, 1)                     # :39-41  (B,N,3)
        diff = pred - X2[None, :, :]
        errors = np.sqrt((diff * diff).sum(axis=2, dtype=np.float32)).astype(np.float32)  # :153
        counts = (errors < np.float32(threshold)).sum(axis=1).astype(np.int64)       # :156-157
    return errors, counts, int(np.argmax(counts))                                    # :160 (first maximum)


def build_occupancy(points_xyz: np.ndarray, voxel_size: float, ceiling_z: float, height_thresh: float):
    """-> centers (M,3) f32, is_blocked (M,) bool, cell_keys (M,2) int64, minz (M,) f32, in np.unique(axis=0) order."""
    pts = np.asarray(points_xyz, dtype=np.float32)
    pts = pts[np.isfinite(pts).all(axis=1)]                       # get_occupancy.py:146
    pts = pts[pts[:, 2] <= ceiling_z]                             # :148
    if pts.shape[0] == 0:                                         # :151-157
        return (np.zeros((0, 3), np.float32), np.zeros((0,), bool), np.zeros((0, 2), np.int64), np.zeros((0,), np.float32))
    ix = np.floor(pts[:, 0] / voxel_size).astype(np.int64)        # :158-159 (float32 division)
    iy = np.floor(pts[:, 1] / voxel_size).astype(np.int64)
    uniq, inv = np.unique(np.stack([ix, iy], axis=1), axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    m = uniq.shape[0]
    z = pts[:, 2].astype(np.float32)
    minz = np.full((m,), np.inf, dtype=np.float32)
    maxz = np.full((m,), -np.inf, dtype=np.float32)
    np.minimum.at(minz, inv, z)                                   # :166-167
    np.maximum.at(maxz, inv, z)
    blocked = (maxz - minz) > float(height_thresh)                # :170-172
    centers = np.zeros((m, 3), dtype=np.float32)
    centers[:, 0] = (uniq[:, 0].astype(np.float32) + 0.5) * float(voxel_size)
    centers[:, 1] = (uniq[:, 1].astype(np.float32) + 0.5) * float(voxel_size)
    centers[:, 2] = minz + float(voxel_size) * 0.5
    return centers, blocked, uniq, minz


def closed_form_inverse_se3(extrinsic: np.ndarray) -> np.ndarray:
    """(S,3,4) or (S,4,4) world-to-camera -> (S,4,4) float64 camera-to-world, R^T and -R^T t formed in the input dtype
    (vggt.utils.geometry.closed_form_inverse_se3 as used at solver.py:263)."""
    R = extrinsic[:, :3, :3]
    T = extrinsic[:, :3, 3:]
    Rt = np.transpose(R, (0, 2, 1))
    out = np.tile(np.eye(4), (len(R), 1, 1))
    out[:, :3, :3] = Rt
    out[:, :3, 3:] = -np.matmul(Rt, T)
    return out


def unproject_depth(depth: np.ndarray, extrinsic: np.ndarray, intrinsic: np.ndarray) -> np.ndarray:
    """depth (S,H,W,1) or (S,H,W) float32, extrinsic (S,3,4), intrinsic (S,3,3) -> world points (S,H,W,3) float64."""
    depth = np.asarray(depth)
    if depth.ndim == 4:
        depth = depth[..., 0]
    S, H, W = depth.shape
    c2w = closed_form_inverse_se3(np.asarray(extrinsic))
    out = []
    for s in range(S):
        K = intrinsic[s]
        fu, fv, cu, cv = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        u, v = np.meshgrid(np.arange(W), np.arange(H))
        x = (u - cu) * depth[s] / fu
        y = (v - cv) * depth[s] / fv
        cam = np.stack((x, y, depth[s]), axis=-1).astype(np.float32)
        out.append(np.dot(cam, c2w[s, :3, :3].T) + c2w[s, :3, 3])
    return np.stack(out, axis=0)


def images_to_colors(images: np.ndarray) -> np.ndarray:
    """(S,3,H,W) float32 -> (S,H,W,3) uint8 (solver.py:260)."""
    return (np.asarray(images).transpose(0, 2, 3, 1) * 255).astype(np.uint8)
