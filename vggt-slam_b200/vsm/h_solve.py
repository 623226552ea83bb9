"""Projective RANSAC between two point maps with the reference's API (vggt_slam/h_solve.py), the scoring of the
hypotheses on the GPU (vsm_ransac_score, csrc/ransac.cu).

The minimal solver (null space of a (3*5) x 16 system per hypothesis, h_solve.py:43-93) runs on the host with
numpy / scipy, exactly where the reference runs it; what the reference then does with three (B, N, 3) float32 torch
tensors -- apply every hypothesis to every point, count inliers, take the argmax (h_solve.py:16-41, 150-160) -- is one
kernel here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .voxel_map import _ptr, _stream_ptr, as_device, require_cuda


def to_homogeneous(X):
    return np.hstack([X, np.ones((X.shape[0], 1))])


def apply_homography(H, X, debug=False):
    Xt = (H @ to_homogeneous(X).T).T
    return Xt[:, :3] / Xt[:, 3:]


def estimate_3D_homography(X_src_batch: np.ndarray, X_dst_batch: np.ndarray) -> np.ndarray:
    """(B,n,3) x2 -> (B,4,4) float32 hypotheses, det-normalised; identity where the system is degenerate
    (h_solve.py:43-93).  Host code, as upstream."""
    from scipy.linalg import null_space

    B, n, _ = X_src_batch.shape
    ones = np.ones((B, n))
    x, y, z = X_src_batch[:, :, 0], X_src_batch[:, :, 1], X_src_batch[:, :, 2]
    xp, yp, zp = X_dst_batch[:, :, 0], X_dst_batch[:, :, 1], X_dst_batch[:, :, 2]
    A = np.zeros((B, 3 * n, 16))
    src = np.stack([x, y, z, ones], axis=2)
    for row, p in enumerate((xp, yp, zp)):
        A[:, row::3, 4 * row:4 * row + 4] = -src
        A[:, row::3, 12:16] = np.stack([x * p, y * p, z * p, p], axis=2)
    H_batch = np.zeros((B, 4, 4))
    for i in range(B):
        nv = null_space(A[i])
        if nv.shape[1] == 0:
            H_batch[i] = np.eye(4)
            continue
        H = nv[:, 0].reshape(4, 4)
        if H[3, 3] == 0:
            H_batch[i] = np.eye(4)
            continue
        H = H / H[3, 3]
        det = np.linalg.det(H)
        H_batch[i] = np.eye(4) if (np.isnan(det) or det < 0.0001) else H / det ** 0.25
    return H_batch.astype(np.float32)


def score_hypotheses(H_batch, X1, X2, threshold: float, device: Optional[torch.device] = None
                     ) -> Tuple[torch.Tensor, int, int]:
    """Inlier counts (B,) int32 device tensor, best index, best count: vsm_ransac_score.  Inputs may be numpy arrays or
    torch tensors (CUDA tensors are used in place: the point maps of the producer hand-off never leave the device)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    Ht = as_device(H_batch, dev, torch.float32).reshape(-1, 16)
    x1 = as_device(X1, dev, torch.float32).reshape(-1, 3)
    x2 = as_device(X2, dev, torch.float32).reshape(-1, 3)
    if x1.shape != x2.shape:
        raise ValueError(f"point sets differ in shape: {tuple(x1.shape)} vs {tuple(x2.shape)}")
    B = int(Ht.shape[0])
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    best = torch.empty((2,), dtype=torch.int32, device=dev)
    bi, bc = C.c_int32(0), C.c_int32(0)
    N.check(N.lib.vsm_ransac_score(_ptr(Ht), _ptr(x1), _ptr(x2), int(x1.shape[0]), B, float(threshold), _ptr(counts),
                                   _ptr(best), C.byref(bi), C.byref(bc), _stream_ptr(dev)))
    return counts, int(bi.value), int(bc.value)


def ransac_projective(X1_np, X2_np, threshold=0.01, max_iter=300, sample_size=5,
                      generator: Optional[np.random.Generator] = None) -> np.ndarray:
    """Best of `max_iter` 5-point hypotheses by inlier count (h_solve.py:132-163); returns the 4x4 float32 matrix.
    The sample indices come from `generator` (numpy) instead of torch's CUDA generator."""
    rng = np.random.default_rng() if generator is None else generator
    X1h = X1_np.detach().cpu().numpy() if isinstance(X1_np, torch.Tensor) else np.asarray(X1_np)
    X2h = X2_np.detach().cpu().numpy() if isinstance(X2_np, torch.Tensor) else np.asarray(X2_np)
    n = X1h.shape[0]
    idx = rng.integers(0, n, size=(max_iter, sample_size))
    H = estimate_3D_homography(X1h[idx].astype(np.float32), X2h[idx].astype(np.float32))
    _, best, _ = score_hypotheses(H, X1_np, X2_np, threshold)
    return H[best]
