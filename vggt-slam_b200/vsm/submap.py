"""Submap with the reference's hot-path API (vggt_slam/submap.py), computing on the GPU.

Only what the semantic voxel path touches is mirrored: storage of points / confidence / colours /
embeddings / the 4x4 world transform, the confidence threshold, world-frame point extraction and the
per-submap semantic voxelisation.  Arrays may be numpy arrays (the reference's contract) or torch CUDA
tensors (the producer keeps VGGT outputs on the device); embeddings may be float32 or bfloat16
(torch.bfloat16 tensors, or numpy uint16 bit patterns with ``embeddings_are_bf16_bits=True``).
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as N
from . import voxel_map as vm
from .semantic_voxel import LazyContributors, SemanticVoxel


def _shape(x) -> Tuple[int, ...]:
    return tuple(int(v) for v in x.shape)


class Submap:
    def __init__(self, submap_id):
        self.submap_id = submap_id
        self.H_world_map = None
        self.R_world_map = None
        self.poses = None
        self.frames = None
        self.vggt_intrinscs = None
        self.retrieval_vectors = None
        self.colors = None  # (S, H, W, 3)
        self.conf = None  # (S, H, W)
        self.conf_masks = None  # (S, H, W)
        self.conf_threshold = None
        self.pointclouds = None  # (S, H, W, 3)
        self.voxelized_points = None
        self.last_non_loop_frame_index = None
        self.frame_ids = None
        self.frame_names = None
        self.frame_id_to_name = None
        self.semantic_embeddings = None  # (S, H, W, d) -- or the (M, d) table when semantic_index is set
        self.semantic_index = None       # (S, H, W) integer rows of that table (indexed embeddings)
        self.embeddings_are_bf16_bits = False
        self._dev_cache: Dict[str, torch.Tensor] = {}

    # -- storage (submap.py:31-39, 135-152) -----------------------------------
    def add_all_poses(self, poses):
        self.poses = poses

    def add_all_points(self, points, colors, conf, conf_threshold_percentile, intrinsics):
        """Stores the arrays; the threshold is np.percentile(conf, pct) over the whole (S,H,W) volume,
        loop-closure frames included (submap.py:34-39), computed by vsm_conf_threshold."""
        self.pointclouds = points
        self.colors = colors
        self.conf = conf
        self._dev_cache.clear()
        self.conf_threshold = vm.conf_threshold(self._device("conf"), conf_threshold_percentile)
        self.vggt_intrinscs = intrinsics

    def add_all_semantic_embeddings(self, semantic_embeddings, embeddings_are_bf16_bits: bool = False):
        """(S,H,W,d) array aligned to the point maps.  Same checks and exception types as
        submap.py:41-65, except that torch tensors (CPU or CUDA, float32 or bfloat16) are accepted too."""
        if semantic_embeddings is None:
            self.semantic_embeddings = None
            return
        if not isinstance(semantic_embeddings, (np.ndarray, torch.Tensor)):
            raise TypeError("semantic_embeddings must be a numpy array of shape (S,H,W,d)")
        if semantic_embeddings.ndim != 4:
            raise ValueError(
                f"semantic_embeddings must have 4 dims (S,H,W,d), got shape={_shape(semantic_embeddings)}")
        if self.pointclouds is not None:
            if _shape(semantic_embeddings)[:3] != _shape(self.pointclouds)[:3]:
                raise ValueError("semantic_embeddings spatial dims must match pointclouds. "
                                 f"semantic={_shape(semantic_embeddings)[:3]} vs points={_shape(self.pointclouds)[:3]}")
        self.semantic_embeddings = semantic_embeddings
        self.semantic_index = None
        self.embeddings_are_bf16_bits = bool(embeddings_are_bf16_bits)
        self._dev_cache.pop("emb", None)

    def add_all_semantic_embeddings_indexed(self, mask_ids, table, embeddings_are_bf16_bits: bool = False):
        """Indexed form of ``add_all_semantic_embeddings`` (no upstream counterpart): ``mask_ids`` (S,H,W)
        integers and ``table`` (M,d) such that pixel (s,h,w) carries ``table[mask_ids[s,h,w]]``.  That is what the
        embedder paints into its dense image -- one CLIP vector per SAM mask, zeros where there is none
        (semantic_embedder.py:324-349): row 0 = zeros, row i+1 = the i-th mask's vector, later masks overwrite earlier
        ones.  The fused map equals that of the expanded array ``table[mask_ids]``; the submap holds 4 bytes per pixel
        instead of 2 KB.  Same exception types as the dense call (submap.py:52-63)."""
        if not isinstance(mask_ids, (np.ndarray, torch.Tensor)) or not isinstance(table, (np.ndarray, torch.Tensor)):
            raise TypeError("mask_ids must be an integer array of shape (S,H,W) and table an array of shape (M,d)")
        if mask_ids.ndim != 3:
            raise ValueError(f"mask_ids must have 3 dims (S,H,W), got shape={_shape(mask_ids)}")
        if table.ndim != 2 or _shape(table)[0] < 1:
            raise ValueError(f"table must have 2 dims (M,d) with M >= 1, got shape={_shape(table)}")
        is_int = (mask_ids.dtype in (torch.int16, torch.int32, torch.int64, torch.uint8)) if isinstance(mask_ids, torch.Tensor) \
            else np.issubdtype(mask_ids.dtype, np.integer)
        if not is_int:
            raise TypeError(f"mask_ids must be integers, got {mask_ids.dtype}")
        if self.pointclouds is not None and _shape(mask_ids) != _shape(self.pointclouds)[:3]:
            raise ValueError("mask_ids spatial dims must match pointclouds. "
                             f"semantic={_shape(mask_ids)} vs points={_shape(self.pointclouds)[:3]}")
        self.semantic_embeddings = table
        self.semantic_index = mask_ids
        self.embeddings_are_bf16_bits = bool(embeddings_are_bf16_bits)
        self._dev_cache.pop("emb", None)
        self._dev_cache.pop("emb_index", None)

    def index_on_device(self) -> Optional[torch.Tensor]:
        """int32 device tensor of the embedding indices (cached: 4 bytes per pixel), or None for dense embeddings."""
        if self.semantic_index is None:
            return None
        t = self._dev_cache.get("emb_index")
        if t is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            src = self.semantic_index
            src = torch.from_numpy(np.ascontiguousarray(src)) if isinstance(src, np.ndarray) else src
            t = src.to(dev).to(torch.int32).contiguous()
            self._dev_cache["emb_index"] = t
        return t

    def dense_semantic_embeddings(self):
        """The (S,H,W,d) array the reference would hold: table[mask_ids] for indexed embeddings (host numpy, float32)."""
        if self.semantic_index is None:
            return self.semantic_embeddings
        tab = self.semantic_embeddings
        tab = tab.float().cpu().numpy() if isinstance(tab, torch.Tensor) else np.asarray(tab, dtype=np.float32)
        ids = self.semantic_index
        ids = ids.cpu().numpy() if isinstance(ids, torch.Tensor) else np.asarray(ids)
        return tab[ids]

    def add_all_frames(self, frames):
        self.frames = frames

    def add_all_retrieval_vectors(self, retrieval_vectors):
        self.retrieval_vectors = retrieval_vectors

    def set_all_retrieval_vectors(self, retrieval_vectors):
        self.retrieval_vectors = retrieval_vectors

    def get_id(self):
        return self.submap_id

    def get_conf_threshold(self):
        return self.conf_threshold

    def get_frame_at_index(self, index):
        return self.frames[index, ...]

    def get_last_non_loop_frame_index(self):
        return self.last_non_loop_frame_index

    def get_all_frames(self):
        return self.frames

    def get_all_retrieval_vectors(self):
        return self.retrieval_vectors

    def get_all_poses_world(self, ignore_loop_closure_frames=False):
        """World camera poses of the frames (submap.py:91-104): the projection K @ inv(pose)[:3] @ inv(H_world_map) is
        decomposed (cv2.decomposeProjectionMatrix) into rotation and camera centre.  Host code of the SLAM back end,
        kept because GraphMap.save_frame_outputs writes its result next to the point maps."""
        import cv2

        poses_in = np.asarray(self.poses)
        K = np.asarray(self.vggt_intrinscs)
        projection_mat_list = K @ np.linalg.inv(poses_in)[:, 0:3, :] @ np.linalg.inv(np.asarray(self.H_world_map))
        poses = []
        for index, projection_mat in enumerate(projection_mat_list):
            cal, rot, trans = cv2.decomposeProjectionMatrix(projection_mat)[0:3]
            trans = trans / trans[3, 0]
            pose = np.eye(4)
            pose[0:3, 0:3] = np.linalg.inv(rot)
            pose[0:3, 3] = trans[0:3, 0]
            poses.append(pose)
            if ignore_loop_closure_frames and index == self.last_non_loop_frame_index:
                break
        return np.stack(poses, axis=0)

    def get_frame_pointcloud(self, pose_index):
        return self.pointclouds[pose_index]

    def set_frame_ids(self, file_paths):
        """First integer / decimal in each basename as float; str(float) keys the name map
        (submap.py:109-131).  Loop-closure frames are not part of file_paths."""
        ids, names, id_to_name = [], [], {}
        for path in file_paths:
            filename = os.path.basename(path)
            m = re.search(r"\d+(?:\.\d+)?", filename)
            if m is None:
                raise ValueError(f"No number found in image name: {filename}")
            fid = float(m.group())
            ids.append(fid)
            names.append(filename)
            id_to_name[str(fid)] = filename
        self.frame_ids = ids
        self.frame_names = names
        self.frame_id_to_name = id_to_name

    def set_last_non_loop_frame_index(self, last_non_loop_frame_index):
        self.last_non_loop_frame_index = last_non_loop_frame_index

    def set_reference_homography(self, H_world_map):
        self.H_world_map = H_world_map

    def set_conf_masks(self, conf_masks):
        self.conf_masks = conf_masks

    def get_reference_homography(self):
        return self.H_world_map

    def get_pose_subframe(self, pose_index):
        return np.linalg.inv(self.poses[pose_index])

    def get_frame_ids(self):
        return self.frame_ids

    # -- device views ----------------------------------------------------------
    def _device(self, what: str) -> torch.Tensor:
        """Device tensor of points / conf / colors / conf_masks (small arrays: cached)."""
        t = self._dev_cache.get(what)
        if t is None:
            vm.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device())
            src = {"points": self.pointclouds, "conf": self.conf, "colors": self.colors,
                   "conf_masks": self.conf_masks}[what]
            dt = torch.uint8 if what == "colors" else torch.float32
            t = vm.as_device(src, dev, dt)
            self._dev_cache[what] = t
        return t

    def embedding_dtype_code(self) -> int:
        e = self.semantic_embeddings
        if isinstance(e, torch.Tensor):
            if e.dtype == torch.bfloat16:
                return N.BF16
            if e.dtype == torch.float32:
                return N.F32
            raise TypeError(f"semantic embeddings must be float32 or bfloat16, got {e.dtype}")
        if e.dtype == np.uint16 and self.embeddings_are_bf16_bits:
            return N.BF16
        return N.F32

    def embeddings_on_device(self, cache: bool = False) -> torch.Tensor:
        """Embeddings as a device tensor in the dtype the kernels read (no conversion of values)."""
        t = self._dev_cache.get("emb")
        if t is not None:
            return t
        dev = torch.device("cuda", torch.cuda.current_device())
        e = self.semantic_embeddings
        if isinstance(e, torch.Tensor):
            t = e.to(dev).contiguous()
        elif e.dtype == np.uint16 and self.embeddings_are_bf16_bits:
            t = torch.from_numpy(np.ascontiguousarray(e).view(np.int16)).to(dev).view(torch.bfloat16)
        else:
            t = vm.as_device(e, dev, torch.float32)
        if cache:
            self._dev_cache["emb"] = t
        return t

    def prefetch_to_device(self, copy_stream: "torch.cuda.Stream", with_embeddings: bool) -> None:
        """Start the host->device copies of this submap's small arrays (points, confidences and -- for indexed
        embeddings -- the index image and the table) on ``copy_stream`` and make the CURRENT stream wait for them.
        Copies from pinned memory then overlap the fuse calls already queued on the current stream (the copies of
        submap i+1 run beside the kernels of submap i); pageable sources are copied synchronously, as before."""
        dev = torch.device("cuda", torch.cuda.current_device())
        main = torch.cuda.current_stream(dev)
        todo = [("points", self.pointclouds, torch.float32), ("conf", self.conf, torch.float32)]
        if with_embeddings and self.semantic_index is not None:
            todo += [("emb_index", self.semantic_index, torch.int32), ("emb", self.semantic_embeddings, None)]
        started = False
        with torch.cuda.stream(copy_stream):
            for key, src, dt in todo:
                if key in self._dev_cache or src is None:
                    continue
                if key == "emb" and isinstance(src, np.ndarray) and src.dtype == np.uint16 and self.embeddings_are_bf16_bits:
                    t = torch.from_numpy(np.ascontiguousarray(src).view(np.int16))
                else:
                    t = torch.from_numpy(np.ascontiguousarray(src)) if isinstance(src, np.ndarray) else src
                if t.is_cuda:
                    continue
                if dt is not None and t.dtype != dt:
                    t = t.to(dt)
                elif key == "emb" and t.dtype not in (torch.float32, torch.bfloat16, torch.int16):
                    t = t.to(torch.float32)
                d = t.to(dev, non_blocking=True).contiguous()
                if key == "emb" and d.dtype == torch.int16:
                    d = d.view(torch.bfloat16)
                d.record_stream(main)
                self._dev_cache[key] = d
                started = True
        if started:
            main.wait_stream(copy_stream)

    def release_device_cache(self) -> None:
        self._dev_cache.clear()

    # -- a4 / a5: world-frame extraction (submap.py:155-188, 217-219) ------------
    def filter_data_by_confidence(self, data, stride=1):
        """data[conf >= thr] on the [::stride, ::stride] grid.  Points and colours go through the device
        gather; other arrays use the same mask on the host."""
        if data is self.colors:
            _, c = vm.select_points(None, self._device("conf"), self._device("colors"), stride, self.conf_threshold,
                                    None, False, True)
            return c.cpu().numpy()
        if data is self.pointclouds:
            w, _ = vm.select_points(self._device("points"), self._device("conf"), None, stride, self.conf_threshold,
                                    np.eye(4), True, False)
            return w.cpu().numpy().astype(np.float32)
        conf = self.conf if isinstance(self.conf, np.ndarray) else self.conf.cpu().numpy()
        if stride == 1:
            return data[conf >= self.conf_threshold]
        return data[:, ::stride, ::stride, :][conf[:, ::stride, ::stride] >= self.conf_threshold]

    def get_points_in_world_frame(self, stride=1):
        """(N,3) float64 = (H @ [p;1]) / w of the confident points (submap.py:182-188)."""
        w, _ = vm.select_points(self._device("points"), self._device("conf"), None, stride, self.conf_threshold,
                                self.H_world_map, True, False)
        return w.cpu().numpy()

    def get_points_colors(self, stride=1):
        _, c = vm.select_points(None, self._device("conf"), self._device("colors"), stride, self.conf_threshold, None,
                                False, True)
        return c.cpu().numpy().reshape(-1, 3)

    def get_points_list_in_world_frame(self, ignore_loop_closure_frames=False):
        """Per-frame unfiltered (H,W,3) float64 maps, frame ids and per-frame masks (submap.py:166-180)."""
        S = _shape(self.pointclouds)[0]
        n = S
        if ignore_loop_closure_frames and self.last_non_loop_frame_index is not None:
            n = min(S, int(self.last_non_loop_frame_index) + 1)
        if self.frame_ids is not None and n > len(self.frame_ids):
            # the reference indexes frame_ids[index] inside its loop and raises here
            raise IndexError("list index out of range")
        world = vm.transform_points(self._device("points")[:n], self.H_world_map, out_f64=True).cpu().numpy()
        masks = (self._device("conf_masks")[:n] >= float(self.conf_threshold)).cpu().numpy()
        return [world[i] for i in range(n)], [self.frame_ids[i] for i in range(n)], [masks[i] for i in range(n)]

    # -- a6: per-submap semantic voxelisation (submap.py:221-311) ----------------
    def get_semantic_voxel_in_world_frame(self, voxel_size: float, stride: int = 1,
                                          ignore_loop_closure_frames: bool = False) -> SemanticVoxel:
        """Voxel-average the embeddings of this submap's confident points in the world frame.  No outlier
        filters; ``stride`` is accepted and ignored, like upstream (Appendix A-12); contributors hold one
        (submap_id, frame_id) tuple per point, in point order."""
        if voxel_size <= 0.0:
            raise ValueError("voxel_size must be > 0")
        if self.pointclouds is None:
            raise RuntimeError("No pointclouds in submap. Run add_all_points() first.")
        if self.semantic_embeddings is None:
            raise RuntimeError("No semantic embeddings in submap. Run add_all_semantic_embeddings() first.")
        if self.H_world_map is None:
            raise RuntimeError("No reference homography in submap. Run set_reference_homography() first.")
        S, H, W = _shape(self.pointclouds)[:3]
        d = _shape(self.semantic_embeddings)[-1]
        end_idx = S
        if ignore_loop_closure_frames and self.last_non_loop_frame_index is not None:
            end_idx = min(end_idx, int(self.last_non_loop_frame_index) + 1)
        dm = vm.DeviceVoxelMap(float(voxel_size), d, self.embedding_dtype_code(), capacity=1 << 16)
        index = self.index_on_device()
        params = dm.make_params(S, H, W, end_idx, 1, self.conf_threshold, self.H_world_map, int(self.submap_id),
                                N.FUSE_KEEP_POINT_INDEX, emb_index=index,
                                emb_rows=_shape(self.semantic_embeddings)[0] if index is not None else 0)
        stats = dm.fuse(self._device("points"), self._device("conf"), self.embeddings_on_device(), params)
        dm.finalize()
        V = dm.num_voxels
        if stats["n_fused"] == 0:
            return SemanticVoxel(voxel_size=voxel_size, centers_world=np.zeros((0, 3), dtype=np.float32),
                                 features=np.zeros((0, d), dtype=np.float32), contributors=[])
        _, centers, _, _ = dm.export_geometry(coords=False, centers=True, counts=False, recon=False)
        inverse = dm.export_point_index(0, S * H * W)
        contributors = _per_point_contributors(int(self.submap_id), self.frame_ids, inverse, H * W, V)
        vox = SemanticVoxel.lazy(voxel_size, centers.cpu().numpy(), dm.features_to_host, contributors)
        vox._device_map = dm
        return vox


def _per_point_contributors(submap_id: int, frame_ids, inverse: torch.Tensor, px_per_frame: int, V: int):
    """One (submap_id, frame_id) tuple per fused point, grouped by voxel, in point order
    (submap.py:295-304).  ``inverse`` is the device array pixel -> sorted voxel index (-1: not fused)."""
    state = {}

    def build():
        inv = inverse.cpu().numpy()
        pix = np.nonzero(inv >= 0)[0]
        vox = inv[pix]
        order = np.argsort(vox, kind="stable")
        frames = (pix[order] // px_per_frame).astype(np.int64)
        bounds = np.searchsorted(vox[order], np.arange(V + 1))
        n_ids = 0 if frame_ids is None else len(frame_ids)
        names = [str(frame_ids[f]) if f < n_ids else str(int(f)) for f in range(int(frames.max()) + 1)] if len(frames) else []
        state["frames"], state["bounds"], state["names"] = frames, bounds, names

    def maker(i):
        if not state:
            build()
        lo, hi = state["bounds"][i], state["bounds"][i + 1]
        names = state["names"]
        return [(submap_id, names[f]) for f in state["frames"][lo:hi].tolist()]

    return LazyContributors(V, maker)
