"""GraphMap with the reference's hot-path API (vggt_slam/map.py): the dict of submaps, the global
semantic voxel build and the point-cloud / per-frame dumps.  SLAM back-end methods (loop-closure
retrieval, pose files, COLMAP alignment) are out of scope.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _native as N
from . import voxel_map as vm
from .semantic_voxel import LazyContributors, SemanticVoxel, SemanticVoxelMap
from .submap import _shape


class GraphMap:
    def __init__(self):
        self.submaps = dict()
        self.last_build_stats: List[dict] = []
        self.last_profile = None

    def get_num_submaps(self):
        return len(self.submaps)

    def add_submap(self, submap):
        self.submaps[submap.get_id()] = submap

    def get_largest_key(self):
        return -1 if len(self.submaps) == 0 else max(self.submaps.keys())

    def get_submap(self, id):
        return self.submaps[id]

    def get_latest_submap(self):
        return self.get_submap(self.get_largest_key())

    def get_submaps(self):
        return self.submaps.values()

    def ordered_submaps_by_key(self):
        for k in sorted(self.submaps):
            yield self.submaps[k]

    def update_submap_homographies(self, graph):
        """H_world_map := optimised 4x4 of the pose graph (map.py:73-76)."""
        for key, submap in self.submaps.items():
            submap.set_reference_homography(graph.get_homography(key).matrix())

    def apply_similarity_transform(self, T_world_from_pred: np.ndarray) -> None:
        """H_world_map := T @ H_world_map for every submap, float64 (map.py:383-396)."""
        T = np.asarray(T_world_from_pred, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError(f"T_world_from_pred must be 4x4, got {T.shape}")
        for submap in self.ordered_submaps_by_key():
            H = submap.get_reference_homography()
            if H is None:
                continue
            submap.set_reference_homography((T @ H).astype(np.float64))

    # -- a9: point-cloud build (map.py:98-104, 106-151, 154-168) ------------------
    def get_points_and_colors(self):
        """Concatenated world-frame points (N,3) float64 and colours (N,3) in [0,1], as
        write_points_to_file assembles them before handing them to open3d (map.py:154-165)."""
        pts, cols = [], []
        for submap in self.ordered_submaps_by_key():
            pts.append(submap.get_points_in_world_frame().reshape(-1, 3))
            cols.append(submap.get_points_colors())
        pts = np.concatenate(pts, axis=0)
        cols = np.concatenate(cols, axis=0)
        if cols.max() > 1.0:
            cols = cols / 255.0
        return pts, cols

    def write_points_to_file(self, file_name):
        """Binary PCD (x y z rgb) written directly; the reference goes through open3d (map.py:166-168)."""
        pts, cols = self.get_points_and_colors()
        rgb = (np.clip(cols, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint32)
        packed = ((rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]).astype(np.uint32).view(np.float32)
        rec = np.empty(pts.shape[0], dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<f4")])
        rec["x"], rec["y"], rec["z"], rec["rgb"] = pts[:, 0], pts[:, 1], pts[:, 2], packed
        n = pts.shape[0]
        header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\n"
                  f"TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
        with open(file_name, "wb") as f:
            f.write(header.encode("ascii"))
            f.write(rec.tobytes())

    def save_framewise_pointclouds(self, file_name):
        os.makedirs(file_name, exist_ok=True)
        for submap in self.ordered_submaps_by_key():
            pointclouds, frame_ids, conf_masks = submap.get_points_list_in_world_frame(ignore_loop_closure_frames=True)
            for frame_id, pointcloud, mask in zip(frame_ids, pointclouds, conf_masks):
                np.savez(f"{file_name}/{frame_id}.npz", pointcloud=pointcloud, mask=mask)

    def save_frame_outputs(self, output_dir, ignore_loop_closure_frames=True):
        """Per-frame world point map, confidence mask, world extrinsic and intrinsic as {output_dir}/{stem}.npz
        (map.py:106-151): same keys, same file names (frame name stem, else str(frame id)), same skips (submaps
        without points or transform; a length mismatch between point maps and extrinsics).  The point maps come
        from vsm_transform_points; the extrinsics are Submap.get_all_poses_world's (host, SLAM back end)."""
        os.makedirs(output_dir, exist_ok=True)
        for submap in self.ordered_submaps_by_key():
            if submap.pointclouds is None or submap.H_world_map is None:
                continue
            end_idx = _shape(submap.pointclouds)[0]
            if ignore_loop_closure_frames and (submap.last_non_loop_frame_index is not None):
                end_idx = min(end_idx, submap.last_non_loop_frame_index + 1)
            pointclouds, frame_ids, conf_masks = submap.get_points_list_in_world_frame(
                ignore_loop_closure_frames=ignore_loop_closure_frames)
            extrinsics_world = submap.get_all_poses_world(ignore_loop_closure_frames=ignore_loop_closure_frames)
            intrinsics = submap.vggt_intrinscs
            if intrinsics is None:
                intrinsics = [None] * len(pointclouds)
            if len(pointclouds) != len(extrinsics_world):
                print(f"Skipping submap {submap.get_id()} due to length mismatch: "
                      f"{len(pointclouds)} point maps vs {len(extrinsics_world)} extrinsics.")
                continue
            frame_names = getattr(submap, "frame_names", None)
            for idx in range(min(end_idx, len(pointclouds))):
                if frame_names is not None and idx < len(frame_names):
                    stem, _ = os.path.splitext(str(frame_names[idx]))
                    filename = f"{stem}.npz"
                else:
                    frame_id = frame_ids[idx] if frame_ids is not None else idx
                    filename = f"{str(frame_id)}.npz"
                np.savez(os.path.join(output_dir, filename), point_map_world=pointclouds[idx],
                         conf_mask=conf_masks[idx], extrinsic_world=extrinsics_world[idx],
                         intrinsic=intrinsics[idx] if intrinsics is not None else None)

    # -- a7: global semantic voxel map (map.py:170-381) -----------------------------
    def build_semantic_voxel_map(self, voxel_size: float, stride: int = 1, ignore_loop_closure_frames: bool = True,
                                 deduplicate_contributors: bool = True, use_torch: bool = True,
                                 capacity_hint: Optional[int] = None, host_streaming: Optional[bool] = None,
                                 exact_coords: bool = False, profile: bool = False) -> SemanticVoxelMap:
        """Fuse every submap into one global voxel map on the GPU.

        Per submap: confidence mask, float64 world transform, the finite / percentile-box / coarse-cell
        filters, voxel keys, scatter-accumulate of the embeddings (vsm_fuse_submap with VSM_FUSE_FILTERS);
        then one finalisation.  ``use_torch`` is accepted for signature compatibility; there is one path.
        ``host_streaming``: True streams host embeddings frame by frame through pinned buffers
        (vsm_fuse_submap_host); default: used when the embeddings are host arrays.
        """
        dm, fused, frame_name_maps = self.fuse_into_device_map(voxel_size, stride, ignore_loop_closure_frames,
                                                               deduplicate_contributors, capacity_hint,
                                                               host_streaming, profile)
        if dm is None or dm.num_voxels == 0:
            return SemanticVoxelMap(_empty_voxels(voxel_size), frame_name_maps=frame_name_maps)
        dm.finalize()
        return wrap_device_map(dm, fused, frame_name_maps, voxel_size, deduplicate_contributors, exact_coords)

    def usable_submaps(self):
        """The submaps build_semantic_voxel_map fuses (map.py:196-203 skips the others), in key order."""
        return [sm for sm in self.ordered_submaps_by_key()
                if getattr(sm, "semantic_embeddings", None) is not None and sm.pointclouds is not None
                and sm.conf is not None and sm.conf_threshold is not None and sm.H_world_map is not None]

    def fuse_into_device_map(self, voxel_size, stride=1, ignore_loop_closure_frames=True,
                             deduplicate_contributors=True, capacity_hint=None, host_streaming=None, profile=False,
                             dm=None, submaps=None):
        """The per-submap loop of build_semantic_voxel_map: returns (DeviceVoxelMap or None, fused-call records,
        frame_name_maps) before finalisation (the multi-GPU build exchanges voxels at this point).
        ``submaps``: fuse only these (a round of the streaming multi-GPU build); ``dm``: fuse into this map."""
        if voxel_size <= 0.0:
            raise ValueError("voxel_size must be > 0")
        if stride < 1:
            raise ValueError("stride must be >= 1")
        vm.require_cuda()
        todo = []
        for submap in (self.ordered_submaps_by_key() if submaps is None else submaps):
            if getattr(submap, "semantic_embeddings", None) is None:
                continue
            if submap.pointclouds is None or submap.conf is None or submap.conf_threshold is None:
                continue
            if submap.H_world_map is None:
                continue
            todo.append(submap)
        frame_name_maps: Dict[str, Dict[str, str]] = {}
        self.last_build_stats = []
        self.last_profile = None
        if not todo:
            return dm, [], frame_name_maps
        d = _shape(todo[0].semantic_embeddings)[-1]
        code = todo[0].embedding_dtype_code()
        if dm is None:
            dm = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=capacity_hint or (1 << 18))
        if profile:
            dm.profile_enable(True)
        flags = N.FUSE_FILTERS | (0 if deduplicate_contributors else N.FUSE_KEEP_POINT_INDEX)
        fused: List[dict] = []
        try:
            self._fuse_all(dm, todo, stride, ignore_loop_closure_frames, flags, host_streaming, fused, frame_name_maps)
        except N.NonFiniteEmbeddingError:
            # a non-finite embedding row took part in the optimistic pass: redo exactly, with the row mask
            # computed before the percentiles (map.py:247)
            dm.clear()
            fused.clear()
            frame_name_maps.clear()
            self.last_build_stats = []
            if profile:
                dm.profile_enable(True)
            self._fuse_all(dm, todo, stride, ignore_loop_closure_frames, flags | N.FUSE_EMB_PRECHECK, False, fused,
                           frame_name_maps)
        if profile:
            self.last_profile = dm.profile()
        return dm, fused, frame_name_maps

    # Queued device-path calls keep their input tensors alive until they are collected.  For dense embeddings that
    # arrive as HOST arrays the tensor is a full device copy (5-10 GB per submap at the benchmark's shape): collect
    # once this many bytes of such copies are queued instead of letting up to 64 calls pile up.
    MAX_QUEUED_COPY_BYTES = 24 << 30

    def _fuse_all(self, dm, todo, stride, ignore_loop, flags, host_streaming, fused, frame_name_maps):
        queued = []
        queued_copy_bytes = 0

        def index_is_set(sm):
            return getattr(sm, "semantic_index", None) is not None

        def flush():
            waiting = [i for i, (st, _) in enumerate(queued) if st is None]
            if not waiting:
                return
            got = dm.collect()
            assert len(got) == len(waiting), (len(got), len(waiting))
            for i, st in zip(waiting, got):
                queued[i] = (st, queued[i][1])

        for submap in todo:
            S, H, W = _shape(submap.pointclouds)[:3]
            end_idx = S
            if ignore_loop and submap.last_non_loop_frame_index is not None:
                end_idx = min(end_idx, int(submap.last_non_loop_frame_index) + 1)
            if submap.embedding_dtype_code() != dm.emb_dtype or _shape(submap.semantic_embeddings)[-1] != dm.dim:
                raise ValueError("all submaps must carry embeddings of the same dtype and dimension")
            n_ids = 0 if submap.frame_ids is None else len(submap.frame_ids)
            if end_idx > n_ids:
                conf_dev = submap._device("conf")
                # upstream builds str(frame_ids[i]) for every kept point (map.py:239-240) and fails on
                # frames without an id (loop-closure frames)
                hs = conf_dev[n_ids:end_idx, ::stride, ::stride]
                if bool((hs >= float(submap.conf_threshold)).any()):
                    raise IndexError("list index out of range")
            sid = int(submap.get_id())
            emb = submap.semantic_embeddings
            on_host = isinstance(emb, np.ndarray) or (isinstance(emb, torch.Tensor) and not emb.is_cuda)
            stream_it = on_host if host_streaming is None else (host_streaming and on_host)
            if index_is_set(submap):
                stream_it = False  # an index image and its table are a few MB: copied whole, fused on the device path
            stream_it = stream_it and not (flags & N.FUSE_EMB_PRECHECK)
            if not stream_it:
                # device path: host arrays (geometry; index image + table) are copied on a side stream, beside the
                # kernels of the calls already queued
                submap.prefetch_to_device(_copy_stream(), with_embeddings=index_is_set(submap))
            index = submap.index_on_device() if index_is_set(submap) else None
            params = dm.make_params(S, H, W, end_idx, stride, submap.conf_threshold, submap.H_world_map, sid, flags,
                                    emb_index=index, emb_rows=_shape(emb)[0] if index is not None else 0)
            if stream_it:
                pts = submap.pointclouds if isinstance(submap.pointclouds, np.ndarray) else None
                if pts is None or not isinstance(submap.conf, np.ndarray):
                    pts_h = submap._device("points").cpu().numpy()
                    conf_h = submap._device("conf").cpu().numpy()
                else:
                    pts_h = np.ascontiguousarray(submap.pointclouds, dtype=np.float32)
                    conf_h = np.ascontiguousarray(submap.conf, dtype=np.float32)
                emb_h = emb if isinstance(emb, torch.Tensor) else np.ascontiguousarray(emb)
                if isinstance(emb_h, np.ndarray) and emb_h.dtype not in (np.float32, np.uint16):
                    emb_h = emb_h.astype(np.float32)
                flush()  # the host-streamed call synchronises: take the queued calls' stats first
                stats = dm.fuse_host(pts_h, conf_h, emb_h, params)
            else:
                # queue the call; all queued calls are collected with one synchronisation below
                emb_dev = submap.embeddings_on_device(cache=index is not None)
                if on_host and index is None:
                    nbytes = emb_dev.numel() * emb_dev.element_size()
                    try:
                        budget = min(self.MAX_QUEUED_COPY_BYTES, torch.cuda.mem_get_info()[0] // 2 + queued_copy_bytes)
                    except Exception:
                        budget = self.MAX_QUEUED_COPY_BYTES
                    if queued_copy_bytes and queued_copy_bytes + nbytes > budget:
                        flush()
                        queued_copy_bytes = 0
                    queued_copy_bytes += nbytes
                dm.fuse_async(submap._device("points"), submap._device("conf"), emb_dev, params, keep_alive=index)
                del emb_dev
                stats = None
            queued.append((stats, {"fuse_index": dm.fuse_calls - 1, "submap": submap, "S": S, "H": H, "W": W,
                                   "end_idx": end_idx, "sid": sid}))
        flush()
        for stats, rec in queued:
            stats = dict(stats, submap_id=rec["sid"])
            self.last_build_stats.append(stats)
            if stats["n_fused"] == 0:
                continue
            fused.append(rec)
            submap = rec["submap"]
            if getattr(submap, "frame_id_to_name", None) is not None:
                frame_name_maps[str(rec["sid"])] = dict(submap.frame_id_to_name)
        if self.last_build_stats:
            self.last_build_stats[-1]["n_map_voxels"] = dm.num_voxels


_COPY_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _copy_stream() -> "torch.cuda.Stream":
    """ONE side stream per device for host->device prefetches.  A fresh torch.cuda.Stream per build would make the
    caching allocator keep a separate pool of blocks per stream: every build then allocates its staging tensors from
    the driver again and the dead pools pile up (measured: builds of 40 ms turning into 300-1000 ms at random)."""
    dev = torch.cuda.current_device()
    s = _COPY_STREAMS.get(dev)
    if s is None:
        s = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


def wrap_device_map(dm, fused, frame_name_maps, voxel_size, deduplicate_contributors=True, exact_coords=False):
    """SemanticVoxelMap over a finalised DeviceVoxelMap: centres now, features / contributors lazily."""
    V = dm.num_voxels

    def centers():  # fetched on first read, like the features
        return dm.export_geometry(coords=False, centers=True, counts=False, recon=False)[1].cpu().numpy()

    if deduplicate_contributors:
        contributors = _dedup_contributors(dm, fused, V)
    else:
        contributors = _per_point_contributors_global(dm, fused, V)
    vox = SemanticVoxel.lazy(float(voxel_size), centers, dm.features_to_host, contributors)
    m = SemanticVoxelMap(vox, frame_name_maps=frame_name_maps, _device_map=dm, exact_coords=exact_coords)
    if deduplicate_contributors:
        # the device's contributor CSR + the frame ids of every submap: what the binary side-car stores instead of V lists
        m._contrib_csr = {"dm": dm, "frame_ids": {int(f["submap"].get_id()): list(f["submap"].frame_ids) for f in fused}}
    return m


def _empty_voxels(voxel_size) -> SemanticVoxel:
    return SemanticVoxel(voxel_size=float(voxel_size), centers_world=np.zeros((0, 3), dtype=np.float32),
                         features=np.zeros((0, 0), dtype=np.float32), contributors=[])


def _dedup_contributors(dm, fused, V: int) -> LazyContributors:
    """sorted(set((submap_id, frame_id_str))) per voxel (map.py:365-369) from the device's contributor CSR
    (one entry per (fuse call, voxel) with a frame bit mask)."""
    state = {}
    frame_ids = {int(f["submap"].get_id()): f["submap"].frame_ids for f in fused}

    def maker(i):
        if not state:
            state["off"], state["sub"], state["mask"] = dm.export_contributors()
        off, sub, mask = state["off"], state["sub"], state["mask"]
        out = set()
        for e in range(int(off[i]), int(off[i + 1])):
            sid = int(sub[e])
            ids = frame_ids[sid]
            for w in range(2):
                bits = int(mask[e, w])
                while bits:
                    b = (bits & -bits).bit_length() - 1
                    out.add((sid, str(ids[64 * w + b])))
                    bits &= bits - 1
        return sorted(out)

    return LazyContributors(V, maker)


def _per_point_contributors_global(dm, fused, V: int) -> LazyContributors:
    """deduplicate_contributors=False: one tuple per fused point in global point order -- submaps ascending,
    then (frame, row, column) (map.py:370-373)."""
    state = {}

    def build():
        vox_all, sid_all, name_all = [], [], []
        for f in fused:
            sm = f["submap"]
            inv = dm.export_point_index(f["fuse_index"], f["S"] * f["H"] * f["W"]).cpu().numpy()
            pix = np.nonzero(inv >= 0)[0]
            frames = pix // (f["H"] * f["W"])
            names = np.array([str(x) for x in sm.frame_ids], dtype=object)
            vox_all.append(inv[pix].astype(np.int64))
            sid_all.append(np.full(pix.shape, int(sm.get_id()), dtype=np.int64))
            name_all.append(names[frames])
        vox = np.concatenate(vox_all)
        order = np.argsort(vox, kind="stable")
        state["sid"] = np.concatenate(sid_all)[order]
        state["name"] = np.concatenate(name_all)[order]
        state["bounds"] = np.searchsorted(vox[order], np.arange(V + 1))

    def maker(i):
        if not state:
            build()
        lo, hi = state["bounds"][i], state["bounds"][i + 1]
        return [(int(s), str(n)) for s, n in zip(state["sid"][lo:hi].tolist(), state["name"][lo:hi].tolist())]

    return LazyContributors(V, maker)
