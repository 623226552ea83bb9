"""SemanticVoxel / SemanticVoxelMap with the reference's API (vggt_slam/semantic_voxel.py:12-165),
backed by a device-resident map: queries, position lookups and the index order come from libvsm
kernels; numpy arrays and Python contributor lists are materialised from the device lazily.

Behaviour kept from the reference, on purpose:
  * voxel features are NOT normalised by ``query_with_embedding`` (A-1); ``normalize=True`` is an opt-in;
  * ``_voxel_coords`` is the reference's lossy float32 reconstruction floor(centers/vs - 0.5) (A-4) and is what
    ``query_with_embedding`` returns and what ``get_index_at_position`` looks up; ``exact_coords=True`` switches
    the lookup to the true integer keys;
  * ``get_latest_frame_at_voxel`` orders frame ids as *strings* and sorts the stored list in place (A-6).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .voxel_map import DeviceVoxelMap, require_cuda


class LazyContributors(Sequence):
    """list[list[(submap_id, frame_id_str)]] produced on demand from the device's contributor tables.

    ``maker(i)`` builds the list of voxel i; built lists are cached so that in-place mutation by
    ``get_latest_frame_at_voxel`` persists, as it does for the reference's plain lists."""

    def __init__(self, n: int, maker):
        self._n = int(n)
        self._maker = maker
        self._cache: Dict[int, list] = {}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("voxel index out of range")
        got = self._cache.get(i)
        if got is None:
            got = self._maker(i)
            self._cache[i] = got
        return got

    def __iter__(self):
        for i in range(self._n):
            yield self[i]

    def __eq__(self, other):
        if isinstance(other, (list, LazyContributors)):
            return len(other) == self._n and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def tolist(self) -> list:
        return [self[i] for i in range(self._n)]


class SemanticVoxel:
    """Same fields as the reference dataclass (vggt_slam/semantic_voxel.py:12-26): voxel_size,
    centers_world (N,3), features (N,d), contributors (length-N list of lists of (submap_id, frame_id)).

    When built by the device path, ``centers_world``, ``features`` and ``contributors`` are fetched from the GPU
    the first time they are read (``centers_world`` may be passed as a function returning the array)."""

    def __init__(self, voxel_size, centers_world, features, contributors):
        self.voxel_size = voxel_size
        self._centers = None if callable(centers_world) else centers_world
        self._centers_fn = centers_world if callable(centers_world) else None
        self._features = features
        self._features_fn = None
        self.contributors = contributors

    @classmethod
    def lazy(cls, voxel_size, centers_world, features_fn, contributors):
        v = cls(voxel_size, centers_world, None, contributors)
        v._features_fn = features_fn
        return v

    @property
    def centers_world(self):
        if self._centers is None and self._centers_fn is not None:
            self._centers = self._centers_fn()
        return self._centers

    @centers_world.setter
    def centers_world(self, value):
        self._centers, self._centers_fn = value, None

    def __len__(self):
        if self._centers is None and self._centers_fn is not None:
            return len(self.contributors)
        return 0 if self._centers is None else int(np.asarray(self._centers).shape[0])

    @property
    def features(self):
        if self._features is None and self._features_fn is not None:
            self._features = self._features_fn()
        return self._features

    @features.setter
    def features(self, value):
        self._features = value

    def __repr__(self):
        return f"SemanticVoxel(voxel_size={self.voxel_size}, n_voxels={len(self)})"


class SemanticVoxelMap:
    """Semantic voxel map: saving, loading, querying (vggt_slam/semantic_voxel.py:29-165)."""

    def __init__(self, voxels: SemanticVoxel, frame_name_maps: Dict[str, Dict[str, str]], _device_map=None,
                 exact_coords: bool = False):
        self.voxels = voxels
        self.voxel_size = float(voxels.voxel_size)
        self.frame_name_maps = frame_name_maps
        self.exact_coords = bool(exact_coords)
        self._dm: Optional[DeviceVoxelMap] = _device_map
        self._coord_dict = None
        n = len(voxels) if isinstance(voxels, SemanticVoxel) else int(np.asarray(voxels.centers_world).shape[0])
        if self._dm is None and n > 0:
            # a map made from host arrays (load_from_directory, or user code): upload once
            require_cuda()
            feats = np.ascontiguousarray(np.asarray(voxels.features), dtype=np.float32)
            d = int(feats.shape[1])
            if d % 8 != 0:
                raise ValueError("feature dimension must be a multiple of 8")
            self._dm = DeviceVoxelMap(self.voxel_size, d, N.F32, capacity=max(n, 1024))
            self._dm.load_dense(np.ascontiguousarray(voxels.centers_world, dtype=np.float32), feats)
        self._recon_coords = None  # fetched from the device on first use (6 MB of int64 at 250 k voxels)

    @property
    def _voxel_coords(self) -> np.ndarray:
        """The reference's reconstructed coordinates floor(centers/vs - 0.5) (semantic_voxel.py:38, 62-66)."""
        if self._recon_coords is None:
            if self._dm is not None and self._dm.num_voxels > 0:
                _, _, _, recon = self._dm.export_geometry(coords=False, centers=False, counts=False, recon=True)
                self._recon_coords = recon.cpu().numpy()
            else:
                self._recon_coords = np.zeros((0, 3), dtype=np.int64)
        return self._recon_coords

    # -- getters (semantic_voxel.py:43-56) -----------------------------------
    def get_voxels(self) -> SemanticVoxel:
        return self.voxels

    def get_voxel_size(self) -> float:
        return self.voxel_size

    def get_centers_world(self) -> np.ndarray:
        return self.voxels.centers_world

    def get_features(self) -> np.ndarray:
        return self.voxels.features

    def get_contributors(self):
        return self.voxels.contributors

    def resolve_contributor(self, submap_id: int, frame_id: str) -> Optional[str]:
        return self.frame_name_maps[str(submap_id)][str(frame_id)]

    # -- coordinates -----------------------------------------------------------
    @staticmethod
    def _centers_to_voxel_coords(centers_world: np.ndarray, voxel_size: float) -> np.ndarray:
        """Host restatement kept for API compatibility (semantic_voxel.py:62-66); the map itself takes the
        same values from the device (vsm_export_geometry's recon_coords)."""
        return np.floor(centers_world / voxel_size - 0.5).astype(np.int64)

    @staticmethod
    def _position_to_voxel_coord(position_world: np.ndarray, voxel_size: float) -> Tuple[int, int, int]:
        p = np.asarray(position_world, dtype=np.float32).reshape(3)
        c = np.floor(p / voxel_size).astype(np.int64)
        return int(c[0]), int(c[1]), int(c[2])

    @property
    def _coord_to_index(self) -> Dict[Tuple[int, int, int], int]:
        """The reference's dict; built only if somebody asks for it (O(V) Python)."""
        if self._coord_dict is None:
            self._coord_dict = {(int(c[0]), int(c[1]), int(c[2])): i for i, c in enumerate(self._voxel_coords)}
        return self._coord_dict

    def get_indices_at_positions(self, positions_world: np.ndarray) -> np.ndarray:
        """Batched get_index_at_position: (M,3) -> (M,) int64, -1 where no voxel exists."""
        if self._dm is None:
            return np.full((np.asarray(positions_world).reshape(-1, 3).shape[0],), -1, dtype=np.int64)
        return self._dm.lookup(positions_world, compat=not self.exact_coords)

    def get_index_at_position(self, position_world: np.ndarray) -> Optional[int]:
        idx = int(self.get_indices_at_positions(np.asarray(position_world, dtype=np.float32).reshape(1, 3))[0])
        return None if idx < 0 else idx

    def get_features_at_position(self, position_world: np.ndarray) -> Optional[np.ndarray]:
        idx = self.get_index_at_position(position_world)
        if idx is None:
            return None
        return self.voxels.features[idx]

    def get_voxel_coord_at_index(self, index: int):
        return self._voxel_coords[index]

    def get_contributors_at_position(self, position_world: np.ndarray):
        idx = self.get_index_at_position(position_world)
        if idx is None:
            return None
        return self.voxels.contributors[idx]

    # -- query (semantic_voxel.py:97-116) --------------------------------------
    def query_with_embeddings(self, qe: np.ndarray, top_k: int = 1, normalize: bool = False, engine: int = 0):
        """Batched query: qe (P,d) -> (indices (P,k) int64, coords (P,k,3) int64, scores (P,k) float32), numpy."""
        if self._dm is None or self._dm.num_voxels == 0:
            raise RuntimeError("selected index k out of range")  # torch.topk on an empty map (Appendix A-13)
        q = np.asarray(qe, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if int(top_k) > N.MAX_TOPK:
            raise ValueError(f"top_k={top_k} exceeds the library limit VSM_MAX_TOPK={N.MAX_TOPK}")
        # vsm_query scores at most VSM_MAX_PROMPTS prompts per call: larger batches go block by block (the
        # reference loops one prompt at a time and has no limit, voxel_evaluators.py:61)
        idx_blocks, sc_blocks = [], []
        for p0 in range(0, max(q.shape[0], 1), N.MAX_PROMPTS):
            try:
                idx, sc = self._dm.query(q[p0:p0 + N.MAX_PROMPTS], top_k=top_k, normalize=normalize, engine=engine)
            except ValueError as e:
                if "exceeds" in str(e):
                    raise RuntimeError("selected index k out of range") from e
                raise
            idx_blocks.append(idx.cpu().numpy())
            sc_blocks.append(sc.cpu().numpy())
        idx = np.concatenate(idx_blocks, axis=0)
        return idx, self._voxel_coords[idx], np.concatenate(sc_blocks, axis=0)

    def query_with_embedding(self, qe: np.ndarray, top_k: int = 1, normalize: bool = False):
        """One prompt, the reference's return types: (list[int], (k,3) int64 array, list[float])."""
        q = np.asarray(qe, dtype=np.float32)
        if q.ndim == 2 and q.shape[0] != 1:
            raise ValueError("query_with_embedding scores one prompt; use query_with_embeddings for (P,d) batches")
        idx, coords, sc = self.query_with_embeddings(q.reshape(1, -1), top_k=top_k, normalize=normalize)
        return idx[0].tolist(), coords[0], [float(s) for s in sc[0]]

    def get_latest_frame_at_voxel(self, voxel_index: int):
        voxel_contributors = self.voxels.contributors[voxel_index]
        voxel_contributors.sort(key=lambda x: (x[0], x[1]), reverse=True)
        submap_id, frame_id = voxel_contributors[0]
        return self.resolve_contributor(submap_id, frame_id), submap_id, frame_id

    # -- persistence (semantic_voxel.py:128-165) ---------------------------------------------------
    # Two layouts in one directory:
    #   semantic_voxels.npz + frame_names.json   the reference's files, byte-compatible (its loader reads ours, ours reads
    #                                            its): a pickled object array of V Python lists -- fine for small maps
    #   sidecar/                                 plain .npy arrays, memory-mappable, written and read in row blocks
    #                                            straight from / to the device: meta.json, centers_world.npy (V,3) f32,
    #                                            features.npy (V,d) f32, contributors as a CSR -- contrib_offsets.npy
    #                                            (V+1) i64 + either contrib_submap.npy (M) i32 / contrib_mask.npy (M,2)
    #                                            u64 / frame_ids.json (device-built maps: one entry per submap and voxel,
    #                                            bit f = frame f) or contrib_pair.npy (M) i32 / pairs.json (any list)
    NPZ_MAX_VOXELS = 2_000_000
    SIDECAR = "sidecar"

    def save_to_directory(self, directory_path: str, sidecar: Optional[bool] = None, npz: Optional[bool] = None,
                          chunk_rows: int = 1 << 18) -> None:
        """``npz``: write the reference's file (default: when the map has at most NPZ_MAX_VOXELS voxels -- the pickled
        contributor lists do not scale); ``sidecar``: write the binary side-car (default: when the npz is skipped, or
        the map has more than 100 k voxels)."""
        os.makedirs(directory_path, exist_ok=True)
        V = len(self.voxels.contributors) if self.voxels.contributors is not None else 0
        npz = (V <= self.NPZ_MAX_VOXELS) if npz is None else bool(npz)
        sidecar = ((not npz) or V > 100_000) if sidecar is None else bool(sidecar)
        if npz:
            contribs = self.voxels.contributors
            contribs = contribs.tolist() if isinstance(contribs, LazyContributors) else contribs
            np.savez_compressed(
                os.path.join(directory_path, "semantic_voxels.npz"),
                voxel_size=np.float32(self.voxel_size),
                centers_world=np.asarray(self.voxels.centers_world).astype(np.float32),
                features=np.asarray(self.voxels.features).astype(np.float32),
                contributors=np.array(contribs, dtype=object),
            )
        with open(os.path.join(directory_path, "frame_names.json"), "w") as f:
            json.dump(self.frame_name_maps, f, indent=2)
        if sidecar:
            self._save_sidecar(os.path.join(directory_path, self.SIDECAR), chunk_rows)

    def _save_sidecar(self, path: str, chunk_rows: int) -> None:
        import torch
        from numpy.lib.format import write_array_header_1_0

        os.makedirs(path, exist_ok=True)
        dm = self._dm
        V = 0 if dm is None else dm.num_voxels
        d = 0 if dm is None else dm.dim
        np.save(os.path.join(path, "centers_world.npy"), np.asarray(self.voxels.centers_world, dtype=np.float32).reshape(V, 3))
        # features.npy: header, then the rows block by block through two pinned buffers (the device->host copy of
        # block i+1 runs while block i is written): the (V,d) matrix is never resident twice, on the device or on the host
        with open(os.path.join(path, "features.npy"), "wb") as fp:
            write_array_header_1_0(fp, {"descr": "<f4", "fortran_order": False, "shape": (int(V), int(d))})
            if V:
                n_blk = min(chunk_rows, V)
                stage = [torch.empty((n_blk, d), dtype=torch.float32, pin_memory=True) for _ in range(2)]
                buf = [torch.empty((n_blk, d), dtype=torch.float32, device=dm.device) for _ in range(2)]
                done = [torch.cuda.Event(), torch.cuda.Event()]
                stream = torch.cuda.current_stream(dm.device)
                blocks = [(r0, min(V, r0 + chunk_rows)) for r0 in range(0, V, chunk_rows)]

                def issue(i):
                    r0, r1 = blocks[i]
                    dm.export_features(r0, r1, buf[i & 1][: r1 - r0])
                    stage[i & 1][: r1 - r0].copy_(buf[i & 1][: r1 - r0], non_blocking=True)
                    done[i & 1].record(stream)

                issue(0)
                for i, (r0, r1) in enumerate(blocks):
                    if i + 1 < len(blocks):
                        issue(i + 1)
                    done[i & 1].synchronize()
                    fp.write(memoryview(stage[i & 1][: r1 - r0].numpy()).cast("B"))
        meta = {"version": 1, "voxel_size": float(self.voxel_size), "dim": int(d), "num_voxels": int(V)}
        src = getattr(self, "_contrib_csr", None)
        if src is not None:
            off, sub, mask = src["dm"].export_contributors()
            np.save(os.path.join(path, "contrib_offsets.npy"), off.astype(np.int64))
            np.save(os.path.join(path, "contrib_submap.npy"), sub.astype(np.int32))
            np.save(os.path.join(path, "contrib_mask.npy"), mask.astype(np.uint64))
            with open(os.path.join(path, "frame_ids.json"), "w") as f:
                json.dump({str(k): [str(x) for x in v] for k, v in src["frame_ids"].items()}, f)
            meta["contributors"] = "masks"
        else:
            contribs = self.voxels.contributors
            pairs, index, off, flat = [], {}, np.zeros(V + 1, dtype=np.int64), []
            for i in range(V):
                for sid, fid in contribs[i]:
                    key = (int(sid), str(fid))
                    j = index.get(key)
                    if j is None:
                        j = index[key] = len(pairs)
                        pairs.append([key[0], key[1]])
                    flat.append(j)
                off[i + 1] = len(flat)
            np.save(os.path.join(path, "contrib_offsets.npy"), off)
            np.save(os.path.join(path, "contrib_pair.npy"), np.asarray(flat, dtype=np.int32))
            with open(os.path.join(path, "pairs.json"), "w") as f:
                json.dump(pairs, f)
            meta["contributors"] = "pairs"
        with open(os.path.join(path, "meta.json"), "w") as f:  # written last: its presence marks a complete side-car
            json.dump(meta, f)

    @classmethod
    def _load_sidecar(cls, path: str, frame_name_maps, chunk_rows: int = 1 << 18) -> "SemanticVoxelMap":
        import torch

        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        V, d, vs = int(meta["num_voxels"]), int(meta["dim"]), float(meta["voxel_size"])
        centers = np.load(os.path.join(path, "centers_world.npy"), mmap_mode="r")
        feats = np.load(os.path.join(path, "features.npy"), mmap_mode="r")
        off = np.load(os.path.join(path, "contrib_offsets.npy"), mmap_mode="r")
        if meta.get("contributors") == "masks":
            sub = np.load(os.path.join(path, "contrib_submap.npy"), mmap_mode="r")
            mask = np.load(os.path.join(path, "contrib_mask.npy"), mmap_mode="r")
            with open(os.path.join(path, "frame_ids.json")) as f:
                frame_ids = {int(k): v for k, v in json.load(f).items()}

            def maker(i):
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
        else:
            pair = np.load(os.path.join(path, "contrib_pair.npy"), mmap_mode="r")
            with open(os.path.join(path, "pairs.json")) as f:
                pairs = [(int(a), str(b)) for a, b in json.load(f)]

            def maker(i):
                return [pairs[int(j)] for j in pair[int(off[i]):int(off[i + 1])]]

        dm = None
        if V:
            require_cuda()
            dm = DeviceVoxelMap(vs, d, N.F32, capacity=max(V, 1024))
            N.check(N.lib.vsm_map_load_begin(dm._h, V, C.c_void_p(torch.cuda.current_stream(dm.device).cuda_stream)))
            for r0 in range(0, V, chunk_rows):
                r1 = min(V, r0 + chunk_rows)
                c = torch.from_numpy(np.ascontiguousarray(centers[r0:r1], dtype=np.float32)).to(dm.device)
                f = torch.from_numpy(np.ascontiguousarray(feats[r0:r1], dtype=np.float32)).to(dm.device)
                N.check(N.lib.vsm_map_load_rows(dm._h, r0, r1 - r0, C.c_void_p(c.data_ptr()), C.c_void_p(f.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream(dm.device).cuda_stream)))
                torch.cuda.current_stream(dm.device).synchronize()
            dm.finalize()
        vox = SemanticVoxel.lazy(vs, lambda: np.asarray(centers), lambda: np.asarray(feats), LazyContributors(V, maker))
        return cls(vox, frame_name_maps=frame_name_maps, _device_map=dm)

    @staticmethod
    def load_from_directory(directory_path: str, prefer_sidecar: bool = True) -> "SemanticVoxelMap":
        json_path = os.path.join(directory_path, "frame_names.json")
        frame_name_maps: Dict[str, Dict[str, str]] = {}
        if os.path.exists(json_path):
            with open(json_path, "r") as f:
                frame_name_maps = json.load(f)
        side = os.path.join(directory_path, SemanticVoxelMap.SIDECAR)
        if prefer_sidecar and os.path.exists(os.path.join(side, "meta.json")):
            return SemanticVoxelMap._load_sidecar(side, frame_name_maps)
        data = np.load(os.path.join(directory_path, "semantic_voxels.npz"), allow_pickle=True)
        vox = SemanticVoxel(
            voxel_size=float(data["voxel_size"]),
            centers_world=data["centers_world"],
            features=data["features"],
            contributors=list(data["contributors"].tolist()),
        )
        return SemanticVoxelMap(vox, frame_name_maps=frame_name_maps)
