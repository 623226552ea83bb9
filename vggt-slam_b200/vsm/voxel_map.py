"""DeviceVoxelMap: the device-resident voxel map behind the reference-facing classes.

A thin object over the C ABI (include/vsm.h).  torch is used for device memory
and streams only; every computation is a libvsm kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as N


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("vsm needs a CUDA device (sm_100a): there is no CPU fallback for the voxel-mapping path")


def as_device(x, device: torch.device, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """numpy array or torch tensor -> contiguous tensor on `device` (no copy if it already is one)."""
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        t = t.to(device, non_blocking=False)
    return t.contiguous()


def emb_dtype_code(t) -> int:
    dt = t.dtype
    if dt in (torch.bfloat16,):
        return N.BF16
    if dt in (torch.float32, np.float32) or dt == np.dtype("float32"):
        return N.F32
    raise TypeError(f"semantic embeddings must be float32 or bfloat16, got {dt}")


class DeviceVoxelMap:
    """Owns one ``vsm_map`` handle."""

    def __init__(self, voxel_size: float, dim: int, emb_dtype: int = N.F32, capacity: int = 1 << 16,
                 device: Optional[torch.device] = None):
        require_cuda()
        if not float(voxel_size) > 0.0:
            raise ValueError("voxel_size must be > 0")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.voxel_size = float(voxel_size)
        self.dim = int(dim)
        self.emb_dtype = int(emb_dtype)
        cfg = N.Config(float(voxel_size), int(dim), int(emb_dtype), int(capacity), int(self.device.index or 0), 0)
        h = C.c_void_p()
        N.check(N.lib.vsm_map_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.fuse_calls = 0
        self._inflight = []

    # -- life cycle -------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            N.lib.vsm_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self) -> None:
        N.check(N.lib.vsm_map_clear(self._h, _stream_ptr(self.device)))
        self.fuse_calls = 0

    def reserve(self, n_voxels: int) -> None:
        N.check(N.lib.vsm_map_reserve(self._h, int(n_voxels), _stream_ptr(self.device)))

    def reserve_log(self, n_entries: int) -> None:
        N.check(N.lib.vsm_map_reserve_log(self._h, int(n_entries), _stream_ptr(self.device)))

    def clear_async(self) -> None:
        """vsm_map_clear_async on the CURRENT stream: queued behind what still reads the map there, no host wait."""
        N.check(N.lib.vsm_map_clear_async(self._h, _stream_ptr(self.device)))
        self.fuse_calls = 0

    @property
    def num_voxels(self) -> int:
        out = C.c_int64()
        N.check(N.lib.vsm_num_voxels(self._h, C.byref(out)))
        return int(out.value)

    @property
    def num_log_entries(self) -> int:
        """Contributor-log entries held (one per fuse call and voxel): what an exchange push sends besides the rows."""
        out = C.c_int64()
        N.check(N.lib.vsm_num_log_entries(self._h, C.byref(out)))
        return int(out.value)

    # -- fusion -----------------------------------------------------------
    def make_params(self, S, H, W, end_idx, stride, conf_threshold, H_world_map, submap_id, flags,
                    frame_base: int = 0, emb_index: Optional[torch.Tensor] = None, emb_rows: int = 0) -> N.FuseParams:
        """emb_index: int32 device tensor (S,H,W) of table rows -- the `emb` of the fuse call is then the (emb_rows, d)
        table (indexed embeddings).  The caller keeps the tensor alive until the call has been collected."""
        p = N.FuseParams()
        p.S, p.H, p.W, p.end_idx, p.stride = int(S), int(H), int(W), int(end_idx), int(stride)
        p.conf_threshold = float(conf_threshold)
        Hm = np.asarray(H_world_map, dtype=np.float64).reshape(16)
        for i in range(16):
            p.H_world_map[i] = float(Hm[i])
        p.submap_id = int(submap_id)
        p.flags = int(flags)
        p.bbox_lo_pct, p.bbox_hi_pct, p.coarse_factor, p.coarse_min_points = 0.5, 99.5, 3.0, 10
        p.frame_base = int(frame_base)
        if emb_index is not None:
            assert emb_index.is_cuda and emb_index.dtype == torch.int32 and emb_index.is_contiguous()
            assert tuple(emb_index.shape) == (int(S), int(H), int(W)), (tuple(emb_index.shape), (S, H, W))
            p.emb_index = emb_index.data_ptr()
            p.emb_rows = int(emb_rows)
        return p

    def fuse(self, points: torch.Tensor, conf: torch.Tensor, emb: torch.Tensor, params: N.FuseParams,
             emb_ok: Optional[torch.Tensor] = None) -> dict:
        """vsm_fuse_submap on device tensors: points (S,H,W,3) f32, conf (S,H,W) f32, emb (S,H,W,d)."""
        assert points.is_cuda and conf.is_cuda and emb.is_cuda
        assert points.dtype == torch.float32 and conf.dtype == torch.float32
        assert points.is_contiguous() and conf.is_contiguous() and emb.is_contiguous()
        assert emb_dtype_code(emb) == self.emb_dtype and emb.shape[-1] == self.dim
        assert not params.emb_index or (emb.ndim == 2 and emb.shape[0] == params.emb_rows)
        st = N.FuseStats()
        rc = N.lib.vsm_fuse_submap(self._h, _ptr(points), _ptr(conf), _ptr(emb), _ptr(emb_ok), C.byref(params),
                                   C.byref(st), _stream_ptr(self.device))
        self.last_stats = st.as_dict()
        N.check(rc)
        self.fuse_calls += 1
        return self.last_stats

    def fuse_async(self, points: torch.Tensor, conf: torch.Tensor, emb: torch.Tensor, params: N.FuseParams,
                   emb_ok: Optional[torch.Tensor] = None, keep_alive=None) -> None:
        """vsm_fuse_submap_async: queue the call on the current stream and return; `collect` gets the stats.
        The tensors are kept alive here until then."""
        assert points.is_cuda and conf.is_cuda and emb.is_cuda
        assert points.dtype == torch.float32 and conf.dtype == torch.float32
        assert points.is_contiguous() and conf.is_contiguous() and emb.is_contiguous()
        assert emb_dtype_code(emb) == self.emb_dtype and emb.shape[-1] == self.dim
        assert not params.emb_index or (emb.ndim == 2 and emb.shape[0] == params.emb_rows)
        N.check(N.lib.vsm_fuse_submap_async(self._h, _ptr(points), _ptr(conf), _ptr(emb), _ptr(emb_ok),
                                            C.byref(params), _stream_ptr(self.device)))
        self._inflight.append((points, conf, emb, emb_ok, keep_alive))
        self.fuse_calls += 1

    def collect(self) -> list:
        """vsm_fuse_collect: one synchronisation for all queued calls; list of stats dicts in call order."""
        n_max = max(len(self._inflight), 1) + 64
        arr = (N.FuseStats * n_max)()
        n = C.c_int32(0)
        rc = N.lib.vsm_fuse_collect(self._h, arr, n_max, C.byref(n), _stream_ptr(self.device))
        self._inflight.clear()
        out = [arr[i].as_dict() for i in range(min(int(n.value), n_max))]
        if out:
            self.last_stats = out[-1]
        N.check(rc)
        return out

    def fuse_host(self, points: np.ndarray, conf: np.ndarray, emb, params: N.FuseParams) -> dict:
        """vsm_fuse_submap_host on HOST arrays (numpy, or CPU torch tensors for bf16)."""
        pts_p = points.ctypes.data if isinstance(points, np.ndarray) else points.data_ptr()
        conf_p = conf.ctypes.data if isinstance(conf, np.ndarray) else conf.data_ptr()
        emb_p = emb.ctypes.data if isinstance(emb, np.ndarray) else emb.data_ptr()
        st = N.FuseStats()
        rc = N.lib.vsm_fuse_submap_host(self._h, C.c_void_p(pts_p), C.c_void_p(conf_p), C.c_void_p(emb_p),
                                        C.byref(params), C.byref(st), _stream_ptr(self.device))
        self.last_stats = st.as_dict()
        N.check(rc)
        self.fuse_calls += 1
        return self.last_stats

    def embedding_row_mask(self, conf: torch.Tensor, emb: torch.Tensor, params: N.FuseParams) -> torch.Tensor:
        out = torch.zeros(params.S * params.H * params.W, dtype=torch.uint8, device=self.device)
        N.check(N.lib.vsm_embedding_row_mask(self._h, _ptr(conf), _ptr(emb), C.byref(params), _ptr(out),
                                             _stream_ptr(self.device)))
        return out

    def profile_enable(self, on: bool = True) -> None:
        N.check(N.lib.vsm_profile_enable(self._h, int(bool(on))))

    def profile(self) -> dict:
        p = N.Profile()
        N.check(N.lib.vsm_profile_get(self._h, C.byref(p)))
        return {k: getattr(p, k) for k, _ in N.Profile._fields_}

    # -- finalisation / export ---------------------------------------------
    def finalize(self) -> None:
        N.check(N.lib.vsm_finalize(self._h, _stream_ptr(self.device)))

    def export_geometry(self, coords=True, centers=True, counts=True, recon=True):
        V = self.num_voxels
        dev = self.device
        t_coords = torch.empty((V, 3), dtype=torch.int64, device=dev) if coords else None
        t_centers = torch.empty((V, 3), dtype=torch.float32, device=dev) if centers else None
        t_counts = torch.empty((V,), dtype=torch.int64, device=dev) if counts else None
        t_recon = torch.empty((V, 3), dtype=torch.int64, device=dev) if recon else None
        N.check(N.lib.vsm_export_geometry(self._h, _ptr(t_coords), _ptr(t_centers), _ptr(t_counts), _ptr(t_recon),
                                          _stream_ptr(dev)))
        return t_coords, t_centers, t_counts, t_recon

    def export_features(self, r0: int = 0, r1: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        V = self.num_voxels
        r1 = V if r1 is None else r1
        if out is None:
            out = torch.empty((r1 - r0, self.dim), dtype=torch.float32, device=self.device)
        N.check(N.lib.vsm_export_features(self._h, int(r0), int(r1), _ptr(out), _stream_ptr(self.device)))
        return out

    def features_to_host(self, chunk_rows: int = 1 << 18) -> np.ndarray:
        """(V,d) float32 on the host, exported through a bounded device buffer."""
        V = self.num_voxels
        host = torch.empty((V, self.dim), dtype=torch.float32, pin_memory=V > 0)
        buf = None
        for r0 in range(0, V, chunk_rows):
            r1 = min(V, r0 + chunk_rows)
            if buf is None or buf.shape[0] < r1 - r0:
                buf = torch.empty((r1 - r0, self.dim), dtype=torch.float32, device=self.device)
            self.export_features(r0, r1, buf[: r1 - r0])
            host[r0:r1].copy_(buf[: r1 - r0], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def export_contributors(self):
        """CSR in sorted voxel order: offsets (V+1,) int64, submap ids (M,) int32, frame masks (M,2) uint64 (as int64)."""
        V = self.num_voxels
        M = C.c_int64()
        N.check(N.lib.vsm_num_contributor_entries(self._h, C.byref(M)))
        M = int(M.value)
        dev = self.device
        off = torch.empty((V + 1,), dtype=torch.int64, device=dev)
        sub = torch.empty((M,), dtype=torch.int32, device=dev)
        mask = torch.empty((M, 2), dtype=torch.int64, device=dev)
        N.check(N.lib.vsm_export_contributors(self._h, _ptr(off), _ptr(sub), _ptr(mask), _stream_ptr(dev)))
        return off.cpu().numpy(), sub.cpu().numpy(), mask.cpu().numpy().view(np.uint64)

    def export_point_index(self, fuse_index: int, n_pixels: int) -> torch.Tensor:
        out = torch.empty((n_pixels,), dtype=torch.int32, device=self.device)
        N.check(N.lib.vsm_export_point_index(self._h, int(fuse_index), _ptr(out), int(n_pixels),
                                             _stream_ptr(self.device)))
        return out

    def export_packed_keys(self) -> torch.Tensor:
        out = torch.empty((self.num_voxels,), dtype=torch.int64, device=self.device)
        N.check(N.lib.vsm_export_packed_keys(self._h, _ptr(out), _stream_ptr(self.device)))
        return out

    def load_dense(self, centers: torch.Tensor, features: torch.Tensor) -> None:
        centers = as_device(centers, self.device, torch.float32)
        features = as_device(features, self.device, torch.float32)
        V = int(features.shape[0])
        if V and int(features.shape[1]) != self.dim:
            raise ValueError("feature dimension mismatch")
        N.check(N.lib.vsm_map_load_dense(self._h, _ptr(centers), _ptr(features), V, _stream_ptr(self.device)))

    # -- lookup / query ------------------------------------------------------
    def lookup(self, positions, compat: bool = True) -> np.ndarray:
        pos = as_device(np.asarray(positions, dtype=np.float32).reshape(-1, 3), self.device, torch.float32)
        out = torch.empty((pos.shape[0],), dtype=torch.int64, device=self.device)
        N.check(N.lib.vsm_lookup(self._h, _ptr(pos), int(pos.shape[0]), _ptr(out), int(bool(compat)),
                                 _stream_ptr(self.device)))
        return out.cpu().numpy()

    def query(self, q, top_k: int = 1, normalize: bool = False, engine: int = 0):
        """q: (P,d) float32 (numpy or tensor).  Returns device tensors idx (P,k) int64, scores (P,k) float32."""
        qt = as_device(q, self.device, torch.float32)
        if qt.ndim == 1:
            qt = qt[None, :]
        if qt.ndim != 2 or qt.shape[1] != self.dim:
            raise ValueError(f"query embeddings must have shape (P,{self.dim}), got {tuple(qt.shape)}")
        P = int(qt.shape[0])
        idx = torch.empty((P, top_k), dtype=torch.int64, device=self.device)
        sc = torch.empty((P, top_k), dtype=torch.float32, device=self.device)
        N.check(N.lib.vsm_query(self._h, _ptr(qt), P, int(top_k), int(bool(normalize)), int(engine), _ptr(idx),
                                _ptr(sc), _stream_ptr(self.device)))
        return idx, sc

    def query_stats(self) -> dict:
        a, b = C.c_int64(), C.c_int64()
        N.check(N.lib.vsm_query_stats(self._h, C.byref(a), C.byref(b)))
        return {"last_candidates": int(a.value), "fallbacks": int(b.value)}

    # -- multi-GPU partials ---------------------------------------------------
    def partials_counts(self, world: int) -> np.ndarray:
        cnt = (C.c_int64 * world)()
        N.check(N.lib.vsm_partials_pack(self._h, world, None, None, None, cnt, _stream_ptr(self.device)))
        return np.asarray(list(cnt), dtype=np.int64)

    def partials_pack(self, world: int):
        V = self.num_voxels
        dev = self.device
        keys = torch.empty((V,), dtype=torch.int64, device=dev)
        counts = torch.empty((V,), dtype=torch.int32, device=dev)
        sums = torch.empty((V, self.dim), dtype=torch.float32, device=dev)
        cnt = (C.c_int64 * world)()
        N.check(N.lib.vsm_partials_pack(self._h, world, _ptr(keys), _ptr(counts), _ptr(sums), cnt, _stream_ptr(dev)))
        return keys, counts, sums, np.asarray(list(cnt), dtype=np.int64)

    def partials_merge(self, keys: torch.Tensor, counts: torch.Tensor, sums: torch.Tensor) -> None:
        n = int(keys.shape[0])
        N.check(N.lib.vsm_partials_merge(self._h, _ptr(keys), _ptr(counts), _ptr(sums), n, _stream_ptr(self.device)))

    def contrib_pack(self, world: int):
        dev = self.device
        cnt = (C.c_int64 * world)()
        N.check(N.lib.vsm_contrib_pack(self._h, world, None, None, None, cnt, _stream_ptr(dev)))
        M = int(sum(cnt))
        keys = torch.empty((M,), dtype=torch.int64, device=dev)
        subs = torch.empty((M,), dtype=torch.int32, device=dev)
        masks = torch.empty((M, 2), dtype=torch.int64, device=dev)
        N.check(N.lib.vsm_contrib_pack(self._h, world, _ptr(keys), _ptr(subs), _ptr(masks), cnt, _stream_ptr(dev)))
        return keys, subs, masks, np.asarray(list(cnt), dtype=np.int64)

    def contrib_merge(self, keys: torch.Tensor, subs: torch.Tensor, masks: torch.Tensor) -> None:
        n = int(keys.shape[0])
        N.check(N.lib.vsm_contrib_merge(self._h, _ptr(keys), _ptr(subs), _ptr(masks), n, _stream_ptr(self.device)))


def conf_threshold(conf, percentile: float, device: Optional[torch.device] = None) -> np.float32:
    """np.percentile(conf, percentile) on the device (vggt_slam/submap.py:38)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = as_device(conf, dev, torch.float32).reshape(-1)
    out = C.c_float()
    N.check(N.lib.vsm_conf_threshold(_ptr(t), int(t.numel()), float(percentile), C.byref(out), _stream_ptr(dev)))
    return np.float32(out.value)


def transform_points(points, H_world_map, out_f64: bool = True, device: Optional[torch.device] = None) -> torch.Tensor:
    """(H @ [p;1]) / w for points (...,3) float32 -> same leading shape, float64 (or float32)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = as_device(points, dev, torch.float32)
    n = t.numel() // 3
    out = torch.empty(t.shape, dtype=torch.float64 if out_f64 else torch.float32, device=dev)
    Hm = (C.c_double * 16)(*np.asarray(H_world_map, dtype=np.float64).reshape(16).tolist())
    N.check(N.lib.vsm_transform_points(_ptr(t), int(n), Hm, _ptr(out), int(out_f64), _stream_ptr(dev)))
    return out


def select_points(points, conf, colors, stride: int, conf_threshold_value, H_world_map, want_world: bool,
                  want_colors: bool, device: Optional[torch.device] = None):
    """Boolean gather conf >= thr on the stride grid (+ transform): submap.py:155-164, 182-188, 217-219."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    conf_t = as_device(conf, dev, torch.float32)
    S, H, W = conf_t.shape
    hs, ws = -(-H // stride), -(-W // stride)
    cap = S * hs * ws
    pts_t = as_device(points, dev, torch.float32) if want_world else None
    col_t = as_device(colors, dev, torch.uint8) if want_colors else None
    out_w = torch.empty((cap, 3), dtype=torch.float64, device=dev) if want_world else None
    out_c = torch.empty((cap, 3), dtype=torch.uint8, device=dev) if want_colors else None
    Hm = None
    if want_world:
        Hm = (C.c_double * 16)(*np.asarray(H_world_map, dtype=np.float64).reshape(16).tolist())
    n = C.c_int64()
    N.check(N.lib.vsm_select_points(_ptr(pts_t), _ptr(conf_t), _ptr(col_t), S, H, W, int(stride),
                                    float(conf_threshold_value), Hm, _ptr(out_w), _ptr(out_c), C.byref(n),
                                    _stream_ptr(dev)))
    n = int(n.value)
    return (out_w[:n] if want_world else None), (out_c[:n] if want_colors else None)
