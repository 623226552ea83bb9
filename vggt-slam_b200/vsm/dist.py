"""Multi-GPU build and query.  One process per GPU, torch.distributed (NCCL over
NVLink) for the plumbing; the reference is single-process and has no counterpart.

  * submaps are sharded over ranks (every rank's GraphMap holds its own submaps);
  * each rank fuses its submaps into a local map (the three per-submap filters need no communication);
  * a voxel is owned by rank  hash(key) % world  (libvsm's owner_of): local partial voxels
    (key, count, fp32 sums) and contributor entries are grouped by owner on the device
    (vsm_partials_pack / vsm_contrib_pack), moved with one all-to-all per array, and merged into the
    owner's map (vsm_partials_merge / vsm_contrib_merge);
  * the owner shard is finalised; a global index (np.unique order over ALL voxels) comes from an
    all-gather of the sorted packed keys (8 bytes per voxel);
  * queries run on every shard and the P x k candidates are all-gathered and merged.

The exchange helpers take plain tensors so that the plumbing is testable on CPU with gloo.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def exchange_counts(send_counts: Sequence[int], group=None, device=None) -> List[int]:
    """all-to-all of one int64 per peer: how many records each peer will send me."""
    world = dist.get_world_size(group)
    t = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(out, t, group=group)
    return [int(x) for x in out.cpu().tolist()]


def exchange_rows(x: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int], group=None) -> torch.Tensor:
    """all-to-all of rows: x is owner-major (first send_counts[0] rows go to rank 0, ...)."""
    out = torch.empty((int(sum(recv_counts)),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(out, x.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                           group=group)
    return out


def merge_topk(idx: torch.Tensor, score: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge gathered candidates: idx, score are (P, world*k) global indices and scores -> (P,k) best,
    ties broken by the lower global index (the library's policy)."""
    P, n = idx.shape
    order = torch.argsort(idx, dim=1, stable=True)
    idx_s = torch.gather(idx, 1, order)
    sc_s = torch.gather(score, 1, order)
    sc_key = torch.where(torch.isnan(sc_s), torch.full_like(sc_s, float("inf")), sc_s)  # NaN ranks first, like topk
    order2 = torch.argsort(sc_key, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(idx_s, 1, order2), torch.gather(sc_s, 1, order2)


def global_ranks(my_sorted_keys: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """Global index of each of my (sorted, unique, disjoint across ranks) keys in the order of ALL keys."""
    world = dist.get_world_size(group)
    n = torch.tensor([my_sorted_keys.numel()], dtype=torch.int64, device=my_sorted_keys.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.full((mx,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=my_sorted_keys.device)
    pad[: my_sorted_keys.numel()] = my_sorted_keys
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    ranks = torch.zeros_like(my_sorted_keys)
    for r, b in enumerate(bufs):
        ranks += torch.searchsorted(b[: sizes[r]], my_sorted_keys)  # keys smaller than mine on rank r
    return ranks, int(sum(sizes))


class ShardedVoxelMap:
    """The owner shard of a multi-GPU map plus what is needed to talk about global voxel indices."""

    def __init__(self, local_map, device_map, global_index: torch.Tensor, n_global: int, group=None):
        self.local = local_map  # SemanticVoxelMap over this rank's voxels (sorted order within the shard)
        self._dm = device_map
        self.global_index = global_index  # (V_local,) int64: index in the order of all voxels
        self.n_global = n_global
        self.group = group

    def query_with_embeddings(self, qe: np.ndarray, top_k: int = 1, normalize: bool = False):
        """Collective: every rank passes the same prompts; returns (global indices (P,k), scores (P,k))."""
        world = dist.get_world_size(self.group)
        dev = self._dm.device
        q = torch.as_tensor(np.asarray(qe, dtype=np.float32), device=dev)
        if q.ndim == 1:
            q = q[None, :]
        P = q.shape[0]
        kk = min(top_k, self._dm.num_voxels)
        idx = torch.full((P, top_k), -1, dtype=torch.int64, device=dev)
        sc = torch.full((P, top_k), float("-inf"), dtype=torch.float32, device=dev)
        if kk > 0:
            li, ls = self._dm.query(q, top_k=kk, normalize=normalize)
            idx[:, :kk] = self.global_index[li]
            sc[:, :kk] = ls
        gi = [torch.empty_like(idx) for _ in range(world)]
        gs = [torch.empty_like(sc) for _ in range(world)]
        dist.all_gather(gi, idx, group=self.group)
        dist.all_gather(gs, sc, group=self.group)
        ai, as_ = torch.cat(gi, dim=1), torch.cat(gs, dim=1)
        ai = torch.where(ai < 0, torch.full_like(ai, torch.iinfo(torch.int64).max), ai)
        bi, bs = merge_topk(ai, as_, top_k)
        return bi.cpu().numpy(), bs.cpu().numpy()


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    host-streaming path are allocated on the GPU's NUMA node (one process per GPU on a two-socket box otherwise
    leaves half the ranks copying across sockets).  Returns the CPU list, or None if NVML / affinity is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        want = f"{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0" if hasattr(props, "pci_bus_id") else None
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        if want is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                b = pynvml.nvmlDeviceGetPciInfo(hi).busId
                b = b.decode() if isinstance(b, bytes) else b
                if b.lower().endswith(want):
                    h = hi
                    break
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


class _Phases:
    """Wall-clock phases of build_sharded when VSM_DIST_TRACE=1 (synchronises the device at every mark)."""

    def __init__(self, device):
        import os
        import time

        self.on = os.environ.get("VSM_DIST_TRACE") == "1"
        self.device, self.time, self.marks = device, time, []
        self.mark("start")

    def mark(self, name):
        if self.on:
            torch.cuda.synchronize(self.device)
            self.marks.append((name, self.time.perf_counter()))

    def report(self):
        if self.on and dist.get_rank() == 0:
            t = self.marks
            print("[vsm dist] " + "  ".join(f"{t[i][0]} {1e3 * (t[i][1] - t[i - 1][1]):.2f}" for i in range(1, len(t))),
                  flush=True)


_NAMES_CACHE: dict = {}
_PEER_BROKEN: dict = {}  # group -> True once the peer-memory exchange has failed to set up (transport="auto")


def _names_signature(fused) -> int:
    """62-bit signature of this rank's (submap id, frame ids, frame names); stable within the process."""
    items = []
    for f in fused:
        sm = f["submap"]
        items.append((int(sm.get_id()), tuple(map(str, sm.frame_ids)), tuple(sorted((sm.frame_id_to_name or {}).items()))))
    return hash(tuple(sorted(items))) & ((1 << 62) - 1)


def _sizes_all(values, group, device) -> np.ndarray:
    """all-gather of a few int64 per rank -> (world, len(values))"""
    world = dist.get_world_size(group)
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy()


def build_sharded(graph_map, voxel_size: float, stride: int = 1, ignore_loop_closure_frames: bool = True,
                  capacity_hint: Optional[int] = None, host_streaming: Optional[bool] = None, profile: bool = False,
                  group=None, transport: str = "auto"):
    """Collective multi-GPU build.  Returns (ShardedVoxelMap, this rank's per-submap fuse stats).

    transport "peer": voxel partials and contributor entries are pushed straight into the owners' inboxes over
    NVLink peer memory by one kernel pass (vsm/peer.py, csrc/peer.cu); "collective": pack on the device, one
    torch.distributed all-to-all per array, merge (works with any backend); "auto" (default): "peer", and
    "collective" from then on if CUDA IPC / peer access turns out not to work between the GPUs of the group (all
    ranks learn that together)."""
    from . import voxel_map as vm
    from .map import wrap_device_map

    world = dist.get_world_size(group)
    ph = _Phases(torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None)
    dm, fused, names = graph_map.fuse_into_device_map(voxel_size, stride, ignore_loop_closure_frames, True,
                                                      capacity_hint, host_streaming, profile)
    stats = graph_map.last_build_stats
    if dm is None:
        raise RuntimeError("build_sharded: this rank has no submap to fuse")
    dev = dm.device
    d, code = dm.dim, dm.emb_dtype
    names_sigs = None
    ph.mark("fuse")
    auto_key = id(group) if group is not None else 0
    if transport == "auto":
        transport = "collective" if _PEER_BROKEN.get(auto_key) else "peer"
        fall_back = True
    else:
        fall_back = False
    ex = None
    if transport == "peer":
        from . import peer

        # the inboxes must hold what the owners receive: a hash spreads sum(V_r) voxels evenly over the owners;
        # 25 % + 64 K slots of slack, never more than everything
        n_log = int(sum(s["n_submap_voxels"] for s in stats))
        sizes = _sizes_all([dm.num_voxels, n_log, _names_signature(fused)], group, dev)
        names_sigs = sizes[:, 2]
        tot_v, tot_c = int(sizes[:, 0].sum()), int(sizes[:, 1].sum())
        need_rows = min(tot_v, int(1.25 * tot_v / world) + (1 << 16)) + 1
        need_contrib = min(tot_c, int(1.25 * tot_c / world) + (1 << 16)) + 1
        ph.mark("sizes")
        try:
            ex = peer.exchange_for(dev, d, need_rows, need_contrib, group)
        except peer.PeerUnavailable:
            if not fall_back:
                raise
            _PEER_BROKEN[auto_key] = True
            transport = "collective"
    if transport == "peer":
        ex.push(dm)
        ph.mark("push")
        owner = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=ex.cap_rows, device=dev)
        ex.drain(owner)
        ph.mark("drain")
        dm.close()
        ph.mark("close")
    if transport == "collective":
        # ---- voxels -> owners ------------------------------------------------------
        keys, counts, sums, send = dm.partials_pack(world)
        recv = exchange_counts(send.tolist(), group, dev)
        r_keys = exchange_rows(keys, send.tolist(), recv, group)
        r_counts = exchange_rows(counts, send.tolist(), recv, group)
        r_sums = exchange_rows(sums, send.tolist(), recv, group)
        ckeys, csubs, cmasks, csend = dm.contrib_pack(world)
        crecv = exchange_counts(csend.tolist(), group, dev)
        rc_keys = exchange_rows(ckeys, csend.tolist(), crecv, group)
        rc_subs = exchange_rows(csubs, csend.tolist(), crecv, group)
        rc_masks = exchange_rows(cmasks, csend.tolist(), crecv, group)
        del keys, counts, sums, ckeys, csubs, cmasks
        dm.close()
        # ---- owner merge --------------------------------------------------------------
        owner = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=max(int(sum(recv)), 1024), device=dev)
        owner.partials_merge(r_keys, r_counts, r_sums)
        owner.contrib_merge(rc_keys, rc_subs, rc_masks)
        del r_keys, r_counts, r_sums
    if transport not in ("peer", "collective"):
        raise ValueError(f"unknown transport {transport!r}")
    owner.finalize()
    ph.mark("finalize")
    # frame ids of every submap, from every rank (contributor lists name frames of remote submaps too)
    mine = {int(f["submap"].get_id()): (list(f["submap"].frame_ids), dict(f["submap"].frame_id_to_name or {}))
            for f in fused}
    # The pickled all-gather costs ~1 ms: skip it while no rank's frame ids have changed since the last build.  Every
    # rank sees the same vector of per-rank signatures (all-gathered with the sizes), so all ranks decide alike.
    my_sig = _names_signature(fused)
    if names_sigs is None:
        names_sigs = tuple(int(x) for x in _sizes_all([my_sig], group, dev)[:, 0])
    else:
        names_sigs = tuple(int(x) for x in names_sigs)
    cache_key = (id(group) if group is not None else 0, dev.index)
    cached = _NAMES_CACHE.get(cache_key)
    if cached is not None and cached[0] == names_sigs and cached[1] == my_sig:
        everyone = cached[2]
    else:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        _NAMES_CACHE[cache_key] = (names_sigs, my_sig, everyone)
    frame_ids, all_names = {}, {}
    for part in everyone:
        for sid, (ids, nm) in part.items():
            frame_ids[sid] = ids
            all_names[str(sid)] = nm

    class _SubmapStub:
        def __init__(self, sid, ids):
            self._sid, self.frame_ids = sid, ids

        def get_id(self):
            return self._sid

    fused_all = [{"submap": _SubmapStub(sid, ids)} for sid, ids in frame_ids.items()]
    ph.mark("names")
    local = wrap_device_map(owner, fused_all, all_names, voxel_size, True, False) if owner.num_voxels else None
    ph.mark("wrap")
    gidx, n_global = global_ranks(owner.export_packed_keys(), group)
    ph.mark("ranks")
    ph.report()
    return ShardedVoxelMap(local, owner, gidx, n_global, group), stats
