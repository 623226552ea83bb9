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


def _global_ranks_torch(my_sorted_keys: torch.Tensor, bufs, sizes) -> torch.Tensor:
    """CPU tensors only (the gloo plumbing test): one searchsorted per shard."""
    ranks = torch.zeros_like(my_sorted_keys)
    for r, b in enumerate(bufs):
        ranks += torch.searchsorted(b[: sizes[r]], my_sorted_keys)  # keys smaller than mine on rank r
    return ranks


def global_ranks(my_sorted_keys: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """Global index of each of my (sorted, unique, disjoint across ranks) keys in the order of ALL keys: the sorted
    keys of every shard are all-gathered (8 bytes per voxel) and ONE kernel ranks mine among them (vsm_global_ranks:
    a binary search per shard)."""
    world = dist.get_world_size(group)
    n = torch.tensor([my_sorted_keys.numel()], dtype=torch.int64, device=my_sorted_keys.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in torch.cat(sizes).cpu().tolist()]
    mx = max(max(sizes), 1)
    pad = torch.full((mx,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=my_sorted_keys.device)
    pad[: my_sorted_keys.numel()] = my_sorted_keys
    if not my_sorted_keys.is_cuda:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        return _global_ranks_torch(my_sorted_keys, bufs, sizes), int(sum(sizes))
    import ctypes as C

    from . import _native as N

    allk = torch.empty((world * mx,), dtype=torch.int64, device=my_sorted_keys.device)
    dist.all_gather_into_tensor(allk, pad, group=group)
    ranks = torch.empty_like(my_sorted_keys)
    sz = (C.c_int64 * world)(*sizes)
    N.check(N.lib.vsm_global_ranks(C.c_void_p(my_sorted_keys.data_ptr()), int(my_sorted_keys.numel()),
                                   C.c_void_p(allk.data_ptr()), int(mx), sz, world, C.c_void_p(ranks.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream(my_sorted_keys.device).cuda_stream)))
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


def _settle_before_next_build(dev, group) -> None:
    """With SM partitions on (vsm_set_option("green_prep_sms")), a rank must not start the fuse kernels of its next build
    -- they run in green contexts -- while a peer is still finishing this build's exchange / collectives against this
    rank's memory: measured on 2 x B200, steps then stall at random for 35-150 ms (10 steps: 20.2 19.7 113 109 56 143 93
    143 19.8 19.7 ms); with the device synchronised and a barrier at the end of every build all steps take 18.8-19.3 ms.
    (On 8 GPUs that is not enough: 26-41 ms per step.  There the partition goes with the collective exchange, bench.py.)"""
    from . import _native as N

    if N.option("green_prep_sms", 0) > 0 and dev is not None and dev.type == "cuda":
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)


def _finish_shard(owner, fused, voxel_size, group, dev, ph, names_sigs=None, settle=False):
    """Tail of a sharded build: finalise the owner shard, learn every rank's frame ids (contributor lists name frames
    of remote submaps too), wrap the shard and rank its voxels among all shards."""
    from .map import wrap_device_map

    world = dist.get_world_size(group)
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
    if settle:  # (the peer-memory exchange; the collective route synchronises the host at every step anyway)
        _settle_before_next_build(dev, group)
    ph.report()
    return ShardedVoxelMap(local, owner, gidx, n_global, group)


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
        ckeys, csubs, cmasks, csend = dm.contrib_pack(world)
        ph.mark("pack")
        # one small all-to-all tells every rank what it will receive from whom -- voxel rows, contributor entries --
        # and carries the senders' frame-name signatures (which decide whether the pickled all-gather of names is needed)
        my_sig = _names_signature(fused)
        meta = torch.tensor([[int(send[r]), int(csend[r]), my_sig] for r in range(world)], dtype=torch.int64, device=dev)
        got = torch.empty_like(meta)
        dist.all_to_all_single(got, meta, group=group)
        got = got.cpu().numpy()
        recv, crecv, names_sigs = [int(x) for x in got[:, 0]], [int(x) for x in got[:, 1]], got[:, 2]
        ph.mark("counts")
        r_keys = exchange_rows(keys, send.tolist(), recv, group)
        r_counts = exchange_rows(counts, send.tolist(), recv, group)
        r_sums = exchange_rows(sums, send.tolist(), recv, group)
        ph.mark("a2a_rows")
        rc_keys = exchange_rows(ckeys, csend.tolist(), crecv, group)
        rc_subs = exchange_rows(csubs, csend.tolist(), crecv, group)
        rc_masks = exchange_rows(cmasks, csend.tolist(), crecv, group)
        ph.mark("a2a_contrib")
        del keys, counts, sums, ckeys, csubs, cmasks
        dm.close()
        ph.mark("close")
        # ---- owner merge --------------------------------------------------------------
        owner = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=max(int(sum(recv)), 1024), device=dev)
        owner.partials_merge(r_keys, r_counts, r_sums)
        owner.contrib_merge(rc_keys, rc_subs, rc_masks)
        ph.mark("merge")
        del r_keys, r_counts, r_sums
    if transport not in ("peer", "collective"):
        raise ValueError(f"unknown transport {transport!r}")
    return _finish_shard(owner, fused, voxel_size, group, dev, ph, names_sigs, settle=transport == "peer"), stats


_XCHG_STREAMS: dict = {}


def _exchange_stream(dev) -> "torch.cuda.Stream":
    st = _XCHG_STREAMS.get(dev.index)
    if st is None:
        st = _XCHG_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


def build_sharded_streaming(graph_map, voxel_size: float, round_submaps: int, stride: int = 1,
                            ignore_loop_closure_frames: bool = True, owner_capacity: Optional[int] = None,
                            round_capacity: Optional[int] = None, profile: bool = False, group=None,
                            timings: Optional[dict] = None):
    """Collective multi-GPU build for LONG trajectories (BASELINE configs[2]): the rank's submaps are fused in rounds of
    `round_submaps`; the voxels of round r are pushed to their owners over NVLink peer memory and merged there on an
    exchange stream WHILE round r+1 is being fused on the main stream (two local maps alternate; a drained local map is
    cleared on the exchange stream).  The local map therefore never holds more than a round, the inboxes are sized for
    one round, and the exchange is off the critical path except for the last round.  Peer memory only (no collective
    fallback: a long trajectory does not fit the pack -> all-to-all -> merge route's staging buffers).

    owner_capacity / round_capacity: voxels the owner shard / one round's local map must hold (estimated from the
    first round when omitted; too small a value fails loudly with MemoryError at the final collect).
    Returns (ShardedVoxelMap, this rank's per-submap fuse stats)."""
    from . import peer
    from . import voxel_map as vm
    from .map import wrap_device_map

    world = dist.get_world_size(group)
    if round_submaps < 1:
        raise ValueError("round_submaps must be >= 1")
    todo = graph_map.usable_submaps()
    if not todo:
        raise RuntimeError("build_sharded_streaming: this rank has no submap to fuse")
    dev = torch.device("cuda", torch.cuda.current_device())
    main = torch.cuda.current_stream(dev)
    xs = _exchange_stream(dev)
    d, code = int(todo[0].semantic_embeddings.shape[-1]), todo[0].embedding_dtype_code()
    n_rounds = (len(todo) + round_submaps - 1) // round_submaps
    t_r = torch.tensor([n_rounds], dtype=torch.int64, device=dev)
    dist.all_reduce(t_r, op=dist.ReduceOp.MAX, group=group)
    n_rounds = int(t_r.item())
    ph = _Phases(dev)
    locals_ = [None, None]
    free_ev = [None, None]
    owner, ex = None, None
    stats_all, fused_all, names = [], [], {}
    prof = None
    pushed_rows = pushed_contrib = 0
    round_ev = [] if (timings is not None and timings.get("per_round")) else None  # exchange-stream events per round
    for r in range(n_rounds):
        chunk = todo[r * round_submaps:(r + 1) * round_submaps]
        b = r & 1
        if locals_[b] is None:
            locals_[b] = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=round_capacity or (1 << 18))
        L = locals_[b]
        if free_ev[b] is not None:
            main.wait_event(free_ev[b])  # the push of round r-2 has read this map and it has been cleared
        if chunk:
            _, fused, nm = graph_map.fuse_into_device_map(voxel_size, stride, ignore_loop_closure_frames, True, None,
                                                          False, profile, dm=L, submaps=chunk)
            st = graph_map.last_build_stats
            stats_all += st
            fused_all += fused
            names.update(nm)
            if profile and graph_map.last_profile is not None:
                prof = dict(graph_map.last_profile) if prof is None else {k: prof[k] + v for k, v in graph_map.last_profile.items()}
        n_log = L.num_log_entries
        pushed_rows += L.num_voxels
        pushed_contrib += n_log
        ev = torch.cuda.Event()
        ev.record(main)
        if r == 0:
            # the first round sizes the inboxes and, unless the caller knows better, the owner shard
            sizes = _sizes_all([L.num_voxels, n_log], group, dev)
            tot_v, tot_c = int(sizes[:, 0].sum()), int(sizes[:, 1].sum())
            need_rows = int(1.5 * tot_v / world) + (1 << 16)
            need_contrib = int(1.5 * tot_c / world) + (1 << 16)
            ex = peer.exchange_for(dev, d, need_rows, need_contrib, group)
            cap = owner_capacity or (int(1.3 * tot_v * n_rounds / world) + (1 << 16))
            owner = vm.DeviceVoxelMap(float(voxel_size), d, code, capacity=cap, device=dev)
            owner.reserve_log(int(1.3 * tot_c * n_rounds / world) + (1 << 16))
            torch.cuda.synchronize(dev)  # the owner's allocation (zeroed rows) is complete before any stream uses it
        with torch.cuda.stream(xs):
            xs.wait_event(ev)
            if round_ev is not None:
                marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                round_ev.append(marks)
                marks[0].record(xs)
            ex.push(L)
            if round_ev is not None:
                marks[1].record(xs)
            L.clear_async()
            free_ev[b] = torch.cuda.Event()
            free_ev[b].record(xs)
            if round_ev is not None:
                marks[2].record(xs)
            ex.drain_async(owner, slot=0)
            if round_ev is not None:
                marks[3].record(xs)
    ph.mark("rounds")
    with torch.cuda.stream(xs):
        got_rows, got_contrib = peer.drain_collect(owner, slot=0)  # synchronises the exchange stream; errors surface here
    main.wait_stream(xs)
    ph.mark("last_exchange")
    for L in locals_:
        if L is not None:
            L.close()
    graph_map.last_build_stats = stats_all
    graph_map.last_profile = prof
    shard = _finish_shard(owner, fused_all, voxel_size, group, dev, ph, settle=True)
    if round_ev:
        torch.cuda.synchronize(dev)
        timings["per_round_ms"] = {"push": [m[0].elapsed_time(m[1]) for m in round_ev],
                                   "clear": [m[1].elapsed_time(m[2]) for m in round_ev],
                                   "wait_and_drain": [m[2].elapsed_time(m[3]) for m in round_ev]}
    if timings is not None:
        timings.update({"rounds": n_rounds, "pushed_rows": int(pushed_rows), "pushed_contrib": int(pushed_contrib),
                        "received_rows": int(got_rows), "received_contrib": int(got_contrib),
                        "inbox_rows": int(ex.cap_rows), "row_bytes": 16 + 4 * d,
                        "phases_ms": {ph.marks[i][0]: 1e3 * (ph.marks[i][1] - ph.marks[i - 1][1])
                                      for i in range(1, len(ph.marks))} if ph.on else None})
    return shard, stats_all


def shard_invariants(shard: "ShardedVoxelMap", stats, group=None) -> dict:
    """Collective, cheap at any size: sum of the per-voxel counts over all shards == points fused by all ranks; the
    shards' sizes add up to the global voxel count; the global indices are a permutation (sum and sum of squares)."""
    dm = shard._dm
    dev = dm.device
    V = dm.num_voxels
    counts = dm.export_geometry(coords=False, centers=False, counts=True, recon=False)[2] if V else torch.zeros(0, dtype=torch.int64, device=dev)
    g = shard.global_index.to(torch.float64)
    t = torch.tensor([float(counts.sum().item()) if V else 0.0, float(sum(s["n_fused"] for s in stats)), float(V),
                      float(g.sum().item()) if V else 0.0, float((g * g).sum().item()) if V else 0.0],
                     dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = float(shard.n_global)
    out = {"points_in_voxels": int(t[0].item()), "points_fused": int(t[1].item()), "voxels": int(t[2].item()),
           "n_global": int(shard.n_global)}
    out["counts_conserved"] = out["points_in_voxels"] == out["points_fused"]
    out["shards_cover_map"] = out["voxels"] == out["n_global"]
    out["indices_are_a_permutation"] = bool(abs(t[3].item() - n * (n - 1) / 2) < 0.5 and
                                            abs(t[4].item() - (n - 1) * n * (2 * n - 1) / 6) <= 1e-9 * max(t[4].item(), 1.0))
    out["ok"] = bool(out["counts_conserved"] and out["shards_cover_map"] and out["indices_are_a_permutation"])
    return out


def parity_check(make_submap, n_submaps: int, voxel_size: float, dim: int, group=None, round_submaps: Optional[int] = None,
                 top_k: int = 5) -> dict:
    """Collective correctness check of the multi-GPU build, small enough to run before a benchmark: `n_submaps`
    submaps (``make_submap(i) -> vsm.Submap``, a pure function of i) are built sharded over the group -- one-shot
    exchange, or streaming rounds when `round_submaps` is given -- and rank 0 also builds ALL of them on its own GPU;
    the union of the shards must equal that map: keys, counts, contributors bit-exact, features to 1e-3, and the
    sharded query must return the single-GPU query's indices.  Returns a report dict ("ok": bool) on every rank."""
    from .map import GraphMap

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gm = GraphMap()
    for i in range(rank, n_submaps, world):
        gm.add_submap(make_submap(i))
    if round_submaps:
        sh, stats = build_sharded_streaming(gm, voxel_size, round_submaps, group=group)
    else:
        sh, stats = build_sharded(gm, voxel_size, group=group)
    inv = shard_invariants(sh, stats, group)
    dm = sh._dm
    V = dm.num_voxels
    keys = dm.export_packed_keys().cpu().numpy() if V else np.zeros(0, np.int64)
    counts = dm.export_geometry(coords=False, centers=False, counts=True, recon=False)[2].cpu().numpy() if V else np.zeros(0, np.int64)
    feats = dm.features_to_host() if V else np.zeros((0, dim), np.float32)
    contribs = sh.local.get_contributors().tolist() if sh.local is not None else []
    gidx = sh.global_index.cpu().numpy()
    rng = np.random.default_rng(5)
    Q = rng.normal(size=(4, dim)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    qi, qs = sh.query_with_embeddings(Q, top_k=top_k)
    parts = [None] * world
    dist.all_gather_object(parts, (keys, counts, feats, contribs, gidx), group=group)
    report = {"world": world, "submaps": n_submaps, "mode": "streaming" if round_submaps else "one-shot", "invariants": inv}
    if rank == 0:
        gm1 = GraphMap()
        for i in range(n_submaps):
            gm1.add_submap(make_submap(i))
        single = gm1.build_semantic_voxel_map(voxel_size)
        sdm = single._dm
        s_keys = sdm.export_packed_keys().cpu().numpy()
        s_counts = sdm.export_geometry(coords=False, centers=False, counts=True, recon=False)[2].cpu().numpy()
        all_gidx = np.concatenate([p[4] for p in parts])
        order = np.argsort(all_gidx, kind="stable")
        report["voxels"] = int(len(s_keys))
        report["keys"] = bool(len(all_gidx) == len(s_keys) and np.array_equal(np.sort(all_gidx), np.arange(len(s_keys)))
                              and np.array_equal(np.concatenate([p[0] for p in parts])[order], s_keys))
        report["counts"] = bool(report["keys"] and np.array_equal(np.concatenate([p[1] for p in parts])[order], s_counts))
        f_all = np.concatenate([p[2] for p in parts])[order] if report["keys"] else None
        f_ref = single.get_features()
        report["features_rtol"] = 1e-3
        report["features"] = bool(report["keys"] and np.allclose(f_all, f_ref, rtol=1e-3, atol=1e-5))
        allc = sum([p[3] for p in parts], [])
        report["contributors"] = bool(report["keys"] and [allc[i] for i in order] == single.get_contributors().tolist())
        si, _, ss = single.query_with_embeddings(Q, top_k=top_k)
        report["query"] = bool(np.array_equal(qi, si) and np.allclose(qs, ss, rtol=1e-3, atol=1e-6))
        report["ok"] = bool(inv["ok"] and all(report[k] for k in ("keys", "counts", "features", "contributors", "query")))
    box = [report]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]
