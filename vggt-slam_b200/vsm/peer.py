"""One-sided voxel exchange over NVLink peer memory (csrc/peer.cu).

Every rank owns an inbox -- a cudaMalloc block shared through CUDA IPC -- and pushes its voxel partials and
contributor entries straight into the owners' inboxes from ONE kernel pass (pack + transfer), then signals;
the owner drains its inbox on its own stream.  torch.distributed is used once, to hand the IPC handles around
(and per exchange for a few integers: the sizes that bound the inbox)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _native as N


def inbox_bytes(dim: int, cap_rows: int, cap_contrib: int) -> int:
    out = C.c_int64(0)
    N.check(N.lib.vsm_inbox_bytes(int(dim), int(cap_rows), int(cap_contrib), C.byref(out)))
    return int(out.value)


class Inbox:
    """A zeroed inbox block on `device`; `handle` (64 bytes) lets the other processes of the node map it."""

    def __init__(self, device: torch.device, dim: int, cap_rows: int, cap_contrib: int, shared: bool = True):
        self.device, self.dim, self.cap_rows, self.cap_contrib = device, int(dim), int(cap_rows), int(cap_contrib)
        self.bytes = inbox_bytes(dim, cap_rows, cap_contrib)
        if shared and os.environ.get("VSM_PEER_DISABLE") == "1":  # e.g. containers that must not use CUDA IPC
            raise RuntimeError("peer-memory exchange disabled by VSM_PEER_DISABLE=1")
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64)()
        N.check(N.lib.vsm_peer_alloc(device.index or 0, self.bytes, C.byref(ptr), buf if shared else None))
        self.ptr = int(ptr.value)
        self.handle = bytes(buf) if shared else None

    def free(self) -> None:
        if self.ptr:
            N.check(N.lib.vsm_peer_free(self.device.index or 0, C.c_void_p(self.ptr)))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def push(dm, inbox_ptrs: Sequence[int], cap_rows: int, cap_contrib: int, epoch: int) -> None:
    """vsm_partials_push of DeviceVoxelMap `dm` into the inboxes at `inbox_ptrs` (one per rank, own included)."""
    from .voxel_map import _stream_ptr

    world = len(inbox_ptrs)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in inbox_ptrs])
    N.check(N.lib.vsm_partials_push(dm._h, world, arr, int(cap_rows), int(cap_contrib), int(epoch), _stream_ptr(dm.device)))


def drain(owner, inbox_ptr: int, world: int, cap_rows: int, cap_contrib: int, epoch: int, timeout_s: float = 30.0):
    """vsm_partials_drain into DeviceVoxelMap `owner`; returns (rows received, contributor entries received)."""
    from .voxel_map import _stream_ptr

    n_rows, n_contrib, flags = C.c_int64(0), C.c_int64(0), C.c_uint32(0)
    rc = N.lib.vsm_partials_drain(owner._h, C.c_void_p(int(inbox_ptr)), int(world), int(cap_rows), int(cap_contrib),
                                  int(epoch), float(timeout_s), C.byref(n_rows), C.byref(n_contrib), C.byref(flags),
                                  _stream_ptr(owner.device))
    N.check(rc)
    return int(n_rows.value), int(n_contrib.value)


def drain_async(owner, inbox_ptr: int, world: int, cap_rows: int, cap_contrib: int, epoch: int, slot: int = 0,
                timeout_s: float = 30.0) -> None:
    """vsm_partials_drain_async on the current stream (the owner map must have room: reserve / reserve_log)."""
    from .voxel_map import _stream_ptr

    N.check(N.lib.vsm_partials_drain_async(owner._h, C.c_void_p(int(inbox_ptr)), int(world), int(cap_rows), int(cap_contrib),
                                           int(epoch), float(timeout_s), int(slot), _stream_ptr(owner.device)))


def drain_collect(owner, slot: int = 0):
    """vsm_partials_drain_collect: synchronises the current stream; (rows, contributor entries) received by the drains
    queued on `slot` since its last collect."""
    from .voxel_map import _stream_ptr

    n_rows, n_contrib, flags = C.c_int64(0), C.c_int64(0), C.c_uint32(0)
    N.check(N.lib.vsm_partials_drain_collect(owner._h, int(slot), C.byref(n_rows), C.byref(n_contrib), C.byref(flags),
                                             _stream_ptr(owner.device)))
    return int(n_rows.value), int(n_contrib.value)


class PeerUnavailable(RuntimeError):
    """CUDA IPC / peer access does not work between the GPUs of this group (raised on EVERY rank)."""


class PeerExchange:
    """The inboxes of a process group (one node, one process per GPU), mapped into this process."""

    def __init__(self, device: torch.device, dim: int, cap_rows: int, cap_contrib: int, group=None):
        self.group, self.device, self.dim = group, device, int(dim)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap_rows, self.cap_contrib = int(cap_rows), int(cap_contrib)
        self.epoch = 0
        self.ptrs: List[int] = []
        self._opened: List[int] = []
        self.mine = None
        err = None
        try:
            self.mine = Inbox(device, dim, cap_rows, cap_contrib, shared=True)
        except Exception as e:  # e.g. no CUDA IPC in this environment
            err = e
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, None if self.mine is None else self.mine.handle, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.ptrs.append(self.mine.ptr)
                        continue
                    p = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(h)
                    N.check(N.lib.vsm_peer_open(device.index or 0, hb, C.byref(p)))
                    self.ptrs.append(int(p.value))
                    self._opened.append(int(p.value))
            except Exception as e:  # no peer access between some pair of GPUs
                err = e
        # every rank must reach the same verdict: one failure anywhere makes the exchange unusable everywhere
        oks: List[Optional[bool]] = [None] * self.world
        dist.all_gather_object(oks, err is None and all(h is not None for h in handles), group=group)
        if not all(oks):
            self.close()
            raise PeerUnavailable(f"peer-memory exchange unavailable on rank(s) {[r for r, ok in enumerate(oks) if not ok]}"
                                  + (f": {err}" if err is not None else ""))

    def push(self, dm) -> None:
        push(dm, self.ptrs, self.cap_rows, self.cap_contrib, self.epoch)

    def drain(self, owner, timeout_s: float = 30.0):
        out = drain(owner, self.mine.ptr, self.world, self.cap_rows, self.cap_contrib, self.epoch, timeout_s)
        self.epoch += 1
        return out

    def drain_async(self, owner, slot: int = 0, timeout_s: float = 30.0) -> None:
        drain_async(owner, self.mine.ptr, self.world, self.cap_rows, self.cap_contrib, self.epoch, slot, timeout_s)
        self.epoch += 1

    def close(self) -> None:
        """Collective: nobody may unmap or free while a peer could still be writing."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for p in self._opened:
            N.lib.vsm_peer_close(self.device.index or 0, C.c_void_p(p))
        self._opened = []
        dist.barrier(group=self.group)
        if self.mine is not None:
            self.mine.free()


_EXCHANGES: dict = {}


def exchange_for(device: torch.device, dim: int, need_rows: int, need_contrib: int, group=None) -> PeerExchange:
    """The cached exchange of (group, device, dim), re-created (collectively: every rank passes the same sizes) when
    an inbox has to grow."""
    key = (id(group) if group is not None else 0, device.index or 0, int(dim))
    ex = _EXCHANGES.get(key)
    if ex is not None and (ex.cap_rows < need_rows or ex.cap_contrib < need_contrib):
        ex.close()
        ex = None
    if ex is None:
        ex = PeerExchange(device, dim, max(int(need_rows * 1.5), 1 << 16), max(int(need_contrib * 1.5), 1 << 16), group)
        _EXCHANGES[key] = ex
    return ex


def close_all() -> None:
    for ex in list(_EXCHANGES.values()):
        ex.close()
    _EXCHANGES.clear()
