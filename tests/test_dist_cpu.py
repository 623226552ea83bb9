"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing in vsm.dist: count / row exchange by owner,
global ranking of disjoint sorted key sets, and the top-k merge.  The CUDA pack / merge kernels themselves
are covered by the gpu tests (test_gpu_dist.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util

        spec = importlib.util.spec_from_file_location("vsm_dist", os.path.join(ROOT, "vggt-slam_b200", "vsm", "dist.py"))
        vd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(vd)
        rng = np.random.default_rng(100 + rank)
        # every rank holds records (key, count, row) grouped by owner = key % world
        keys = rng.choice(1000, size=60 + 10 * rank, replace=False).astype(np.int64)
        owner = keys % world
        order = np.argsort(owner, kind="stable")
        keys = keys[order]
        send = [int((owner == o).sum()) for o in range(world)]
        rows = torch.from_numpy(np.stack([keys * 10 + j for j in range(4)], axis=1).astype(np.float32))
        recv = vd.exchange_counts(send)
        got_keys = vd.exchange_rows(torch.from_numpy(keys), send, recv)
        got_rows = vd.exchange_rows(rows, send, recv)
        assert (got_keys % world == rank).all()
        assert torch.equal(got_rows[:, 2], got_keys.float() * 10 + 2)
        # all records arrive exactly once
        tot = torch.tensor([got_keys.numel(), keys.size], dtype=torch.int64)
        dist.all_reduce(tot)
        assert tot[0] == tot[1]
        # global ranks of disjoint sorted unique key sets
        mine = torch.unique(got_keys)
        ranks, n_global = vd.global_ranks(mine)
        allk = [None] * world
        dist.all_gather_object(allk, mine.tolist())
        union = np.sort(np.concatenate([np.asarray(a, dtype=np.int64) for a in allk]))
        assert n_global == union.size
        np.testing.assert_array_equal(ranks.numpy(), np.searchsorted(union, mine.numpy()))
        # top-k merge with ties and NaN
        idx = torch.tensor([[5, 2, 9, 7], [1, 0, 3, 2]])
        sc = torch.tensor([[0.5, 0.9, 0.9, float("nan")], [0.1, 0.1, 0.1, 0.2]])
        bi, bs = vd.merge_topk(idx, sc, 3)
        assert bi.tolist() == [[7, 2, 9], [2, 0, 1]]
        # sizes / signatures gathered before a peer exchange: every rank sees the same (world, n) table
        table = vd._sizes_all([100 + rank, 7 * rank, (1 << 61) + rank], None, None)
        assert table.shape == (world, 3) and table.dtype == np.int64
        assert table[:, 0].tolist() == [100 + r for r in range(world)]
        assert table[:, 2].tolist() == [(1 << 61) + r for r in range(world)]

        class _Sm:  # what _names_signature reads of a Submap
            def __init__(self, sid, ids, names):
                self._sid, self.frame_ids, self.frame_id_to_name = sid, ids, names

            def get_id(self):
                return self._sid

        fa = [{"submap": _Sm(3, [1.0, 2.0], {"1.0": "a.png", "2.0": "b.png"})}, {"submap": _Sm(1, [5.0], None)}]
        sig = vd._names_signature(fa)
        assert 0 <= sig < (1 << 62)
        assert sig == vd._names_signature(list(reversed(fa)))            # order of the fuse records does not matter
        fb = [{"submap": _Sm(3, [1.0, 2.0], {"1.0": "a.png", "2.0": "c.png"})}, {"submap": _Sm(1, [5.0], None)}]
        assert sig != vd._names_signature(fb)                             # a renamed frame does
        # no NVML / no GPU here: binding must decline quietly
        assert vd.bind_to_gpu_numa(0) is None
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_exchange_plumbing_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
