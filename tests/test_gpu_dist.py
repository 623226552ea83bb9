"""Multi-GPU exchange kernels on ONE GPU: two 'ranks' are two local maps in one process; their packed
partials are routed by hand exactly as vsm.dist.build_sharded routes them with all-to-all, merged into two
owner maps, and the union is compared with the single-map build (keys, counts bit-exact; features 1e-3)."""
import numpy as np
import pytest

import golden_io as gio
from vsm import synth

pytestmark = pytest.mark.gpu


def _local_map(vsm, subs, voxel_size):
    from test_gpu_parity import to_submap

    gm = vsm.GraphMap()
    for s in subs:
        gm.add_submap(to_submap(vsm, s, device_inputs=True))
    return gm


def test_pack_route_merge_equals_single_map():
    import torch
    import vsm
    from vsm import voxel_map as vm
    from vsm.map import wrap_device_map

    world = 2
    subs = [synth.make_submap(61, i, S=4, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=4 * i) for i in range(4)]
    single = _local_map(vsm, subs, 0.05).build_semantic_voxel_map(0.05)
    s_coords, _, s_counts, _ = single._dm.export_geometry()
    s_keys = single._dm.export_packed_keys().cpu().numpy()

    packs, cpacks, fused_all, names = [], [], [], {}
    for r in range(world):
        gm = _local_map(vsm, subs[r::world], 0.05)
        dm, fused, nm = gm.fuse_into_device_map(0.05)
        packs.append(dm.partials_pack(world))
        cpacks.append(dm.contrib_pack(world))
        # the size query agrees with the real pack
        np.testing.assert_array_equal(dm.partials_counts(world), packs[-1][3])
        assert int(packs[-1][3].sum()) == dm.num_voxels
        fused_all += fused
        names.update(nm)
    owners = []
    for o in range(world):
        om = vm.DeviceVoxelMap(0.05, 64, 0, capacity=1024)  # small: exercises growth + rehash
        for r in range(world):
            keys, counts, sums, cnt = packs[r]
            lo = int(cnt[:o].sum())
            hi = lo + int(cnt[o])
            om.partials_merge(keys[lo:hi].contiguous(), counts[lo:hi].contiguous(), sums[lo:hi].contiguous())
            ckeys, csubs, cmasks, ccnt = cpacks[r]
            lo = int(ccnt[:o].sum())
            hi = lo + int(ccnt[o])
            om.contrib_merge(ckeys[lo:hi].contiguous(), csubs[lo:hi].contiguous(), cmasks[lo:hi].contiguous())
        om.finalize()
        owners.append(om)
    all_keys = np.concatenate([o.export_packed_keys().cpu().numpy() for o in owners])
    assert len(np.unique(all_keys)) == len(all_keys) == len(s_keys)   # disjoint ownership, nothing lost
    order = np.argsort(all_keys.view(np.uint64), kind="stable")
    np.testing.assert_array_equal(all_keys[order], s_keys)
    coords = np.concatenate([o.export_geometry()[0].cpu().numpy() for o in owners])[order]
    counts = np.concatenate([o.export_geometry()[2].cpu().numpy() for o in owners])[order]
    feats = np.concatenate([o.features_to_host() for o in owners])[order]
    np.testing.assert_array_equal(coords, s_coords.cpu().numpy())
    np.testing.assert_array_equal(counts, s_counts.cpu().numpy())
    np.testing.assert_allclose(feats, single.get_features(), rtol=1e-3, atol=1e-5)
    # contributors of the union == contributors of the single map
    contribs = []
    for o in owners:
        m = wrap_device_map(o, fused_all, names, 0.05)
        contribs += m.get_contributors().tolist()
    contribs = [contribs[i] for i in order]
    assert contribs == single.get_contributors().tolist()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_push_drain_equals_single_map(world):
    """The one-sided exchange (csrc/peer.cu) with every 'rank' on this GPU: inboxes are plain device blocks, the
    maps of all ranks push into them, each owner drains its own.  Two exchanges back to back use both halves."""
    import torch
    import vsm
    from vsm import peer
    from vsm import voxel_map as vm
    from vsm.map import wrap_device_map

    dev = torch.device("cuda", 0)
    subs = [synth.make_submap(67, i, S=4, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=4 * i) for i in range(6)]
    single = _local_map(vsm, subs, 0.05).build_semantic_voxel_map(0.05)
    s_coords, _, s_counts, _ = single._dm.export_geometry()
    s_keys = single._dm.export_packed_keys().cpu().numpy()
    cap_rows, cap_contrib = len(s_keys) + 8, 6 * len(s_keys)
    inboxes = [peer.Inbox(dev, 64, cap_rows, cap_contrib, shared=False) for _ in range(world)]
    ptrs = [b.ptr for b in inboxes]
    for epoch in range(3):
        fused_all, names = [], {}
        for r in range(world):
            gm = _local_map(vsm, subs[r::world], 0.05)
            dm, fused, nm = gm.fuse_into_device_map(0.05)
            peer.push(dm, ptrs, cap_rows, cap_contrib, epoch)
            fused_all += fused
            names.update(nm)
            dm.close()
        owners = []
        n_rows_total = 0
        for o in range(world):
            om = vm.DeviceVoxelMap(0.05, 64, 0, capacity=1024)
            n_rows, n_contrib = peer.drain(om, ptrs[o], world, cap_rows, cap_contrib, epoch, timeout_s=5.0)
            n_rows_total += n_rows
            om.finalize()
            owners.append(om)
        all_keys = np.concatenate([o.export_packed_keys().cpu().numpy() for o in owners])
        assert len(np.unique(all_keys)) == len(all_keys) == len(s_keys)
        assert n_rows_total >= len(s_keys)  # voxels seen by several ranks arrive once per rank
        order = np.argsort(all_keys.view(np.uint64), kind="stable")
        np.testing.assert_array_equal(all_keys[order], s_keys)
        coords = np.concatenate([o.export_geometry()[0].cpu().numpy() for o in owners])[order]
        counts = np.concatenate([o.export_geometry()[2].cpu().numpy() for o in owners])[order]
        feats = np.concatenate([o.features_to_host() for o in owners])[order]
        np.testing.assert_array_equal(coords, s_coords.cpu().numpy())
        np.testing.assert_array_equal(counts, s_counts.cpu().numpy())
        np.testing.assert_allclose(feats, single.get_features(), rtol=1e-3, atol=1e-5)
        contribs = []
        for o in owners:
            contribs += wrap_device_map(o, fused_all, names, 0.05).get_contributors().tolist()
        assert [contribs[i] for i in order] == single.get_contributors().tolist()
        for o in owners:
            o.close()
    for b in inboxes:
        b.free()


def test_peer_inbox_overflow_and_timeout_are_reported():
    import torch
    import vsm
    from vsm import peer
    from vsm import voxel_map as vm

    dev = torch.device("cuda", 0)
    subs = [synth.make_submap(71, 0, S=3, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2))]
    gm = _local_map(vsm, subs, 0.05)
    dm, _, _ = gm.fuse_into_device_map(0.05)
    V = dm.num_voxels
    assert V > 64
    small = peer.Inbox(dev, 64, 16, 16, shared=False)
    peer.push(dm, [small.ptr], 16, 16, 0)
    om = vm.DeviceVoxelMap(0.05, 64, 0, capacity=1024)
    with pytest.raises(MemoryError, match="inbox too small"):
        peer.drain(om, small.ptr, 1, 16, 16, 0, timeout_s=5.0)
    # the half was reset by the drain: nobody pushes now, so waiting for a sender must time out, not hang
    om.clear()
    with pytest.raises(RuntimeError, match="timed out"):
        peer.drain(om, small.ptr, 1, 16, 16, 0, timeout_s=0.05)
    om.close()
    dm.close()
    small.free()


def test_rounds_with_queued_drains_equal_single_map():
    """The streaming exchange of a long build, every 'rank' on this GPU: per round each rank fuses a few submaps into
    a local map, pushes it, clears it without waiting (vsm_map_clear_async) and every owner QUEUES its drain
    (vsm_partials_drain_async); the reports are collected once, at the end.  Union of the owner shards == one map."""
    import torch
    import vsm
    from vsm import peer
    from vsm import voxel_map as vm
    from vsm.map import wrap_device_map

    world, rounds, per_round = 2, 3, 2
    dev = torch.device("cuda", 0)
    subs = [synth.make_submap(73, i, S=3, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=3 * i) for i in range(world * rounds * per_round)]
    single = _local_map(vsm, subs, 0.05).build_semantic_voxel_map(0.05)
    s_keys = single._dm.export_packed_keys().cpu().numpy()
    s_counts = single._dm.export_geometry()[2].cpu().numpy()
    cap_rows, cap_contrib = len(s_keys) + 8, 4 * len(s_keys)
    inboxes = [peer.Inbox(dev, 64, cap_rows, cap_contrib, shared=False) for _ in range(world)]
    ptrs = [b.ptr for b in inboxes]
    owners = [vm.DeviceVoxelMap(0.05, 64, 0, capacity=len(s_keys) + 1024) for _ in range(world)]
    for o in owners:
        o.reserve_log(rounds * world * cap_contrib)
    locals_ = [vm.DeviceVoxelMap(0.05, 64, 0, capacity=1 << 15) for _ in range(world)]
    fused_all, names = [], {}
    for r in range(rounds):
        for k in range(world):
            mine = subs[k::world][r * per_round:(r + 1) * per_round]
            gm = _local_map(vsm, mine, 0.05)
            _, fused, nm = gm.fuse_into_device_map(0.05, dm=locals_[k], submaps=gm.usable_submaps())
            fused_all += fused
            names.update(nm)
            assert locals_[k].num_log_entries == sum(s["n_submap_voxels"] for s in gm.last_build_stats)
            peer.push(locals_[k], ptrs, cap_rows, cap_contrib, r)
            locals_[k].clear_async()
            assert locals_[k].num_voxels == 0
        for o in range(world):
            peer.drain_async(owners[o], ptrs[o], world, cap_rows, cap_contrib, r, slot=0, timeout_s=5.0)
    got = [peer.drain_collect(o, slot=0) for o in owners]
    assert sum(g[0] for g in got) >= len(s_keys)
    for o in owners:
        o.finalize()
    all_keys = np.concatenate([o.export_packed_keys().cpu().numpy() for o in owners])
    assert len(np.unique(all_keys)) == len(all_keys) == len(s_keys)
    order = np.argsort(all_keys.view(np.uint64), kind="stable")
    np.testing.assert_array_equal(all_keys[order], s_keys)
    np.testing.assert_array_equal(np.concatenate([o.export_geometry()[2].cpu().numpy() for o in owners])[order], s_counts)
    np.testing.assert_allclose(np.concatenate([o.features_to_host() for o in owners])[order], single.get_features(),
                               rtol=1e-3, atol=1e-5)
    contribs = []
    for o in owners:
        contribs += wrap_device_map(o, fused_all, names, 0.05).get_contributors().tolist()
    assert [contribs[i] for i in order] == single.get_contributors().tolist()
    # an owner without room reports it at the collect instead of growing behind the host's back
    tiny = vm.DeviceVoxelMap(0.05, 64, 0, capacity=1024)
    gm = _local_map(vsm, subs[:8], 0.05)
    dm, _, _ = gm.fuse_into_device_map(0.05)
    assert dm.num_voxels > 1024
    peer.push(dm, [ptrs[0]], cap_rows, cap_contrib, rounds)
    peer.drain_async(tiny, ptrs[0], 1, cap_rows, cap_contrib, rounds, slot=1, timeout_s=5.0)
    with pytest.raises(MemoryError):
        peer.drain_collect(tiny, slot=1)
    for m in owners + locals_ + [tiny, dm]:
        m.close()
    for b in inboxes:
        b.free()


def test_global_ranks_kernel():
    """vsm_global_ranks against numpy: disjoint sorted shards of unequal length, padded like the all-gather buffer."""
    import ctypes as C

    import torch
    from vsm import _native as N

    rng = np.random.default_rng(3)
    keys = np.unique(rng.integers(0, 1 << 62, size=200_000, dtype=np.int64))
    owner = rng.integers(0, 3, size=keys.size)
    shards = [np.sort(keys[owner == r]) for r in range(3)]
    shards.append(np.zeros(0, np.int64))  # an empty shard
    stride = max(len(s) for s in shards)
    buf = np.full((4, stride), np.iinfo(np.int64).max, dtype=np.int64)
    for r, sh in enumerate(shards):
        buf[r, :len(sh)] = sh
    allk = torch.from_numpy(buf).cuda()
    sizes = (C.c_int64 * 4)(*[len(s) for s in shards])
    for r, sh in enumerate(shards):
        mine = torch.from_numpy(sh).cuda()
        out = torch.empty_like(mine)
        N.check(N.lib.vsm_global_ranks(C.c_void_p(mine.data_ptr()), len(sh), C.c_void_p(allk.data_ptr()), stride, sizes, 4,
                                       C.c_void_p(out.data_ptr()), None))
        np.testing.assert_array_equal(out.cpu().numpy(), np.searchsorted(keys, sh))
