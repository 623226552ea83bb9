"""Edge cases of the fuse path through the C ABI: capacity growth by retry, more queued calls than the ring holds,
submaps that contribute nothing, strides beyond the image, the frame limit, coordinates outside the packable range."""
import numpy as np
import pytest

import golden_io as gio
from oracle import voxel_oracle as vo
from vsm import synth

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-5


def _dev(s):
    import torch

    return torch.from_numpy(s.points).cuda(), torch.from_numpy(s.conf).cuda(), torch.from_numpy(s.emb).cuda()


def test_tiny_capacity_grows_by_retry_and_ring_overflow_collects_early():
    """70 fuse calls queued into a map of 1024 voxels: calls that do not fit are aborted on the device before they
    touch the map and repeated after growing; the 65th submit collects the first 64 itself.  Same map as one call
    at a time into a large map."""
    import vsm
    from vsm import _native as N
    from vsm import voxel_map as vm

    subs = [synth.make_submap(91, i, S=2, H=40, W=60, d=32, mode="sl4", room=(2.4, 1.8, 1.2), start=0.05 * i)
            for i in range(70)]
    thr = [vm.conf_threshold(s.conf, 25.0) for s in subs]
    big = vm.DeviceVoxelMap(0.02, 32, N.F32, capacity=1 << 18)   # 2 cm voxels: ~10^5 of them
    for s, t in zip(subs, thr):
        big.fuse(*_dev(s), big.make_params(2, 40, 60, 2, 1, t, s.H_world_map, s.submap_id, 0))
    big.finalize()
    r0, e0 = N.get_counter("capacity_retries"), N.get_counter("early_collects")
    small = vm.DeviceVoxelMap(0.02, 32, N.F32, capacity=1024)
    keep = []
    for s, t in zip(subs, thr):
        d = _dev(s)
        keep.append(d)
        small.fuse_async(*d, small.make_params(2, 40, 60, 2, 1, t, s.H_world_map, s.submap_id, 0))
    stats = small.collect()
    assert len(stats) == 70
    assert N.get_counter("capacity_retries") > r0 and N.get_counter("early_collects") > e0
    small.finalize()
    assert small.num_voxels == big.num_voxels > 20000
    for a, b in zip(small.export_geometry(), big.export_geometry()):
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy())
    np.testing.assert_allclose(small.features_to_host(), big.features_to_host(), rtol=1e-5, atol=1e-6)
    oa, sa, ma = small.export_contributors()
    ob, sb, mb = big.export_contributors()
    assert np.array_equal(oa, ob)
    for v in range(0, small.num_voxels, 37):   # entries of a voxel may come in any order: compare as sets
        ea = {(int(sa[e]), int(ma[e, 0]), int(ma[e, 1])) for e in range(int(oa[v]), int(oa[v + 1]))}
        eb = {(int(sb[e]), int(mb[e, 0]), int(mb[e, 1])) for e in range(int(ob[v]), int(ob[v + 1]))}
        assert ea == eb


def test_submap_that_contributes_nothing_and_stride_beyond_the_image():
    import vsm
    from test_gpu_parity import to_submap

    subs = [synth.make_submap(93, i, S=3, H=40, W=60, d=16, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=3 * i) for i in range(3)]
    gm = vsm.GraphMap()
    osubs = []
    for i, s in enumerate(subs):
        sm = to_submap(vsm, s, device_inputs=True)
        o = gio.to_oracle_submap(s)
        if i == 1:  # nothing passes this threshold
            sm.conf_threshold = np.float32(1e9)
            o.conf_threshold = np.float32(1e9)
        gm.add_submap(sm)
        osubs.append(o)
    with np.errstate(all="ignore"):
        want = vo.build_global(osubs, 0.05, exact_order=False)
    m = gm.build_semantic_voxel_map(0.05)
    np.testing.assert_array_equal(m.get_centers_world(), want.centers_world)
    np.testing.assert_allclose(m.get_features(), want.features, rtol=RTOL, atol=ATOL)
    assert m.get_contributors() == want.contributors
    assert [st["n_fused"] for st in gm.last_build_stats][1] == 0
    assert sorted(m.frame_name_maps) == ["0", "2"]          # map.py:282-297: only submaps that contributed points
    # stride beyond the image: one pixel per frame survives the grid, far too few for the coarse filter
    empty = gm.build_semantic_voxel_map(0.05, stride=1000)
    assert empty.get_centers_world().shape == (0, 3) and empty.get_features().shape == (0, 0)


def test_frame_limit_and_coordinate_range():
    import torch
    import vsm
    from vsm import _native as N
    from vsm import voxel_map as vm

    dm = vm.DeviceVoxelMap(0.05, 8, N.F32)
    H = np.eye(4)
    pts = torch.rand((129, 2, 2, 3), device="cuda")
    conf = torch.full((129, 2, 2), 2.0, device="cuda")
    emb = torch.ones((129, 2, 2, 8), device="cuda")
    with pytest.raises(ValueError, match="frames in one submap"):
        dm.fuse(pts, conf, emb, dm.make_params(129, 2, 2, 129, 1, 1.0, H, 0, 0))
    st = dm.fuse(pts[:128].contiguous(), conf[:128].contiguous(), emb[:128].contiguous(),
                 dm.make_params(128, 2, 2, 128, 1, 1.0, H, 0, 0))
    assert st["n_fused"] == 128 * 4
    n_before = dm.num_voxels
    far = pts[:2].clone().contiguous()
    far[0, 0, 0, 0] = 1.0e5                                   # 2e6 cells at 5 cm: beyond +-(2^20 - 1)
    with pytest.raises(OverflowError, match="outside"):
        dm.fuse(far, conf[:2].contiguous(), emb[:2].contiguous(), dm.make_params(2, 2, 2, 2, 1, 1.0, H, 1, 0))
    assert dm.num_voxels == n_before                          # the call was stopped before it touched the map
    # coord_range_policy 1: the far point is dropped and counted, the other seven are fused
    N.set_option("coord_range_policy", 1)
    try:
        st = dm.fuse(far, conf[:2].contiguous(), emb[:2].contiguous(), dm.make_params(2, 2, 2, 2, 1, 1.0, H, 1, 0))
    finally:
        N.set_option("coord_range_policy", 0)
    assert st["n_range_dropped"] == 1 and st["n_fused"] == 7
    dm.finalize()
    _, _, counts, _ = dm.export_geometry()
    assert int(counts.sum()) == 128 * 4 + 7
    dm.close()


def test_two_maps_on_two_streams_share_the_workspace_safely():
    """Two maps of one device fused from two streams (queued calls, nothing collected in between): the shared
    workspace is handed over in stream order, so both maps equal their single-stream builds."""
    import torch
    from vsm import _native as N
    from vsm import voxel_map as vm

    subs = [synth.make_submap(91, i, S=4, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.3 * i)
            for i in range(4)]
    dev = [_dev(s) for s in subs]
    thr = [float(vm.conf_threshold(d[1], 25.0)) for d in dev]

    def build(order):
        maps = [vm.DeviceVoxelMap(0.05, 64, N.F32), vm.DeviceVoxelMap(0.05, 64, N.F32)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()] if order == "two_streams" else [None, None]
        torch.cuda.synchronize()
        for i, (s, d) in enumerate(zip(subs, dev)):
            which = i & 1
            p = maps[which].make_params(4, 56, 84, 4, 1, thr[i], s.H_world_map, s.submap_id, N.FUSE_FILTERS)
            if streams[which] is not None:
                with torch.cuda.stream(streams[which]):
                    maps[which].fuse_async(d[0], d[1], d[2], p)
            else:
                maps[which].fuse_async(d[0], d[1], d[2], p)
        out = []
        for which, mp in enumerate(maps):
            if streams[which] is not None:
                with torch.cuda.stream(streams[which]):
                    mp.collect()
                    mp.finalize()
                    out.append((mp.export_packed_keys().cpu().numpy(), mp.export_geometry()[2].cpu().numpy(),
                                mp.features_to_host()))
            else:
                mp.collect()
                mp.finalize()
                out.append((mp.export_packed_keys().cpu().numpy(), mp.export_geometry()[2].cpu().numpy(),
                            mp.features_to_host()))
            mp.close()
        return out

    want = build("one_stream")
    for _ in range(3):
        got = build("two_streams")
        for (k0, c0, f0), (k1, c1, f1) in zip(want, got):
            np.testing.assert_array_equal(k0, k1)
            np.testing.assert_array_equal(c0, c1)
            np.testing.assert_allclose(f0, f1, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("filters", [False, True])
@pytest.mark.parametrize("variant", [5, 13])
def test_table_prefix_too_small_is_repeated_with_the_whole_table(filters, variant):
    """The submap-local tables use a prefix sized from earlier calls (here: forced to its minimum of 16 K slots).  A call
    with more voxels than 3/4 of the prefix is stopped on the device before it touches the map and repeated with the
    whole table: same map as with the prefix sizing off, and the retry is counted."""
    from vsm import _native as N
    from vsm import voxel_map as vm

    s = synth.make_submap(17, 0, S=5, H=56, W=84, d=16, mode="sl4", room=(2.4, 1.8, 1.2))
    thr = vm.conf_threshold(s.conf, 25.0)
    flags = N.FUSE_FILTERS if filters else 0
    outs = []
    N.set_option("prep_variant", variant)
    try:
        for small in (0, 1):
            N.set_option("small_tables", small)
            N.set_option("table_hint", 1)  # prefix = the minimum: 16 384 slots (4 096 for the coarse cells)
            r0 = N.get_counter("table_retries")
            dm = vm.DeviceVoxelMap(0.004, 16, N.F32, capacity=1 << 16)  # 4 mm voxels: nearly one voxel per point
            st = dm.fuse(*_dev(s), dm.make_params(5, 56, 84, 5, 1, thr, s.H_world_map, 0, flags))
            dm.finalize()
            if not filters:  # (with filters on it is the coarse table's prefix that overflows: 12 mm cells of ~1 point)
                assert st["n_submap_voxels"] > 12288
            if not (filters and variant == 5):  # (two tables + filters: 12 mm cells of ~1 point, neither prefix fills up)
                assert (N.get_counter("table_retries") > r0) == bool(small)
            outs.append((st, [t.cpu().numpy() for t in dm.export_geometry()], dm.features_to_host()))
    finally:
        N.set_option("small_tables", 1)
        N.set_option("table_hint", 0)
        N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
    assert repr(outs[0][0]) == repr(outs[1][0])  # (the percentile box is NaN without filters: compare the text)
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-6)
