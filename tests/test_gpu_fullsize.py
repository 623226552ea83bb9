"""Full-size checks (BASELINE.json config shapes) where the numpy oracle would take minutes: the CUDA path against
an independent torch restatement on the same GPU (float64 transform, float32 quantisation, torch.unique,
index_add_ -- the operations of the reference's own torch branch, vggt_slam/map.py:322-348) and against
size-independent properties (count conservation, idempotent means, frame-by-frame == whole-submap)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch_reference(points, conf, emb, thr, H, voxel_size):
    """Per-submap path (submap.py:246-293) with torch on the GPU: keys / counts exact, means in float32."""
    import torch

    mask = conf >= float(thr)
    p = points[mask].double()
    Ht = torch.tensor(H, dtype=torch.float64, device=p.device)
    hom = torch.cat([p, torch.ones_like(p[:, :1])], dim=1)
    # same operation order as the kernel's FMA chain up to the last float64 bit; equal after float32 rounding
    out = hom @ Ht.T
    pw = (out[:, :3] / out[:, 3:]).float()
    coords = torch.floor(pw / torch.tensor(voxel_size, dtype=torch.float32, device=p.device)).to(torch.int64)
    uniq, inverse, counts = torch.unique(coords, dim=0, return_inverse=True, return_counts=True)
    sums = torch.zeros((uniq.shape[0], emb.shape[-1]), dtype=torch.float32, device=p.device)
    e = emb[mask]
    step = 1 << 19
    for i in range(0, e.shape[0], step):
        sums.index_add_(0, inverse[i:i + step], e[i:i + step].float())
    return uniq, counts, sums / counts[:, None].float(), int(mask.sum())


@pytest.mark.parametrize("mode,S,H,W", [("sl4", 32, 294, 518), ("sim3", 16, 518, 518)])
def test_config_shapes_against_torch(mode, S, H, W):
    """config 1 (office_loop shape: 32 x 518x294, SL(4)) and the config-5 frame shape (518x518, Sim(3))."""
    import torch
    from vsm import _native as N
    from vsm import synth_device
    from vsm import voxel_map as vm

    d = synth_device.make_submap_device(77, 0, S=S, H=H, W=W, d=512, mode=mode, room=(6.0, 4.0, 3.0))
    thr = vm.conf_threshold(d.conf, 25.0)
    assert float(thr) == float(np.percentile(d.conf.cpu().numpy(), 25.0))     # a1 at full size
    uniq, counts, mean, n_sel = _torch_reference(d.points, d.conf, d.emb, thr, d.H_world_map, 0.05)
    dm = vm.DeviceVoxelMap(0.05, 512, N.BF16, capacity=1 << 17)
    p = dm.make_params(S, H, W, S, 1, thr, d.H_world_map, 0, 0)
    st = dm.fuse(d.points, d.conf, d.emb, p)
    dm.finalize()
    assert st["n_fused"] == n_sel == st["n_conf"]
    coords, _, cnt, _ = dm.export_geometry()
    assert torch.equal(coords, uniq)                 # voxel key set and order: bit-exact
    assert torch.equal(cnt, counts)                  # per-voxel counts: bit-exact
    assert int(cnt.sum()) == n_sel                   # conservation
    feats = dm.export_features()
    torch.testing.assert_close(feats, mean, rtol=1e-3, atol=1e-5)
    # idempotent means: fusing the same submap again doubles the counts and leaves the means alone
    dm.fuse(d.points, d.conf, d.emb, p)
    dm.finalize()
    _, _, cnt2, _ = dm.export_geometry()
    assert torch.equal(cnt2, 2 * counts)
    torch.testing.assert_close(dm.export_features(), feats, rtol=1e-5, atol=1e-6)
    # the streaming (pixel-order) kernel gives the same map
    dm2 = vm.DeviceVoxelMap(0.05, 512, N.BF16, capacity=1 << 17)
    dm2.fuse(d.points, d.conf, d.emb, dm2.make_params(S, H, W, S, 1, thr, d.H_world_map, 0, N.FUSE_PIXEL_ORDER))
    dm2.finalize()
    c2, _, n2, _ = dm2.export_geometry()
    assert torch.equal(c2, uniq) and torch.equal(n2, counts)
    torch.testing.assert_close(dm2.export_features(), feats, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mode,variant", [(1, 5), (2, 5), (1, 13), (3, 13)],
                         ids=["radix_select", "bracket_select", "one_table_radix", "one_table_deferred_box"])
def test_filter_bounds_at_config_size(mode, variant):
    """The bbox filter's 0.5 / 99.5 percentiles (map.py:257-258) at config-1 size against np.percentile on the same
    world points, bit-exact, for both select implementations; the one-pass bracket select must answer without
    falling back; both give the same map."""
    import torch
    from vsm import _native as N
    from vsm import synth_device
    from vsm import voxel_map as vm

    S, H, W = 32, 294, 518
    d = synth_device.make_submap_device(79, 0, S=S, H=H, W=W, d=64, mode="sl4", room=(6.0, 4.0, 3.0))
    thr = vm.conf_threshold(d.conf, 25.0)
    pw = vm.transform_points(d.points.reshape(-1, 3), d.H_world_map, out_f64=False).cpu().numpy()
    keep = (d.conf.reshape(-1) >= float(thr)).cpu().numpy() & np.isfinite(pw).all(axis=1)
    want_lo = np.percentile(pw[keep], 0.5, axis=0)
    want_hi = np.percentile(pw[keep], 99.5, axis=0)
    assert want_lo.dtype == np.float32
    N.set_option("select_mode", mode)
    N.set_option("prep_variant", variant)
    try:
        misses0 = N.get_counter("select_misses")
        dm = vm.DeviceVoxelMap(0.05, 64, N.BF16, capacity=1 << 17)
        st = dm.fuse(d.points, d.conf, d.emb, dm.make_params(S, H, W, S, 1, thr, d.H_world_map, 0, N.FUSE_FILTERS))
        assert N.get_counter("select_misses") == misses0
    finally:
        N.set_option("select_mode", 0)
        N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
    np.testing.assert_array_equal(np.asarray(st["bbox_lo"], np.float32), want_lo)
    np.testing.assert_array_equal(np.asarray(st["bbox_hi"], np.float32), want_hi)
    inside = keep & (pw >= want_lo).all(axis=1) & (pw <= want_hi).all(axis=1)
    assert st["n_finite"] == int(keep.sum()) and st["n_bbox"] == int(inside.sum())
    dm.finalize()
    coords, _, cnt, _ = dm.export_geometry()
    assert int(cnt.sum()) == st["n_fused"] <= st["n_bbox"]
    # filter 3 and the voxel keys restated independently (map.py:271-280, 351): cells of float32(3 vs) with >= 10 points
    pin = pw[inside]
    ck = np.floor(pin / np.float32(0.05 * 3.0)).astype(np.int64)
    _, inv, n_cell = np.unique(ck, axis=0, return_inverse=True, return_counts=True)
    ok = n_cell[inv.reshape(-1)] >= 10
    assert st["n_fused"] == int(ok.sum())
    vk, vn = np.unique(np.floor(pin[ok] / np.float32(0.05)).astype(np.int64), axis=0, return_counts=True)
    np.testing.assert_array_equal(coords.cpu().numpy(), vk)
    np.testing.assert_array_equal(cnt.cpu().numpy(), vn)
    assert st["n_submap_voxels"] == vk.shape[0]
    dm.close()


def test_frame_by_frame_equals_whole_submap():
    """Config-5 style per-frame streaming fusion: S=1 calls with frame_base == one call over the submap
    (filters off: they are per call), including the contributor frame masks."""
    import torch
    from vsm import _native as N
    from vsm import synth_device
    from vsm import voxel_map as vm

    S, H, W = 6, 518, 518
    d = synth_device.make_submap_device(78, 3, S=S, H=H, W=W, d=512, mode="sim3", room=(6.0, 4.0, 3.0))
    thr = vm.conf_threshold(d.conf, 25.0)
    whole = vm.DeviceVoxelMap(0.05, 512, N.BF16)
    whole.fuse(d.points, d.conf, d.emb, whole.make_params(S, H, W, S, 1, thr, d.H_world_map, 3, 0))
    whole.finalize()
    stream = vm.DeviceVoxelMap(0.05, 512, N.BF16)
    for f in range(S):
        stream.fuse_async(d.points[f:f + 1], d.conf[f:f + 1], d.emb[f:f + 1],
                          stream.make_params(1, H, W, 1, 1, thr, d.H_world_map, 3, 0, frame_base=f))
    stats = stream.collect()
    assert len(stats) == S and sum(s["n_fused"] for s in stats) == whole.last_stats["n_fused"]
    stream.finalize()
    a, b = whole.export_geometry(), stream.export_geometry()
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    torch.testing.assert_close(stream.export_features(), whole.export_features(), rtol=1e-4, atol=1e-6)
    # contributors: one CSR entry per (call, voxel); OR-ing a voxel's masks gives the whole-submap mask
    off_w, sub_w, mask_w = whole.export_contributors()
    off_s, sub_s, mask_s = stream.export_contributors()
    assert (sub_w == 3).all() and (sub_s == 3).all() and (np.diff(off_w) == 1).all()
    merged = np.zeros_like(mask_w)
    np.bitwise_or.at(merged, np.repeat(np.arange(len(off_s) - 1), np.diff(off_s)), mask_s)
    np.testing.assert_array_equal(merged, mask_w)
