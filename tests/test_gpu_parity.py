"""GPU parity tests: the CUDA path (through the reference-facing classes and the C ABI) against
  (1) the golden vectors produced by running the unmodified reference (tests/golden/*.npz), and
  (2) the numpy oracle on seeded inputs.
Bars: voxel key sets (== centres), per-voxel counts, point->voxel indices, top-k indices: bit-exact;
features and scores: 1e-3 relative (fp32 accumulate), tolerance written at each assert."""
import json
import os

import numpy as np
import pytest

import golden_io as gio
from oracle import voxel_oracle as vo
from vsm import synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-5


@pytest.fixture(scope="module")
def vsm_mod():
    import torch

    assert torch.cuda.is_available()
    import vsm

    return vsm


def to_submap(vsm, s: synth.SynthSubmap, emb=None, device_inputs=False):
    import torch

    sm = vsm.Submap(s.submap_id)
    pts, conf = s.points, s.conf
    if device_inputs:
        pts, conf = torch.from_numpy(pts).cuda(), torch.from_numpy(conf).cuda()
    sm.add_all_points(pts, s.colors, conf, s.conf_percentile, None)
    sm.add_all_semantic_embeddings(s.emb if emb is None else emb)
    sm.set_conf_masks(s.conf)
    sm.set_reference_homography(s.H_world_map)
    sm.set_frame_ids(s.frame_paths)
    sm.set_last_non_loop_frame_index(s.last_non_loop_frame_index)
    return sm


def bf16_tensor(emb_f32: np.ndarray, device="cuda"):
    import torch

    bits = synth.f32_to_bf16_bits(emb_f32).view(np.int16)
    return torch.from_numpy(bits).to(device).view(torch.bfloat16)


# ---------------------------------------------------------------------------
# a1: confidence threshold
# ---------------------------------------------------------------------------
def test_conf_threshold_golden(vsm_mod):
    from vsm import voxel_map as vm

    z = gio.load("case_a_submap_sl4.npz")
    s = gio.inputs(z)[0]
    thr = vm.conf_threshold(s.conf, s.conf_percentile)
    assert thr.dtype == np.float32 and thr == z["conf_threshold"]
    zd = gio.load("case_d_percentile.npz")
    for n, q, want in zd["cases"]:
        got = vm.conf_threshold(zd[f"x_{int(n)}"], float(q))
        assert float(got) == float(want), (n, q, got, want)


def test_conf_threshold_random_and_nan(vsm_mod):
    from vsm import voxel_map as vm

    rng = np.random.default_rng(3)
    for n in (7, 1025, 300001, 2_000_003):
        x = (rng.normal(size=n) * 3).astype(np.float32)
        x[rng.integers(0, n, size=3)] = 0.0
        x[rng.integers(0, n, size=2)] = -0.0
        for q in (0.5, 25.0, 99.5, 100.0, 0.0, 61.7):
            assert float(vm.conf_threshold(x, q)) == float(np.percentile(x, q)), (n, q)
    x = rng.normal(size=1000).astype(np.float32)
    x[17] = np.nan
    assert np.isnan(vm.conf_threshold(x, 25.0))


# ---------------------------------------------------------------------------
# a4 / a5: world-frame points
# ---------------------------------------------------------------------------
def test_world_points_golden(vsm_mod):
    z = gio.load("case_a_submap_sl4.npz")
    s = gio.inputs(z)[0]
    sm = to_submap(vsm_mod, s)
    for stride in (1, 2):
        got = sm.get_points_in_world_frame(stride)
        want = z[f"world_points_s{stride}"]
        assert got.dtype == np.float64 and got.shape == want.shape
        # float64 FMA chain vs BLAS dgemm: equal to the last bit or two
        np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)
        np.testing.assert_array_equal(got.astype(np.float32), want.astype(np.float32))
    np.testing.assert_array_equal(sm.get_points_colors(1), z["colors_s1"])
    np.testing.assert_array_equal(sm.get_points_colors(3), z["colors_s3"])
    pl, fl, ml = sm.get_points_list_in_world_frame(ignore_loop_closure_frames=True)
    np.testing.assert_allclose(np.stack(pl), z["plist_points"], rtol=1e-15, atol=0)
    np.testing.assert_array_equal(np.asarray(fl), z["plist_ids"])
    np.testing.assert_array_equal(np.stack(ml), z["plist_masks"])
    with pytest.raises(IndexError):  # loop frame has no frame id upstream either
        sm.get_points_list_in_world_frame(ignore_loop_closure_frames=False)


# ---------------------------------------------------------------------------
# a6: per-submap fusion
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("tag,ign", [("all", False), ("noloop", True)])
@pytest.mark.parametrize("vs", [0.1, 0.05])
def test_fuse_submap_golden(vsm_mod, tag, ign, vs, dtype):
    z = gio.load("case_a_submap_sl4.npz")
    s = gio.inputs(z)[0]
    sm = to_submap(vsm_mod, s, emb=bf16_tensor(s.emb) if dtype == "bf16" else None)
    v = sm.get_semantic_voxel_in_world_frame(vs, ignore_loop_closure_frames=ign)
    k = f"{tag}_{vs}"
    np.testing.assert_array_equal(v.centers_world, z[f"{k}_centers"])  # key set and order: bit-exact
    np.testing.assert_allclose(v.features, z[f"{k}_features"], rtol=RTOL, atol=ATOL)
    assert v.contributors == gio.contributors(z, f"{k}_contrib")


def test_fuse_submap_nonfinite_quirk(vsm_mod):
    """No filters on the per-submap path: NaN/Inf points land in the INT64_MIN voxel (submap.py:279-292)."""
    z = gio.load("case_b_submap_sim3_bad.npz")
    s = gio.inputs(z)[0]
    sm = to_submap(vsm_mod, s)
    v = sm.get_semantic_voxel_in_world_frame(0.05)
    np.testing.assert_array_equal(v.centers_world, z["centers"])
    want = z["features"]
    got = v.features
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(np.isinf(got), np.isinf(want))
    ok = np.isfinite(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=RTOL, atol=ATOL)
    assert v.contributors == gio.contributors(z, "contrib")


def test_fuse_submap_errors(vsm_mod):
    z = gio.load("case_a_submap_sl4.npz")
    s = gio.inputs(z)[0]
    sm = to_submap(vsm_mod, s)
    with pytest.raises(ValueError):
        sm.get_semantic_voxel_in_world_frame(0.0)
    sm2 = vsm_mod.Submap(1)
    with pytest.raises(RuntimeError):
        sm2.get_semantic_voxel_in_world_frame(0.05)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings([1, 2, 3])
    with pytest.raises(ValueError):
        sm.add_all_semantic_embeddings(np.zeros((2, 2, 2), np.float32))
    with pytest.raises(ValueError):
        sm.add_all_semantic_embeddings(np.zeros((1, 2, 2, 8), np.float32))
    # nothing passes the threshold -> empty result with (0,d) features (submap.py:260-266)
    sm.conf_threshold = np.float32(1e9)
    v = sm.get_semantic_voxel_in_world_frame(0.05)
    assert v.centers_world.shape == (0, 3) and v.features.shape == (0, s.emb.shape[-1]) and v.contributors == []


@pytest.mark.parametrize("seed,S,H,W,d,mode,kind", [
    (21, 6, 56, 98, 64, "sl4", "normal"),
    (22, 5, 42, 70, 512, "sim3", "painted"),
    (23, 3, 29, 41, 8, "se3", "normal"),   # odd sizes: unaligned tails
])
def test_fuse_submap_oracle(vsm_mod, seed, S, H, W, d, mode, kind, kernel_variant):
    import torch

    s = synth.make_submap(seed, 2, S=S, H=H, W=W, d=d, mode=mode, emb_kind=kind, room=(3.0, 2.5, 2.0))
    want = vo.fuse_submap(gio.to_oracle_submap(s), 0.05, exact_order=False, with_contributors=False)
    for dtype in ("f32", "bf16"):
        sm = to_submap(vsm_mod, s, emb=bf16_tensor(s.emb) if dtype == "bf16" else None, device_inputs=True)
        v = sm.get_semantic_voxel_in_world_frame(0.05)
        np.testing.assert_array_equal(v.centers_world, want.centers_world)
        np.testing.assert_allclose(v.features, want.features, rtol=RTOL, atol=ATOL)
        dm = v._device_map
        coords, _, counts, _ = dm.export_geometry()
        np.testing.assert_array_equal(coords.cpu().numpy(), want.coords)        # true keys: bit-exact
        np.testing.assert_array_equal(counts.cpu().numpy(), want.counts)        # counts: bit-exact
        inv = dm.export_point_index(0, S * H * W).cpu().numpy()
        mask = s.conf >= vo.conf_threshold(s.conf, s.conf_percentile)
        keys = vo.voxel_keys(vo.world_points_f32(s.points[mask], s.H_world_map), 0.05)
        _, want_inv = np.unique(keys, axis=0, return_inverse=True)
        np.testing.assert_array_equal(inv[mask.reshape(-1)], want_inv.reshape(-1))  # np.unique's inverse
        assert (inv[~mask.reshape(-1)] == -1).all()
        del v, dm
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------
# a7: global build with the three filters
# ---------------------------------------------------------------------------
@pytest.fixture(params=[1, 2, 3], ids=["radix_select", "bracket_select", "deferred_box"])
def select_mode(request):
    """Both implementations of the bbox percentiles (three-pass radix select / one-pass bracket select with its
    retry path) must give the reference's bounds."""
    from vsm import _native as N

    N.set_option("select_mode", request.param)
    yield request.param
    N.set_option("select_mode", 0)


@pytest.fixture(params=[0, 1], ids=["one_stream", "overlap"])
def fuse_overlap(request):
    """vsm_set_option("overlap", 1): the accumulate kernel of a call runs on a side stream beside the next call's
    preparation kernels (double-buffered entry lists) -- same map."""
    from vsm import _native as N

    N.set_option("overlap", request.param)
    yield request.param
    N.set_option("overlap", 0)


@pytest.fixture(params=[0, 5, 7, 13, 8], ids=["row_kernels", "two_table_kernels", "mask_probe", "one_table_kernels", "one_table_rows"])
def kernel_variant(request):
    """The preparation kernels with a warp per 32-pixel row run / per 4x8 patch (+ merged select steps, + frame-mask
    probe): every combination must give the reference's map."""
    from vsm import _native as N

    N.set_option("prep_variant", request.param)
    yield request.param
    N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)


def graph_from(vsm, subs, **kw):
    gm = vsm.GraphMap()
    for s in subs:
        gm.add_submap(to_submap(vsm, s, **kw))
    return gm


@pytest.mark.parametrize("tag,kw", [
    ("s1_dedup", dict(stride=1)),
    ("s2_dedup", dict(stride=2)),
    ("s1_nodedup", dict(stride=1, deduplicate_contributors=False)),
])
@pytest.mark.parametrize("streaming", [False, True])
def test_build_global_golden(vsm_mod, tag, kw, streaming, select_mode, kernel_variant):
    z = gio.load("case_c_global_sl4.npz")
    gm = graph_from(vsm_mod, gio.inputs(z))
    m = gm.build_semantic_voxel_map(0.05, host_streaming=streaming, **kw)
    np.testing.assert_array_equal(m.get_centers_world(), z[f"{tag}_centers"])
    np.testing.assert_allclose(m.get_features(), z[f"{tag}_features"], rtol=RTOL, atol=ATOL)
    assert m.get_contributors() == gio.contributors(z, f"{tag}_contrib")
    assert m.frame_name_maps == json.loads(str(z[f"{tag}_names"]))
    np.testing.assert_array_equal(m._voxel_coords, z[f"{tag}_recon_coords"])


def test_build_global_filter_stages(vsm_mod, select_mode):
    """Survivor counts and the percentile box of every filter stage against the oracle."""
    z = gio.load("case_c_global_sl4.npz")
    subs = gio.inputs(z)
    gm = graph_from(vsm_mod, subs)
    gm.build_semantic_voxel_map(0.05)
    # the goldens carry NaN embeddings, so the build ran twice; the stats kept are those of the exact pass
    assert len(gm.last_build_stats) == len(subs)
    for s, st in zip(subs, gm.last_build_stats):
        stages = {}
        with np.errstate(all="ignore"):
            vo.submap_observations(gio.to_oracle_submap(s), 0.05, 1, True, stages)
        assert st["n_conf"] == stages["n_conf"]
        assert st["n_finite"] == stages["n_finite"]
        assert st["n_bbox"] == stages["n_bbox"]
        assert st["n_fused"] == stages["n_coarse"]
        np.testing.assert_array_equal(np.asarray(st["bbox_lo"], np.float32), stages["lo"])
        np.testing.assert_array_equal(np.asarray(st["bbox_hi"], np.float32), stages["hi"])


@pytest.mark.parametrize("voxel_size", [0.05, 0.02, 0.1, 0.3])
@pytest.mark.parametrize("variant", [13, 5])
def test_points_on_coarse_cell_boundaries(vsm_mod, voxel_size, variant, select_mode):
    """The coarse-cell filter (map.py:271-280) on points placed within a few float32 steps of the cell boundaries
    k * float32(3 * vs): float32(vs * 3) is not 3 * float32(vs), so floor(p / cell) and floor(floor(p / vs) / 3)
    disagree for some of them -- the one-table preparation derives a voxel's cell from its coordinates and must find
    exactly those points.  A dense cloud keeps most cells above the 10-point limit, a thin shell leaves sparse cells
    next to them, so both outcomes of the filter meet boundary points.  Keys, counts and filter stages: bit-exact."""
    from vsm import _native as N

    rng = np.random.default_rng(int(voxel_size * 1000) + 7)
    cell = np.float32(voxel_size * 3.0)
    S, H, W, d = 2, 96, 128, 8
    n = S * H * W
    pts = np.empty((n, 3), dtype=np.float32)
    half = n // 2
    # boundary points: coordinates k * cell moved by -3 .. +3 float32 steps, on every axis independently
    k = rng.integers(-60, 60, size=(half, 3)).astype(np.float64)
    b = (k * np.float64(cell)).astype(np.float32)
    steps = rng.integers(-3, 4, size=(half, 3))
    for _ in range(3):
        up = np.nextafter(b, np.float32(np.inf))
        dn = np.nextafter(b, np.float32(-np.inf))
        b = np.where(steps > 0, up, np.where(steps < 0, dn, b))
        steps = steps - np.sign(steps)
    on_boundary = rng.random(size=(half, 3)) < 0.5  # the other coordinates anywhere in the box
    box = 60.0 * float(cell)
    pts[:half] = np.where(on_boundary, b, rng.uniform(-box, box, size=(half, 3)).astype(np.float32))
    # dense cloud in a small box around the origin (cells well above 10 points) + nothing else: the boundary points
    # far from it sit in sparse cells
    pts[half:] = rng.uniform(-4.0 * float(cell), 4.0 * float(cell), size=(n - half, 3)).astype(np.float32)
    pts[:half // 2] = np.clip(pts[:half // 2], -4.0 * float(cell), 4.0 * float(cell))  # half of the boundary points inside it
    rng.shuffle(pts, axis=0)
    s = synth.SynthSubmap(0, pts.reshape(S, H, W, 3), (1.0 + rng.gamma(2.0, 2.0, size=(S, H, W))).astype(np.float32),
                          np.zeros((S, H, W, 3), np.uint8), synth.round_to_bf16(rng.normal(size=(S, H, W, d)).astype(np.float32)),
                          np.eye(4), [f"f_{i:03d}.png" for i in range(S)], S - 1)
    stages = {}
    with np.errstate(all="ignore"):
        want = vo.build_global([gio.to_oracle_submap(s)], voxel_size, exact_order=False)
        vo.submap_observations(gio.to_oracle_submap(s), voxel_size, 1, True, stages)
    # the case is adversarial: for some points the reference's cell is not the cell of the point's voxel
    fine = vo.voxel_keys(pts, float(voxel_size))
    assert int((vo.voxel_keys(pts, float(voxel_size) * 3.0) != np.floor_divide(fine, 3)).any(axis=1).sum()) > 20
    N.set_option("prep_variant", variant)
    try:
        gm = graph_from(vsm_mod, [s], device_inputs=True)
        m = gm.build_semantic_voxel_map(voxel_size)
    finally:
        N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
    st = gm.last_build_stats[0]
    assert st["n_bbox"] == stages["n_bbox"] and st["n_fused"] == stages["n_coarse"]
    assert 0 < stages["n_coarse"] < stages["n_bbox"]
    coords, _, counts, _ = m._dm.export_geometry()
    np.testing.assert_array_equal(coords.cpu().numpy(), want.coords)
    np.testing.assert_array_equal(counts.cpu().numpy(), want.counts)
    assert st["n_submap_voxels"] == want.coords.shape[0]
    np.testing.assert_allclose(m.get_features(), want.features, rtol=RTOL, atol=ATOL)
    assert m.get_contributors() == want.contributors


def test_build_global_errors_and_empty(vsm_mod):
    z = gio.load("case_c_global_sl4.npz")
    subs = gio.inputs(z)
    gm = graph_from(vsm_mod, subs)
    assert str(z["s1_withloop_error"]) == "IndexError"
    with pytest.raises(IndexError):
        gm.build_semantic_voxel_map(0.05, ignore_loop_closure_frames=False)
    with pytest.raises(ValueError):
        gm.build_semantic_voxel_map(0.0)
    with pytest.raises(ValueError):
        gm.build_semantic_voxel_map(0.05, stride=0)
    empty = vsm_mod.GraphMap().build_semantic_voxel_map(0.05)
    assert empty.get_centers_world().shape == (0, 3) and empty.get_features().shape == (0, 0)
    assert empty.get_contributors() == [] and empty.frame_name_maps == {}


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("stride", [1, 3])
def test_build_global_oracle(vsm_mod, dtype, stride, select_mode, fuse_overlap, kernel_variant):
    """Four overlapping submaps, SL(4), d=64, clean embeddings (single optimistic pass), device inputs."""
    subs = [synth.make_submap(31, i, S=5, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=5 * i, n_loop_frames=(1 if i == 2 else 0)) for i in range(4)]
    with np.errstate(all="ignore"):
        want = vo.build_global([gio.to_oracle_submap(s) for s in subs], 0.05, stride=stride, exact_order=False)
    gm = vsm_mod.GraphMap()
    for s in subs:
        gm.add_submap(to_submap(vsm_mod, s, emb=bf16_tensor(s.emb) if dtype == "bf16" else None, device_inputs=True))
    m = gm.build_semantic_voxel_map(0.05, stride=stride)
    np.testing.assert_array_equal(m.get_centers_world(), want.centers_world)
    np.testing.assert_allclose(m.get_features(), want.features, rtol=RTOL, atol=ATOL)
    coords, _, counts, _ = m._dm.export_geometry()
    np.testing.assert_array_equal(coords.cpu().numpy(), want.coords)
    np.testing.assert_array_equal(counts.cpu().numpy(), want.counts)
    assert m.get_contributors() == want.contributors
    assert sum(st["n_fused"] for st in gm.last_build_stats) == want.n_points == int(want.counts.sum())


@pytest.mark.parametrize("voxel_size", [0.012, 0.05, 0.4])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_segment_shapes_against_oracle(vsm_mod, voxel_size, dtype, kernel_variant):
    """Voxel segments of ~1 row (1.2 cm voxels), tens of rows and thousands of rows (40 cm voxels: segments that run
    through several 32-entry chunks of the sorted list), d = 256 (full rows: the hot loop of the accumulate kernel)."""
    subs = [synth.make_submap(47, i, S=3, H=56, W=84, d=256, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=3 * i) for i in range(3)]
    with np.errstate(all="ignore"):
        want = vo.build_global([gio.to_oracle_submap(s) for s in subs], voxel_size, exact_order=False)
    gm = vsm_mod.GraphMap()
    for s in subs:
        gm.add_submap(to_submap(vsm_mod, s, emb=bf16_tensor(s.emb) if dtype == "bf16" else None, device_inputs=True))
    m = gm.build_semantic_voxel_map(voxel_size)
    np.testing.assert_array_equal(m.get_centers_world(), want.centers_world)
    coords, _, counts, _ = m._dm.export_geometry()
    np.testing.assert_array_equal(counts.cpu().numpy(), want.counts)
    np.testing.assert_allclose(m.get_features(), want.features, rtol=RTOL, atol=ATOL)
    assert m.get_contributors() == want.contributors


def test_build_orders_agree(vsm_mod):
    """Voxel-sorted accumulate, pixel-order accumulate and host streaming give the same map."""
    import torch
    from vsm import _native as N
    from vsm import voxel_map as vm

    s = synth.make_submap(41, 0, S=4, H=56, W=84, d=128, mode="sl4", room=(2.4, 1.8, 1.2))
    thr = vm.conf_threshold(s.conf, 25.0)
    pts, conf, emb = (torch.from_numpy(x).cuda() for x in (s.points, s.conf, s.emb))
    outs = []
    # "host": pageable numpy arrays, staged frame by frame; "host_pinned": pinned memory, which the accumulate kernel
    # reads in place over PCIe (only the rows of selected pixels); "host_pinned_staged": the same with the option off
    pinned = torch.from_numpy(s.emb).pin_memory()
    for mode in ("sorted", "pixel", "host", "host_pinned", "host_pinned_staged"):
        dm = vm.DeviceVoxelMap(0.05, 128, N.F32)
        flags = N.FUSE_FILTERS | (N.FUSE_PIXEL_ORDER if mode == "pixel" else 0)
        p = dm.make_params(4, 56, 84, 4, 1, thr, s.H_world_map, 0, flags)
        if mode.startswith("host"):
            N.set_option("host_zero_copy", 0 if mode.endswith("staged") else 1)
            try:
                st = dm.fuse_host(s.points, s.conf, s.emb if mode == "host" else pinned.numpy(), p)
            finally:
                N.set_option("host_zero_copy", 1)
        else:
            st = dm.fuse(pts, conf, emb, p)
        dm.finalize()
        coords, _, counts, _ = dm.export_geometry()
        outs.append((coords.cpu().numpy(), counts.cpu().numpy(), dm.features_to_host(), st["n_fused"]))
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[1], outs[0][1])
        np.testing.assert_allclose(o[2], outs[0][2], rtol=1e-5, atol=1e-6)
        assert o[3] == outs[0][3]


# ---------------------------------------------------------------------------
# a11-a15: lookup, query, latest frame, persistence
# ---------------------------------------------------------------------------
def test_query_lookup_latest_golden(vsm_mod):
    z = gio.load("case_c_global_sl4.npz")
    gm = graph_from(vsm_mod, gio.inputs(z))
    m = gm.build_semantic_voxel_map(0.05)
    Q = z["q"]
    for k in (1, 5):
        for p in range(Q.shape[0]):
            qi = Q[p] if p % 2 == 0 else Q[p][None, :]
            idx, coords, sims = m.query_with_embedding(qi, top_k=k)
            assert idx == z[f"q_k{k}_idx"][p].tolist()                       # top-k indices: bit-exact
            np.testing.assert_array_equal(coords, z[f"q_k{k}_coords"][p])
            np.testing.assert_allclose(sims, z[f"q_k{k}_sims"][p], rtol=RTOL, atol=1e-6)
        bi, bc, bs = m.query_with_embeddings(Q, top_k=k)                      # batched == looped
        np.testing.assert_array_equal(bi, z[f"q_k{k}_idx"])
        np.testing.assert_array_equal(bc, z[f"q_k{k}_coords"])
        np.testing.assert_allclose(bs, z[f"q_k{k}_sims"], rtol=RTOL, atol=1e-6)
    lat = json.loads(str(z["latest_every7"]))
    for j, i in enumerate(range(0, m.get_features().shape[0], 7)):
        name, sid, fid = m.get_latest_frame_at_voxel(i)
        assert [name, sid, fid] == lat[j]
    got = m.get_indices_at_positions(z["probe_pos"])
    np.testing.assert_array_equal(got, z["probe_idx"])                        # reference dict semantics (compat)
    one = m.get_index_at_position(z["probe_pos"][0])
    assert (one if one is not None else -1) == int(z["probe_idx"][0])
    # exact-key lookup: every voxel centre finds its own voxel
    ex = m._dm.lookup(m.get_centers_world(), compat=False)
    np.testing.assert_array_equal(ex, np.arange(len(ex)))
    with pytest.raises(RuntimeError):
        m.query_with_embedding(Q[0], top_k=len(ex) + 1)


def test_query_normalised_and_large_k(vsm_mod):
    subs = [synth.make_submap(51, i, S=4, H=56, W=84, d=64, mode="sim3", room=(2.4, 1.8, 1.2), start=0.3 * i)
            for i in range(2)]
    gm = graph_from(vsm_mod, subs)
    m = gm.build_semantic_voxel_map(0.05)
    F = m.get_features()
    rng = np.random.default_rng(1)
    Q = rng.normal(size=(11, 64)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    for k in (1, 10, 300):
        bi, _, bs = m.query_with_embeddings(Q, top_k=k)
        for p in range(Q.shape[0]):
            widx, wsims, allsims = vo.query(F, Q[p], k)
            assert bi[p].tolist() == widx
            np.testing.assert_allclose(bs[p], wsims, rtol=RTOL, atol=1e-6)
        bi, _, bs = m.query_with_embeddings(Q, top_k=k, normalize=True)
        for p in range(Q.shape[0]):
            widx, wsims, _ = vo.query_normalised(F, Q[p], k)
            assert bi[p].tolist() == widx
            np.testing.assert_allclose(bs[p], wsims, rtol=RTOL, atol=1e-6)


def test_persistence_round_trip(vsm_mod, tmp_path):
    z = gio.load("case_c_global_sl4.npz")
    gm = graph_from(vsm_mod, gio.inputs(z))
    m = gm.build_semantic_voxel_map(0.05, stride=2)
    # a map saved by the reference loads here and answers like the reference
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    (ref_dir / "semantic_voxels.npz").write_bytes(z["saved_npz_bytes"].tobytes())
    (ref_dir / "frame_names.json").write_text(str(z["saved_json"]))
    lm = vsm_mod.SemanticVoxelMap.load_from_directory(str(ref_dir))
    np.testing.assert_array_equal(lm.get_centers_world(), m.get_centers_world())
    np.testing.assert_allclose(lm.get_features(), m.get_features(), rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(lm._voxel_coords, m._voxel_coords)
    assert lm.frame_name_maps == m.frame_name_maps
    q = z["q"][0]
    assert lm.query_with_embedding(q, top_k=5)[0] == m.query_with_embedding(q, top_k=5)[0]
    # and a map saved here has the reference's file layout
    out = tmp_path / "ours"
    m.save_to_directory(str(out))
    a = np.load(out / "semantic_voxels.npz", allow_pickle=True)
    b = np.load(ref_dir / "semantic_voxels.npz", allow_pickle=True)
    assert sorted(a.files) == sorted(b.files)
    assert a["voxel_size"].dtype == np.float32 and a["voxel_size"] == b["voxel_size"]
    np.testing.assert_array_equal(a["centers_world"], b["centers_world"])
    assert a["features"].dtype == np.float32
    np.testing.assert_allclose(a["features"], b["features"], rtol=RTOL, atol=ATOL)
    assert [list(map(tuple, c)) for c in a["contributors"].tolist()] == \
        [list(map(tuple, c)) for c in b["contributors"].tolist()]
    assert json.loads((out / "frame_names.json").read_text()) == json.loads(str(z["saved_json"]))
    lm2 = vsm_mod.SemanticVoxelMap.load_from_directory(str(out))
    assert lm2.query_with_embedding(q, top_k=5)[0] == m.query_with_embedding(q, top_k=5)[0]
    name, sid, fid = lm2.get_latest_frame_at_voxel(3)
    assert (name, sid, fid) == m.get_latest_frame_at_voxel(3)


# ---------------------------------------------------------------------------
# a16-a18: evaluators, evaluation manager, query CLI
# ---------------------------------------------------------------------------
def _saved_map(vsm, tmp_path):
    z = gio.load("case_c_global_sl4.npz")
    gm = graph_from(vsm, gio.inputs(z))
    m = gm.build_semantic_voxel_map(0.05)
    d = tmp_path / "voxels"
    m.save_to_directory(str(d))
    return m, str(d)


def test_evaluators_and_manager(vsm_mod, tmp_path):
    from vsm import voxel_evaluation_manager as mgr
    from vsm import voxel_evaluators as ev

    m, vdir = _saved_map(vsm_mod, tmp_path)
    d = m.get_features().shape[1]
    rng = np.random.default_rng(7)
    table = {}

    def encoder(texts):  # stands in for the CLIP text tower (weights are not available offline)
        out = []
        for t in texts:
            if t not in table:
                table[t] = rng.normal(size=d).astype(np.float32) * 3.0
            out.append(table[t])
        return np.stack(out)

    # the frame the map retrieves for "chair" is annotated as a chair (valid) but not as a "lamp"
    q = encoder(["chair"])[0]
    q = q / np.linalg.norm(q)
    idx, _, sims = m.query_with_embedding(q, top_k=3)
    frame, sid, fid = m.get_latest_frame_at_voxel(idx[0])
    ts = ev.get_ts(frame)
    assert ts is not None
    ann = tmp_path / "annotations.json"
    ann.write_text(json.dumps({"images": [{"file_name": frame, "label": "office chair"},
                                          {"timestamp": ts + 10**12, "label": "floor lamp"}]}))
    sve = ev.SearchValidityEvaluator(str(ann), text_encoder=encoder)
    res = sve.evaluate(m, {"queries": ["chair", "lamp"], "top_k": 3})
    assert res[0]["found"] and res[0]["valid"] and res[0]["retrieved_voxel_index"] == idx[0]
    assert res[0]["retrieved_img"] == frame and res[0]["retrieved_submap_id"] == sid and res[0]["retrieved_frame_id"] == fid
    assert abs(res[0]["score"] - sims[0]) <= 1e-6 * max(1.0, abs(sims[0])) and res[0]["time_diff_ns"] == 0.0
    assert res[1]["found"] and not res[1]["valid"] and res[1]["closest_gt_label"] == "floor lamp"
    assert sve.evaluate(m, {}) is None
    cnt = ev.get_evaluator("voxel_count_metric", {}).evaluate(m, {})
    assert cnt == {"num_voxels": len(m.get_centers_world()), "feature_dim": d, "voxel_size": 0.05}
    perf = ev.get_evaluator("performance_metric", {}).evaluate(m, {})
    assert perf["status"] == "ok" and perf["queries"]["1"]["ms"] > 0
    with pytest.raises(NotImplementedError):
        ev.get_evaluator("navigability_metric", {}).evaluate(m, {})

    config = {"experiment_name": "unit test!", "datasets": [{"path": str(tmp_path), "annotation_file": str(ann),
                                                            "voxel_map_dir": vdir + " "}],
              "hyperparameters": [{"queries": ["chair", "lamp"], "top_k": 2}, {"queries": "chair"}],
              "eval_functions": ["search_validity_metric", "voxel_count_metric", "navigability_metric"]}
    out_file, runs = mgr.run(config, workers=1, out_dir=str(tmp_path / "results"), text_encoder=encoder)
    assert os.path.basename(out_file) == "unittest.json" and len(runs) == 2
    saved = json.loads(open(out_file).read())
    assert saved["experiment_meta"]["experiment_name"] == "unit test!" and len(saved["runs"]) == 2
    run0 = saved["runs"][0]
    assert run0["status"] == "success" and run0["history"][0]["voxel_map"] == "voxels"
    met = run0["history"][0]["metrics"]
    assert met["search_validity_metric"][0]["valid"] is True and met["voxel_count_metric"]["feature_dim"] == d
    assert "error" in met["navigability_metric"]  # BaseEvaluator.evaluate raises; the manager records it
    assert saved["runs"][1]["history"][0]["metrics"]["search_validity_metric"][0]["query"] == "chair"
    missing = mgr.run_experiment({"dataset_path": "x", "annotation_path": "y", "voxel_dir": str(tmp_path / "nope"),
                                  "params": {}, "eval_funcs": []})
    assert missing["status"] == "failed"


def test_query_cli(vsm_mod, tmp_path, capsys):
    from vsm import query_voxelmap as cli

    m, vdir = _saved_map(vsm_mod, tmp_path)
    d = m.get_features().shape[1]
    q = np.random.default_rng(3).normal(size=d).astype(np.float32)
    np.save(tmp_path / "q.npy", q)
    res = cli.main(["--voxel_map_dir", vdir, "--query_prompt", "a red chair", "--top_k", "2", "--embedding",
                    str(tmp_path / "q.npy"), "--output_dir", str(tmp_path / "out")])
    widx, wsims, _ = vo.query(m.get_features(), q / np.linalg.norm(q), 2)
    assert [r["voxel_index"] for r in res] == widx
    np.testing.assert_allclose([r["similarity"] for r in res], wsims, rtol=RTOL, atol=1e-6)
    assert os.path.exists(tmp_path / "out" / "a_red_chair" / "retrieval_results.json")
    assert "Top 1 retrieval result" in capsys.readouterr().out
