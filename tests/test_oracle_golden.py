"""The numpy oracle against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

import golden_io as gio
from oracle import voxel_oracle as vo


def _same_float(a, b):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


@pytest.fixture(scope="module")
def case_a():
    return gio.load("case_a_submap_sl4.npz")


@pytest.fixture(scope="module")
def case_c():
    return gio.load("case_c_global_sl4.npz")


def test_conf_threshold_matches_reference(case_a):
    s = gio.inputs(case_a)[0]
    thr = vo.conf_threshold(s.conf, s.conf_percentile)
    assert thr.dtype == np.float32
    _same_float(thr, case_a["conf_threshold"])


@pytest.mark.parametrize("tag,ign", [("all", False), ("noloop", True)])
@pytest.mark.parametrize("vs", [0.1, 0.05])
def test_fuse_submap_bit_exact(case_a, tag, ign, vs):
    sm = gio.to_oracle_submap(gio.inputs(case_a)[0])
    got = vo.fuse_submap(sm, vs, ignore_loop_closure_frames=ign)
    k = f"{tag}_{vs}"
    _same_float(got.centers_world, case_a[f"{k}_centers"])
    assert got.features.dtype == np.float64
    _same_float(got.features, case_a[f"{k}_features"])  # same np.add.at order -> bit exact
    assert got.contributors == gio.contributors(case_a, f"{k}_contrib")
    # the fast (float64, sort based) summation stays inside the 1e-3 band
    fast = vo.fuse_submap(sm, vs, ignore_loop_closure_frames=ign, exact_order=False, with_contributors=False)
    np.testing.assert_array_equal(fast.coords, got.coords)
    np.testing.assert_allclose(fast.features, got.features, rtol=1e-3, atol=1e-5)


def test_world_points(case_a):
    s = gio.inputs(case_a)[0]
    thr = vo.conf_threshold(s.conf, s.conf_percentile)
    for stride in (1, 2):
        got = vo.points_in_world_frame(s.points, s.conf, thr, s.H_world_map, stride)
        assert got.dtype == np.float64
        _same_float(got, case_a[f"world_points_s{stride}"])
    _same_float(vo.filter_by_confidence(s.colors, s.conf, thr, 1).reshape(-1, 3), case_a["colors_s1"])
    _same_float(vo.filter_by_confidence(s.colors, s.conf, thr, 3).reshape(-1, 3), case_a["colors_s3"])
    fids = [vo.frame_id_from_name(p) for p in s.frame_paths]
    pl, fl, ml = vo.points_list_in_world_frame(s.points, s.conf, thr, s.H_world_map, fids,
                                               s.last_non_loop_frame_index, True)
    _same_float(np.stack(pl), case_a["plist_points"])
    _same_float(np.asarray(fl), case_a["plist_ids"])
    _same_float(np.stack(ml), case_a["plist_masks"])


def test_fma_chain_equals_blas(case_a):
    """The kernel's float64 FMA chain vs numpy's matmul on this host: equal
    after the float32 rounding (and here even before it)."""
    s = gio.inputs(case_a)[0]
    p = s.points.reshape(-1, 3)[:1500]
    a = vo.homography_apply_f64(p, s.H_world_map)
    b = vo.homography_apply_fma_chain(p, s.H_world_map)
    _same_float(a.astype(np.float32), b.astype(np.float32))
    assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) < 1e-15


def test_fuse_submap_nonfinite_points_keep_reference_quirk():
    """No filters on the per-submap path: NaN/Inf points land in the
    INT64_MIN voxel and poison it (submap.py:279-292)."""
    z = gio.load("case_b_submap_sim3_bad.npz")
    sm = gio.to_oracle_submap(gio.inputs(z)[0])
    with np.errstate(all="ignore"):
        got = vo.fuse_submap(sm, 0.05)
    _same_float(got.centers_world, z["centers"])
    np.testing.assert_array_equal(got.features, z["features"])  # NaN == NaN under assert_array_equal
    assert got.contributors == gio.contributors(z, "contrib")
    assert (got.coords[0] == np.iinfo(np.int64).min).all()


@pytest.mark.parametrize("tag,kw", [
    ("s1_dedup", dict(stride=1)),
    ("s2_dedup", dict(stride=2)),
    ("s1_nodedup", dict(stride=1, deduplicate_contributors=False)),
])
def test_build_global_bit_exact(case_c, tag, kw):
    subs = [gio.to_oracle_submap(s) for s in gio.inputs(case_c)]
    with np.errstate(all="ignore"):
        got = vo.build_global(subs, 0.05, **kw)
    _same_float(got.centers_world, case_c[f"{tag}_centers"])
    _same_float(got.features, case_c[f"{tag}_features"])
    assert got.contributors == gio.contributors(case_c, f"{tag}_contrib")
    assert got.frame_name_maps == json.loads(str(case_c[f"{tag}_names"]))
    _same_float(vo.coords_from_centers(got.centers_world, 0.05), case_c[f"{tag}_recon_coords"])


def test_build_global_with_loop_frames_raises_like_reference(case_c):
    assert str(case_c["s1_withloop_error"]) == "IndexError"
    subs = [gio.to_oracle_submap(s) for s in gio.inputs(case_c)]
    with pytest.raises(IndexError), np.errstate(all="ignore"):
        vo.build_global(subs, 0.05, ignore_loop_closure_frames=False)


def test_query_and_lookup(case_c):
    subs = [gio.to_oracle_submap(s) for s in gio.inputs(case_c)]
    with np.errstate(all="ignore"):
        m = vo.build_global(subs, 0.05)
    recon = vo.coords_from_centers(m.centers_world, 0.05)
    Q = case_c["q"]
    for k in (1, 5):
        for p in range(Q.shape[0]):
            idx, sims, _ = vo.query(m.features, Q[p], k)
            assert idx == case_c[f"q_k{k}_idx"][p].tolist()
            np.testing.assert_array_equal(recon[idx], case_c[f"q_k{k}_coords"][p])
            np.testing.assert_array_equal(np.asarray(sims, dtype=np.float64), case_c[f"q_k{k}_sims"][p])
    lat = json.loads(str(case_c["latest_every7"]))
    for j, i in enumerate(range(0, m.features.shape[0], 7)):
        sid, fid = vo.latest_contributor(m.contributors[i])
        assert [m.frame_name_maps[str(sid)][fid], sid, fid] == lat[j]
    table = vo.coord_index(recon)
    got = [table.get(vo.position_to_coord(p, 0.05), -1) for p in case_c["probe_pos"]]
    assert got == case_c["probe_idx"].tolist()
    # the lossy reconstruction really is lossy (Appendix A-4): some voxels are off by one
    assert (recon != m.coords).any()


def test_percentile_restatement_matches_numpy():
    z = gio.load("case_d_percentile.npz")
    for n, q, want in z["cases"]:
        x = z[f"x_{int(n)}"]
        got = vo.percentile_linear_restated(x, float(q))
        assert got.dtype == np.float32
        assert float(got) == float(want), (n, q, got, want)
        assert float(np.percentile(x, float(q))) == float(want)  # python float = weak scalar, like argparse gives
    rng = np.random.default_rng(0)
    for n in (5, 77, 12345, 300001):
        x = rng.normal(size=n).astype(np.float32) * 3
        for q in (0.5, 99.5, 25.0):
            assert float(vo.percentile_linear_restated(x, q)) == float(np.percentile(x, q))


def test_latest_uses_string_order():
    c = [(1, "9.0"), (1, "100.0"), (1, "10.0"), (0, "99.0")]
    assert vo.latest_contributor(c) == (1, "9.0")
