"""Indexed embeddings: a (S,H,W) image of table rows + an (M,d) table must give the same map as the
dense (S,H,W,d) array table[ids] the reference fuses (what semantic_embedder.get_fully_embedded_image paints:
semantic_embedder.py:324-349) -- against the dense GPU path and against the oracle on the expanded array."""
import dataclasses

import numpy as np
import pytest

import golden_io as gio
from oracle import voxel_oracle as vo
from vsm import synth

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-5


def _sam_like(seed, S, H, W, d, M, block=7):
    """Piecewise-constant mask ids (blocks of `block` pixels, id 0 = no mask) and a table whose row 0 is zero,
    with bf16-exact values so that the bf16 and float32 paths see the same numbers."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, M, size=(S, (H + block - 1) // block, (W + block - 1) // block))
    ids = np.repeat(np.repeat(coarse, block, axis=1), block, axis=2)[:, :H, :W].astype(np.int32)
    table = rng.normal(size=(M, d)).astype(np.float32)
    table = (table.view(np.uint32) & 0xFFFF0000).view(np.float32)
    table[0] = 0.0
    return np.ascontiguousarray(ids), table


def _subs(n, S=4, H=56, W=84, d=64, M=37, seed=83, **kw):
    out = []
    for i in range(n):
        s = synth.make_submap(seed, i, S=S, H=H, W=W, d=d, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                              first_frame_number=S * i, **kw)
        ids, table = _sam_like(seed * 31 + i, S, H, W, d, M)
        out.append((dataclasses.replace(s, emb=table[ids]), ids, table))
    return out


def _submap(vsm, s, ids=None, table=None, bf16=False, device=True):
    import torch

    sm = vsm.Submap(s.submap_id)
    pts, conf = s.points, s.conf
    if device:
        pts, conf = torch.from_numpy(pts).cuda(), torch.from_numpy(conf).cuda()
    sm.add_all_points(pts, s.colors, conf, s.conf_percentile, None)
    if ids is None:
        emb = torch.from_numpy(s.emb).cuda().to(torch.bfloat16) if bf16 else s.emb
        sm.add_all_semantic_embeddings(emb)
    else:
        tab = torch.from_numpy(table).cuda().to(torch.bfloat16) if bf16 else table
        sm.add_all_semantic_embeddings_indexed(ids, tab)
    sm.set_conf_masks(s.conf)
    sm.set_reference_homography(s.H_world_map)
    sm.set_frame_ids(s.frame_paths)
    sm.set_last_non_loop_frame_index(s.last_non_loop_frame_index)
    return sm


@pytest.mark.parametrize("bf16", [False, True], ids=["f32", "bf16"])
@pytest.mark.parametrize("stride", [1, 2])
def test_indexed_build_equals_dense_and_oracle(bf16, stride):
    import vsm

    data = _subs(3, n_loop_frames=0)
    with np.errstate(all="ignore"):
        want = vo.build_global([gio.to_oracle_submap(s) for s, _, _ in data], 0.05, stride=stride, exact_order=False)
    maps = []
    for indexed in (False, True):
        gm = vsm.GraphMap()
        for s, ids, table in data:
            gm.add_submap(_submap(vsm, s, ids if indexed else None, table, bf16))
        maps.append((gm.build_semantic_voxel_map(0.05, stride=stride), gm.last_build_stats))
    (dense, st_d), (idx, st_i) = maps
    np.testing.assert_array_equal(idx.get_centers_world(), want.centers_world)
    np.testing.assert_array_equal(idx.get_centers_world(), dense.get_centers_world())
    ci, _, ni, _ = idx._dm.export_geometry()
    np.testing.assert_array_equal(ci.cpu().numpy(), want.coords)
    np.testing.assert_array_equal(ni.cpu().numpy(), want.counts)
    np.testing.assert_allclose(idx.get_features(), want.features, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(idx.get_features(), dense.get_features(), rtol=1e-5, atol=1e-6)
    assert idx.get_contributors() == want.contributors == dense.get_contributors()
    assert [s["n_fused"] for s in st_i] == [s["n_fused"] for s in st_d]


def test_indexed_submap_path_and_host_arrays():
    """Submap.get_semantic_voxel_in_world_frame (no filters, per-point contributors) with numpy ids / table."""
    import vsm

    (s, ids, table), = _subs(1, S=3)
    with np.errstate(all="ignore"):
        want = vo.fuse_submap(gio.to_oracle_submap(s), 0.05)
    got = _submap(vsm, s, ids.astype(np.int64), table, device=False).get_semantic_voxel_in_world_frame(0.05)
    np.testing.assert_array_equal(got.centers_world, want.centers_world)
    np.testing.assert_allclose(got.features, want.features, rtol=RTOL, atol=ATOL)
    assert list(got.contributors) == want.contributors


def test_indexed_bad_index_fails_before_the_map_is_touched():
    import torch
    import vsm
    from vsm import _native as N
    from vsm import voxel_map as vm

    (s, ids, table), = _subs(1, S=3)
    bad = ids.copy()
    conf_max = np.unravel_index(np.argmax(s.conf), s.conf.shape)      # a pixel that certainly passes the threshold
    bad[conf_max] = table.shape[0]
    dm = vm.DeviceVoxelMap(0.05, table.shape[1], N.F32)
    pts, conf, tab = (torch.from_numpy(x).cuda() for x in (s.points, s.conf, table))
    thr = vm.conf_threshold(s.conf, 25.0)
    S, H, W = s.conf.shape
    for flags in (0, N.FUSE_FILTERS):
        bad_t = torch.from_numpy(bad).cuda()
        p = dm.make_params(S, H, W, S, 1, thr, s.H_world_map, 0, flags, emb_index=bad_t, emb_rows=table.shape[0])
        with pytest.raises(ValueError, match="embedding index outside the table"):
            dm.fuse(pts, conf, tab, p)
        assert dm.num_voxels == 0
    # an out-of-range index under a pixel that is NOT confident is never read
    worst = np.unravel_index(np.argmin(s.conf), s.conf.shape)
    ok = ids.copy()
    ok[worst] = -5
    ok_t = torch.from_numpy(ok).cuda()
    st = dm.fuse(pts, conf, tab, dm.make_params(S, H, W, S, 1, thr, s.H_world_map, 0, 0, emb_index=ok_t,
                                                emb_rows=table.shape[0]))
    assert st["n_fused"] == int((s.conf >= thr).sum()) and dm.num_voxels > 0
    dm.close()


def test_indexed_nonfinite_table_row_takes_the_exact_path():
    """A NaN in the table: the optimistic pass notices, the build is redone with the row mask computed first
    (map.py:247 drops such rows before the percentiles) -- same result as the oracle on the expanded array."""
    import vsm

    data = _subs(2, n_loop_frames=0)
    for _, ids, table in data:
        table[5, 3] = np.nan
        table[9, 0] = np.inf
    data = [(dataclasses.replace(s, emb=table[ids]), ids, table) for s, ids, table in data]
    with np.errstate(all="ignore"):
        want = vo.build_global([gio.to_oracle_submap(s) for s, _, _ in data], 0.05, exact_order=False)
    gm = vsm.GraphMap()
    for s, ids, table in data:
        gm.add_submap(_submap(vsm, s, ids, table))
    m = gm.build_semantic_voxel_map(0.05)
    np.testing.assert_array_equal(m.get_centers_world(), want.centers_world)
    np.testing.assert_allclose(m.get_features(), want.features, rtol=RTOL, atol=ATOL)
    assert sum(st["n_fused"] for st in gm.last_build_stats) == want.n_points


def test_indexed_argument_errors():
    import vsm

    (s, ids, table), = _subs(1, S=2)
    sm = _submap(vsm, s, ids, table)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings_indexed([[1]], table)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings_indexed(ids.astype(np.float32), table)
    with pytest.raises(ValueError):
        sm.add_all_semantic_embeddings_indexed(ids[0], table)
    with pytest.raises(ValueError):
        sm.add_all_semantic_embeddings_indexed(ids[:, :-1], table)
    with pytest.raises(ValueError):
        sm.add_all_semantic_embeddings_indexed(ids, table[0])
    assert sm.dense_semantic_embeddings().shape == s.emb.shape
    np.testing.assert_array_equal(sm.dense_semantic_embeddings(), s.emb)
