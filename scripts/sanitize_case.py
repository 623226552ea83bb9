"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck), checked against the goldens / the exact
engine while they run:
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py [fuse|query|peer|all]
  fuse   golden case C through GraphMap.build_semantic_voxel_map (one-table preparation + deferred box, and the two-table /
         radix variants), centres and features against the reference's outputs
  query  70 000 voxels x 512, 256 and 64 prompts: engines 2 (TF32) and 3 (bf16 shadow) == engine 1
  peer   the one-sided exchange with two 'ranks' on this GPU (real inboxes), union of the owners == single map"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import golden_io as gio
import vsm
from vsm import _native as N

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("fuse", "all"):
    from test_gpu_parity import graph_from

    z = gio.load("case_c_global_sl4.npz")
    for variant, sel in ((13, 0), (13, 1), (5, 1), (5, 2)):
        N.set_option("prep_variant", variant)
        N.set_option("select_mode", sel)
        m = graph_from(vsm, gio.inputs(z)).build_semantic_voxel_map(0.05)
        np.testing.assert_array_equal(m.get_centers_world(), z["s1_dedup_centers"])
        np.testing.assert_allclose(m.get_features(), z["s1_dedup_features"], rtol=1e-3, atol=1e-5)
        print("fuse ok: prep_variant", variant, "select_mode", sel, m.get_centers_world().shape[0], "voxels", flush=True)
    N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
    N.set_option("select_mode", 0)
if what in ("query", "all"):
    from test_gpu_query_tc import _random_map

    dm, feats = _random_map(70000, 512, 11)
    rng = np.random.default_rng(5)
    for P in (256, 64):
        q = rng.normal(size=(P, 512)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        i1, _ = dm.query(q, top_k=10, engine=1)
        for eng in (2, 3):
            ie, _ = dm.query(q, top_k=10, engine=eng)
            assert torch.equal(i1, ie), (P, eng)
        print("query ok: P", P, dm.query_stats(), flush=True)
    dm.close()
if what in ("peer", "all"):
    import test_gpu_dist as td

    td.test_peer_push_drain_equals_single_map(2)
    print("peer ok", flush=True)
