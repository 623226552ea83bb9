"""torchrun check of the multi-GPU build (one process per GPU, NCCL): the union of the owner shards equals the
single-GPU map (keys / counts bit-exact, features 1e-3), global indices equal the single map's sorted indices,
and the sharded query equals the single-GPU query.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import vsm
from vsm import dist as vdist
from vsm import synth
from test_gpu_parity import to_submap

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_sub = 2 * world + 1
subs = [synth.make_submap(71, i, S=4, H=56, W=84, d=64, mode="sl4", room=(2.4, 1.8, 1.2), start=0.2 * i,
                          first_frame_number=4 * i) for i in range(n_sub)]
gm = vsm.GraphMap()
for s in subs[rank::world]:
    gm.add_submap(to_submap(vsm, s, device_inputs=True))
transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
for _ in range(3):  # repeated exchanges alternate the inbox halves
    sh, stats = vdist.build_sharded(gm, 0.05, transport=transport)
keys = sh._dm.export_packed_keys().cpu().numpy()
coords, _, counts, _ = sh._dm.export_geometry()
feats = sh._dm.features_to_host()
contribs = sh.local.get_contributors().tolist() if sh.local is not None else []
gidx = sh.global_index.cpu().numpy()
rng = np.random.default_rng(5)
Q = rng.normal(size=(5, 64)).astype(np.float32)
Q /= np.linalg.norm(Q, axis=1, keepdims=True)
qi, qs = sh.query_with_embeddings(Q, top_k=7)
parts = [None] * world
dist.all_gather_object(parts, (keys, coords.cpu().numpy(), counts.cpu().numpy(), feats, contribs, gidx))
ok = True
if rank == 0:
    gm1 = vsm.GraphMap()
    for s in subs:
        gm1.add_submap(to_submap(vsm, s, device_inputs=True))
    single = gm1.build_semantic_voxel_map(0.05)
    s_coords, _, s_counts, _ = single._dm.export_geometry()
    V = single._dm.num_voxels
    assert sh.n_global == V, (sh.n_global, V)
    all_gidx = np.concatenate([p[5] for p in parts])
    assert sorted(all_gidx.tolist()) == list(range(V))
    order = np.argsort(all_gidx)
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts])[order], s_coords.cpu().numpy())
    np.testing.assert_array_equal(np.concatenate([p[2] for p in parts])[order], s_counts.cpu().numpy())
    np.testing.assert_allclose(np.concatenate([p[3] for p in parts])[order], single.get_features(), rtol=1e-3, atol=1e-5)
    allc = sum([p[4] for p in parts], [])
    assert [allc[i] for i in order] == single.get_contributors().tolist()
    si, _, ss = single.query_with_embeddings(Q, top_k=7)
    np.testing.assert_array_equal(qi, si)
    np.testing.assert_allclose(qs, ss, rtol=1e-3, atol=1e-6)
    print(f"dist_check ok: transport={transport} (peer broken: {bool(vdist._PEER_BROKEN)}) world={world} voxels={V} "
          f"shards={[len(p[0]) for p in parts]}")
dist.barrier()
if transport in ("peer", "auto"):
    from vsm import peer

    peer.close_all()
dist.destroy_process_group()
