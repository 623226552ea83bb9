"""A/B of the preparation-kernel variants and the two accumulate kernels on the benchmark's submap shape
(libvsm's own CUDA events: whole fuse call and accumulate kernel, per submap).
    python scripts/prep_ab.py [submaps] > gpurun_out/prep_ab.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200")):
    sys.path.insert(0, p)
import torch

import vsm
from vsm import _native as N
from vsm import synth_device

n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
gm = vsm.GraphMap()
for i in range(n_sub):
    d = synth_device.make_submap_device(1234, i, first_frame_number=32 * i)
    gm.add_submap(synth_device.to_submap(d))
torch.cuda.synchronize()
print(f"{'voxel':>6} {'prep':>5} {'acc':>4} {'fuse ms/submap':>15} {'accumulate ms':>14} {'prep ms':>8} {'voxels':>9}")
for vs in (0.05, 0.02):
    for prep, acc in ((5, 0), (13, 0), (13, 2), (13, 3)):
        N.set_option("prep_variant", prep)
        N.set_option("select_mode", acc)  # (the 'acc' column: select mode)
        hint = 1 << 18
        for _ in range(3):
            m = gm.build_semantic_voxel_map(vs, capacity_hint=hint, profile=True)
            hint = max(hint, int(m._dm.num_voxels * 1.1))
        tot = {"fuse_ms": 0.0, "accumulate_ms": 0.0}
        reps = 5
        for _ in range(reps):
            m = gm.build_semantic_voxel_map(vs, capacity_hint=hint, profile=True)
            for k in tot:
                tot[k] += gm.last_profile[k]
        f, a = tot["fuse_ms"] / reps / n_sub, tot["accumulate_ms"] / reps / n_sub
        print(f"{vs:6.2f} {prep:5d} {acc:4d} {f:15.4f} {a:14.4f} {f - a:8.4f} {m._dm.num_voxels:9d}", flush=True)
        del m
N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
N.set_option("select_mode", 0)
