"""Wall-clock of repeated GraphMap.build_semantic_voxel_map calls (GPU box only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
import torch
import vsm
from vsm import synth_device, _native as N

N.set_option("prep_variant", int(os.environ.get("PREP_VARIANT", "5")))
N.set_option("select_mode", int(os.environ.get("SELECT_MODE", "0")))
VS = float(os.environ.get("VOXEL", "0.05"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
gm = vsm.GraphMap()
for i in range(n):
    gm.add_submap(synth_device.to_submap(synth_device.make_submap_device(1234, i), host=False))
torch.cuda.synchronize()
hint = 1 << 18
keep = None
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = gm.build_semantic_voxel_map(VS, capacity_hint=hint, profile=(rep % 2 == 0))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    per = [round(s.get("n_map_voxels", 0) / 1000) for s in gm.last_build_stats]
    print(f"rep {rep}: build {1e3*(t1-t0):.1f} ms  hint {hint} voxels {m._dm.num_voxels} launches {N.launch_count()} prof {gm.last_profile}")
    hint = max(hint, int(m._dm.num_voxels * 1.05) + 1024)
    keep = m
