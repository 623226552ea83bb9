"""Where does host time go in a build?  Prints wall-clock per stage (GPU box only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
import torch
import vsm
from vsm import synth_device, voxel_map as vm, _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
subs = []
for i in range(n):
    d = synth_device.make_submap_device(1234, i)
    subs.append(synth_device.to_submap(d, host=False))
torch.cuda.synchronize()

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

for rep in range(3):
    t0 = sync()
    dm = vm.DeviceVoxelMap(0.05, 512, N.BF16, capacity=1 << 18)
    t1 = sync()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms")
    for sm in subs:
        S, H, W = sm.pointclouds.shape[:3]
        ta = sync()
        p = dm.make_params(S, H, W, S, 1, sm.conf_threshold, sm.H_world_map, sm.submap_id, N.FUSE_FILTERS)
        tb = time.perf_counter()
        st = dm.fuse(sm._device("points"), sm._device("conf"), sm.embeddings_on_device(), p)
        tc = sync()
        print(f"   submap {sm.submap_id}: params {1e3*(tb-ta):.2f} ms fuse {1e3*(tc-tb):.2f} ms  fused {st['n_fused']} vox {st['n_submap_voxels']} map {st['n_map_voxels']}")
    t2 = sync()
    dm.finalize()
    t3 = sync()
    from vsm.map import wrap_device_map; m = wrap_device_map(dm, [], {}, 0.05)
    t4 = sync()
    del m
    dm.close()
    t5 = sync()
    print(f"   finalize {1e3*(t3-t2):.2f} wrap {1e3*(t4-t3):.2f} destroy {1e3*(t5-t4):.2f}")
