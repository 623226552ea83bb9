"""Wall-clock phases of GraphMap.build_semantic_voxel_map (GPU box only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
import torch
import vsm
from vsm import synth_device, voxel_map as vm, _native as N
from vsm.map import wrap_device_map

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N.set_option('select_mode', int(os.environ.get('SELECT_MODE', '0')))
subs = [synth_device.to_submap(synth_device.make_submap_device(1234, i), host=False) for i in range(n)]
torch.cuda.synchronize()
hint = 1 << 18
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dm = vm.DeviceVoxelMap(0.05, 512, N.BF16, capacity=hint)
    dm.profile_enable(True)
    t1 = time.perf_counter()
    for sm in subs:
        S, H, W = sm.pointclouds.shape[:3]
        p = dm.make_params(S, H, W, S, 1, sm.conf_threshold, sm.H_world_map, sm.submap_id, N.FUSE_FILTERS)
        dm.fuse_async(sm._device("points"), sm._device("conf"), sm.embeddings_on_device(), p)
    t2 = time.perf_counter()
    st = dm.collect()
    t3 = time.perf_counter()
    dm.finalize()
    t4 = time.perf_counter()
    m = wrap_device_map(dm, [], {}, 0.05)
    t5 = time.perf_counter()
    hint = max(hint, int(dm.num_voxels * 1.05) + 1024)
    prof = dm.profile()
    del m
    dm.close()
    torch.cuda.synchronize(); t6 = time.perf_counter()
    f = lambda a, b: f"{1e3*(b-a):6.2f}"
    print(f"rep {rep}: total {f(t0,t6)} | create {f(t0,t1)} enqueue {f(t1,t2)} collect-wait {f(t2,t3)} finalize {f(t3,t4)} wrap {f(t4,t5)} destroy {f(t5,t6)} | device fuse_ms {prof['fuse_ms']:.2f} acc_ms {prof['accumulate_ms']:.2f}")
