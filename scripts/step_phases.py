"""Which phase of GraphMap.build_semantic_voxel_map stalls while another process polls NVML? (GPU box only)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
import torch
import vsm
from vsm import synth_device, _native as N
from vsm.map import wrap_device_map
import bench

n = 20
gm = vsm.GraphMap()
for i in range(n):
    gm.add_submap(synth_device.to_submap(synth_device.make_submap_device(1234, i), host=False))
torch.cuda.synchronize()
hint = 1 << 18
child = None
keep = None
for rep in range(40):
    if rep == 10 and os.environ.get('POLL') == '1':
        child = subprocess.Popen([sys.executable, "-c", bench._CLOCK_CHILD, "", "0.01"], stdout=subprocess.DEVNULL)
        time.sleep(0.5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dm, fused, names = gm.fuse_into_device_map(0.05, capacity_hint=hint, profile=True)
    t1 = time.perf_counter()
    dm.finalize(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    m = wrap_device_map(dm, fused, names, 0.05)
    t3 = time.perf_counter()
    hint = max(hint, int(dm.num_voxels * 1.05) + 1024)
    del fused
    if os.environ.get("ALTERNATE") == "1":
        prev, keep = keep, (m, dm)       # the previous map dies here, like in bench.py's loop
        del prev, m
    else:
        del m
        dm.close()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    f = lambda a, b: f"{1e3*(b-a):7.2f}"
    if rep >= 8:
        print(f"rep {rep:2d}{' poll' if child else '     '}: total {f(t0,t4)} | fuse {f(t0,t1)} finalize {f(t1,t2)} wrap {f(t2,t3)} close {f(t3,t4)}")
if child: child.terminate()
