"""Side-car persistence at scale: save / load time of a V-voxel map (d=512 fp32 features).
usage: python scripts/persist_bench.py [V=4000000] [dir=/dev/shm/vsm_persist]   (prints one JSON line)"""
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vggt-slam_b200"))
import numpy as np
import torch

import vsm
from vsm import _native as N
from vsm import voxel_map as vm
from vsm.semantic_voxel import LazyContributors, SemanticVoxel, SemanticVoxelMap

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
root = sys.argv[2] if len(sys.argv) > 2 else "/dev/shm/vsm_persist"
d, vs = 512, 0.02
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
# distinct voxel centres on a lattice, random unit features, generated and loaded block by block
side = int(round(V ** (1 / 3))) + 1
dm = vm.DeviceVoxelMap(vs, d, N.F32, capacity=V)
stream = torch.cuda.current_stream(dev).cuda_stream
import ctypes as C
N.check(N.lib.vsm_map_load_begin(dm._h, V, C.c_void_p(stream)))
B = 1 << 18
for r0 in range(0, V, B):
    n = min(B, V - r0)
    i = torch.arange(r0, r0 + n, device=dev, dtype=torch.int64)
    c = torch.stack([(i % side), (i // side) % side, i // (side * side)], dim=1).to(torch.float32) * vs + 0.5 * vs
    f = torch.randn((n, d), device=dev, generator=g)
    N.check(N.lib.vsm_map_load_rows(dm._h, r0, n, C.c_void_p(c.data_ptr()), C.c_void_p(f.data_ptr()), C.c_void_p(stream)))
    torch.cuda.synchronize()
dm.finalize()
centers = dm.export_geometry(coords=False, centers=True, counts=False, recon=False)[1].cpu().numpy()
vox = SemanticVoxel.lazy(vs, lambda: centers, dm.features_to_host, LazyContributors(V, lambda i: [(int(i) % 200, f"{int(i) % 32:06d}")]))
m = SemanticVoxelMap(vox, frame_name_maps={}, _device_map=dm)
# contributor CSR as a device-built map exports it: one (submap, mask) entry per voxel
m._contrib_csr = None
shutil.rmtree(root, ignore_errors=True)
t0 = time.perf_counter()


class _Csr:
    def export_contributors(self):
        off = np.arange(V + 1, dtype=np.int64)
        sub = (np.arange(V, dtype=np.int64) % 200).astype(np.int32)
        mask = np.zeros((V, 2), dtype=np.uint64)
        mask[:, 0] = np.uint64(1) << (np.arange(V, dtype=np.uint64) % np.uint64(32))
        return off, sub, mask


m._contrib_csr = {"dm": _Csr(), "frame_ids": {s: [f"{k:06d}" for k in range(32)] for s in range(200)}}
m.save_to_directory(root, sidecar=True, npz=False)
t_save = time.perf_counter() - t0
nbytes = sum(os.path.getsize(os.path.join(root, "sidecar", f)) for f in os.listdir(os.path.join(root, "sidecar")))
q = torch.randn(d, generator=torch.Generator().manual_seed(2)).numpy()
ref = m.query_with_embedding(q, top_k=5)[0]
del m, vox, dm
N.lib.vsm_map_cache_release()
torch.cuda.empty_cache()
t0 = time.perf_counter()
lm = SemanticVoxelMap.load_from_directory(root)
torch.cuda.synchronize()
t_load = time.perf_counter() - t0
ok = lm.query_with_embedding(q, top_k=5)[0] == ref and lm.voxels.contributors[V - 1] == [((V - 1) % 200, f"{(V - 1) % 32:06d}")]
print(json.dumps({"voxels": V, "dim": d, "dir": root, "bytes": nbytes, "save_s": t_save, "load_s": t_load,
                  "save_GBps": nbytes / t_save * 1e-9, "load_GBps": nbytes / t_load * 1e-9, "query_equal_after_load": bool(ok)}))
shutil.rmtree(root, ignore_errors=True)
