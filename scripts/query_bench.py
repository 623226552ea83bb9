"""Text-query sweep (BASELINE config 4): V voxels x P prompts, top-k, engines 1 (exact fp32), 2 (tcgen05, TF32 on the fp32 sums)
and 3 (tcgen05, bf16 shadow); the tensor-core engines must return engine 1's indices.
Prints one JSON line per point: latency, effective HBM GB/s over V*d*4 bytes, TFLOP/s, candidates."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
import numpy as np
import torch
from vsm import _native as N
from vsm import voxel_map as vm

V = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
d = 512
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prompts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 8, 16, 64, 128, 256]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
feats = torch.empty((V, d), dtype=torch.float32, device=dev)
for r0 in range(0, V, 1 << 20):
    r1 = min(V, r0 + (1 << 20))
    x = torch.randn((r1 - r0, d), dtype=torch.float32, device=dev, generator=g)
    feats[r0:r1] = x / x.norm(dim=1, keepdim=True) * (0.3 + 0.7 * torch.rand((r1 - r0, 1), device=dev, generator=g))
centers = torch.rand((V, 3), dtype=torch.float32, device=dev, generator=g) * 1000.0
dm = vm.DeviceVoxelMap(0.05, d, N.F32, capacity=V)
dm.load_dense(centers, feats)
del feats, centers
torch.cuda.empty_cache()
rng = np.random.default_rng(0)
for P in prompts:
    q = rng.normal(size=(P, d)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    qt = torch.from_numpy(q).to(dev)
    res = {}
    for eng in (1, 2, 3):
        if eng == 1 and P > 64 and V > 20_000_000:
            continue
        for _ in range(2):
            idx, sc = dm.query(qt, top_k=k, engine=eng)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            idx, sc = dm.query(qt, top_k=k, engine=eng)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[eng] = (ms, idx.clone())
        st = dm.query_stats()
        print(json.dumps({"V": V, "P": P, "k": k, "engine": eng, "ms": round(ms, 3), "GBps_over_Vd4": round(V * d * 4 / ms * 1e-6, 1),
                          "TFLOPs": round(2.0 * V * d * P / ms * 1e-9, 1), "candidates": st["last_candidates"],
                          "fallbacks": st["fallbacks"]}), flush=True)
    for eng in (2, 3):
        if 1 in res and eng in res:
            assert torch.equal(res[1][1], res[eng][1]), "engines disagree"
