"""Where a streaming multi-GPU build spends its time (run under torchrun, 2+ GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/stream_diag.py [submaps per rank]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import vsm
from vsm import _native as N
from vsm import dist as vdist
from vsm import synth_device

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "traj"  # "traj": corridor at 2 cm; "room": BASELINE config 2 (one room, 5 cm)
vs = 0.02 if mode == "traj" else 0.05
S, H, W, d = 32, 294, 518, 512
emb = torch.randn((S, H, W, d), device=dev, dtype=torch.float32).to(torch.bfloat16)
gm = vsm.GraphMap()
for j in range(per_rank):
    i = rank * per_rank + j
    if mode == "traj":
        dd = synth_device.make_trajectory_submap_device(4321, i, S=S, H=H, W=W, d=d, device=dev, with_emb=False, emb_from=emb)
    else:
        dd = synth_device.make_submap_device(1234 + rank, j, S=S, H=H, W=W, d=d, device=dev, first_frame_number=32 * j, with_emb=False)
        dd.emb = emb
    gm.add_submap(synth_device.to_submap(dd))
torch.cuda.synchronize()
out = {}
for rep in range(3):
    N.lib.vsm_map_cache_release()
    dist.barrier()
    torch.cuda.synchronize()
    timings = {"per_round": True}
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if K > 0:
        m, st = vdist.build_sharded_streaming(gm, vs, K, round_capacity=int(1.2 * 330000 * K) + 65536,
                                              owner_capacity=int(1.3 * 330000 * per_rank) + (1 << 18), profile=True, timings=timings)
    else:
        m, st = vdist.build_sharded(gm, vs, capacity_hint=1 << 20, profile=True)
        timings = {}
    e1.record()
    torch.cuda.synchronize()
    timings["device_ms"] = e0.elapsed_time(e1)
    out = {"rank": rank, "wall_ms": 1e3 * (time.perf_counter() - t0), "voxels_owner": m._dm.num_voxels, **timings,
           "fuse_profile": gm.last_profile}
    del m
for r in range(world):
    if r == rank:
        print(json.dumps(out), flush=True)
    dist.barrier()
dist.destroy_process_group()
