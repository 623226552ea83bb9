"""torchrun check of the multi-GPU builds (one process per GPU, NCCL): vsm.dist.parity_check for the one-shot peer
exchange, the collective (pack -> all-to-all -> merge) route and the streaming rounds -- union of the owner shards ==
the single-GPU map (keys, counts, contributors bit-exact; features 1e-3; sharded query == single query) -- repeated so
that both inbox halves and the cached exchange are exercised.  Writes gpurun_out/dist_check_N<world>.json.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from vsm import dist as vdist
from vsm import synth_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reports = []
DIM = 64


def small(i):
    d = synth_device.make_submap_device(71, i, S=4, H=56, W=84, d=DIM, mode="sl4", room=(2.4, 1.8, 1.2),
                                        emb_dtype=torch.float32, first_frame_number=4 * i)
    return synth_device.to_submap(d)


def corridor(i):
    d = synth_device.make_trajectory_submap_device(72, i, S=4, H=56, W=84, d=DIM, room=(2.0, 1.5, 1.2),
                                                   emb_dtype=torch.bfloat16)
    return synth_device.to_submap(d)


n_sub = 4 * world + 1
for rep in range(2):
    reports.append(dict(vdist.parity_check(small, n_sub, 0.05, DIM), case="room, one-shot peer exchange"))
    reports.append(dict(vdist.parity_check(corridor, n_sub, 0.02, DIM, round_submaps=2), case="corridor 2 cm, streaming rounds of 2"))
    reports.append(dict(vdist.parity_check(small, n_sub, 0.05, DIM, round_submaps=1), case="room, streaming rounds of 1"))
ok = all(r["ok"] for r in reports)
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_check_N{world}.json"), "w") as f:
        json.dump({"world": world, "ok": ok, "reports": reports}, f, indent=1)
    for r in reports:
        print(("ok  " if r["ok"] else "FAIL"), r["case"], {k: r.get(k) for k in ("voxels", "keys", "counts", "features", "contributors", "query")},
              r["invariants"])
    print(f"dist_check {'ok' if ok else 'FAILED'}: world={world}")
dist.barrier()
from vsm import peer

peer.close_all()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
