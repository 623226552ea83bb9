"""SM-partition A/B (CUDA green contexts): whole builds of `n_sub` benchmark-shaped submaps, timed with CUDA events on
the caller's stream, for several splits of the SMs between the preparation kernels and the accumulate kernel.
    python scripts/green_ab.py [submaps] [voxel sizes, comma separated] > gpurun_out/green_ab.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vggt-slam_b200")):
    sys.path.insert(0, p)
import torch

import vsm
from vsm import _native as N
from vsm import synth_device

n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sizes = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.05, 0.02]
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
N.set_option("prep_variant", variant)
N.set_option("select_mode", sel)
ring = int(sys.argv[5]) if len(sys.argv) > 5 else 1
N.set_option("acc_ring", ring)
tma = int(sys.argv[6]) if len(sys.argv) > 6 else 0
N.set_option("acc_tma", tma)
splits = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [0, 48, 56, 64, 72, 80, 88, 96]
mix = int(sys.argv[8]) if len(sys.argv) > 8 else 0
N.set_option("acc_mix", mix)
run = int(sys.argv[9]) if len(sys.argv) > 9 else 0
N.set_option("acc_tmarun", run)
print("prep_variant", variant, "select_mode", sel, "acc_ring", ring, "acc_tma", tma, "acc_mix", mix, "acc_tmarun", run)
gm = vsm.GraphMap()
for i in range(n_sub):
    d = synth_device.make_submap_device(1234, i, first_frame_number=32 * i)
    gm.add_submap(synth_device.to_submap(d))
torch.cuda.synchronize()


def run(vs, reps=4):
    hint = 1 << 18
    for _ in range(3):
        m = gm.build_semantic_voxel_map(vs, capacity_hint=hint)
        hint = max(hint, int(m._dm.num_voxels * 1.1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        m = gm.build_semantic_voxel_map(vs, capacity_hint=hint)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    counts = m._dm.export_geometry(coords=False, centers=False, counts=True, recon=False)[2]
    sig = (m._dm.num_voxels, int(counts.sum().item()), float(torch.from_numpy(m.get_features()).double().abs().sum()))
    return best / n_sub, sig


print(f"{'voxel':>6} {'prep SMs':>9} {'acc CTAs/SM':>12} {'ms/submap':>10}  signature (voxels, points, sum|f|)")
for vs in sizes:
    ref = None
    for prep_sms, ctas in [(p, 3 if p else 2) for p in splits]:
        if prep_sms < 0:  # overlap on two plain streams, no partition (round-1 experiment)
            N.set_option("green_prep_sms", 0)
            N.set_option("overlap", 1)
        else:
            N.set_option("overlap", 0)
            N.set_option("green_prep_sms", prep_sms)
        N.set_option("acc_ctas_per_sm", ctas)
        try:
            ms, sig = run(vs)
        except Exception as e:  # noqa: BLE001
            print(f"{vs:6.2f} {prep_sms:9d} {ctas:12d}  failed: {e!r}", flush=True)
            continue
        ref = ref or sig
        ok = sig[:2] == ref[:2] and abs(sig[2] - ref[2]) <= 1e-6 * ref[2]
        print(f"{vs:6.2f} {prep_sms:9d} {ctas:12d} {ms:10.4f}  {sig} {'ok' if ok else 'MISMATCH'}", flush=True)
N.set_option("green_prep_sms", 0)
N.set_option("overlap", 0)
N.set_option("acc_ctas_per_sm", 2)
N.set_option("prep_variant", N.DEFAULT_PREP_VARIANT)
N.set_option("select_mode", 0)
N.set_option("acc_ring", 0)
N.set_option("acc_tma", 0)
N.set_option("acc_mix", 0)
N.set_option("acc_tmarun", 0)
