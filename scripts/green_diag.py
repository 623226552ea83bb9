import os, sys
sys.path.insert(0, "/root/repo/vggt-slam_b200")
import torch, vsm
from vsm import _native as N, synth_device
gm = vsm.GraphMap()
for i in range(6):
    gm.add_submap(synth_device.to_submap(synth_device.make_submap_device(1234, i, first_frame_number=32 * i)))
torch.cuda.synchronize()
def run(tag, vs=0.05):
    hint = 1 << 19
    for _ in range(2):
        m = gm.build_semantic_voxel_map(vs, capacity_hint=hint)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m = gm.build_semantic_voxel_map(vs, capacity_hint=hint, profile=True)
    e1.record(); torch.cuda.synchronize()
    x = torch.empty(1 << 28, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); x.fill_(1.0); x.mul_(2.0); b.record(); torch.cuda.synchronize()
    print(f"{tag:28s} build {e0.elapsed_time(e1)/6:.3f} ms/submap  prof {gm.last_profile['fuse_ms']/6:.3f}  torch fill+mul 1 GiB {a.elapsed_time(b):.3f} ms", flush=True)
run("fresh")
N.set_option("green_prep_sms", 64); N.set_option("acc_ctas_per_sm", 3)
run("green 64")
N.set_option("green_prep_sms", 0); N.set_option("acc_ctas_per_sm", 2)
run("after green off")
run("after green off (2 cm)", 0.02)
N.set_option("green_prep_sms", 64); N.set_option("green_prep_sms", 72); N.set_option("green_prep_sms", 0)
run("after 2 more create/destroy")
