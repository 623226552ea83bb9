"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the upstream ``vggt_slam`` package from /root/reference with a stub
``open3d`` module (the hot-path functions never touch open3d; only the module
top-level import does -- SURVEY.md 8c), feeds it seeded synthetic submaps from
``vsm.synth`` and stores INPUTS and OUTPUTS as small .npz fixtures so that the
GPU box, which has no /root/reference, can check parity against them.

No reference code is copied: this script only *calls* it.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "vggt-slam_b200"))
REF = os.environ.get("VSM_REFERENCE", "/root/reference")

from vsm import synth  # noqa: E402


def import_reference():
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    sys.path.insert(0, REF)
    import torch  # noqa: F401
    from vggt_slam.map import GraphMap
    from vggt_slam.semantic_voxel import SemanticVoxel, SemanticVoxelMap
    from vggt_slam.submap import Submap

    return Submap, GraphMap, SemanticVoxel, SemanticVoxelMap


def to_ref_submap(Submap, s: synth.SynthSubmap):
    sm = Submap(s.submap_id)
    sm.add_all_points(s.points, s.colors, s.conf, s.conf_percentile, None)
    sm.add_all_semantic_embeddings(s.emb)
    sm.set_conf_masks(s.conf)
    sm.set_reference_homography(s.H_world_map)
    sm.set_frame_ids(s.frame_paths)
    sm.set_last_non_loop_frame_index(s.last_non_loop_frame_index)
    return sm


def pack_inputs(subs, prefix="in"):
    out = {f"{prefix}_n": np.int64(len(subs))}
    for i, s in enumerate(subs):
        out[f"{prefix}{i}_id"] = np.int64(s.submap_id)
        out[f"{prefix}{i}_points"] = s.points
        out[f"{prefix}{i}_conf"] = s.conf
        out[f"{prefix}{i}_colors"] = s.colors
        out[f"{prefix}{i}_emb_bf16"] = synth.f32_to_bf16_bits(s.emb) if np.isfinite(s.emb).all() else np.zeros(0, np.uint16)
        out[f"{prefix}{i}_emb"] = s.emb if not np.isfinite(s.emb).all() else np.zeros(0, np.float32)
        out[f"{prefix}{i}_H"] = s.H_world_map
        out[f"{prefix}{i}_paths"] = np.array(json.dumps(s.frame_paths))
        out[f"{prefix}{i}_last"] = np.int64(s.last_non_loop_frame_index)
        out[f"{prefix}{i}_pct"] = np.float64(s.conf_percentile)
    return out


def contrib_json(contributors):
    return np.array(json.dumps([[list(t) for t in c] for c in contributors]))


def main():
    Submap, GraphMap, SemanticVoxel, SemanticVoxelMap = import_reference()
    import numpy
    import torch

    meta = {"numpy": numpy.__version__, "torch": torch.__version__, "reference": REF}
    print("generating goldens with", meta)

    # ---- case A: per-submap fusion, SL(4), one loop frame, N(0,1) embeddings
    sA = synth.make_submap(11, 3, S=4, H=28, W=42, d=16, mode="sl4", n_loop_frames=1, first_frame_number=8)
    rA = to_ref_submap(Submap, sA)
    out = pack_inputs([sA])
    out["meta"] = np.array(json.dumps(meta))
    out["conf_threshold"] = np.asarray(rA.get_conf_threshold())
    for tag, ign in (("all", False), ("noloop", True)):
        for vs in (0.1, 0.05):
            v = rA.get_semantic_voxel_in_world_frame(vs, ignore_loop_closure_frames=ign)
            k = f"{tag}_{vs}"
            out[f"{k}_centers"] = v.centers_world
            out[f"{k}_features"] = v.features
            out[f"{k}_contrib"] = contrib_json(v.contributors)
    # world-frame point extraction (a5)
    out["world_points_s1"] = rA.get_points_in_world_frame(1)
    out["world_points_s2"] = rA.get_points_in_world_frame(2)
    out["colors_s1"] = rA.get_points_colors(1)
    out["colors_s3"] = rA.get_points_colors(3)
    pl, fl, ml = rA.get_points_list_in_world_frame(ignore_loop_closure_frames=True)
    out["plist_points"] = np.stack(pl)
    out["plist_ids"] = np.asarray(fl, dtype=np.float64)
    out["plist_masks"] = np.stack(ml)
    np.savez_compressed(os.path.join(HERE, "case_a_submap_sl4.npz"), **out)

    # ---- case B: per-submap fusion, Sim(3) mode, painted embeddings, NaN/Inf injected, no filters
    sB = synth.make_submap(12, 0, S=3, H=28, W=42, d=8, mode="sim3", emb_kind="painted", bad_fraction=0.003,
                           room=(3.0, 2.0, 1.5))
    rB = to_ref_submap(Submap, sB)
    out = pack_inputs([sB])
    out["meta"] = np.array(json.dumps(meta))
    out["conf_threshold"] = np.asarray(rB.get_conf_threshold())
    with np.errstate(all="ignore"):
        v = rB.get_semantic_voxel_in_world_frame(0.05)
    out["centers"] = v.centers_world
    out["features"] = v.features
    out["contrib"] = contrib_json(v.contributors)
    np.savez_compressed(os.path.join(HERE, "case_b_submap_sim3_bad.npz"), **out)

    # ---- case C: global build, 3 overlapping submaps, SL(4), filters active, NaN/Inf injected
    room = (1.6, 1.2, 0.9)
    subsC = [
        synth.make_submap(13, i, S=4, H=28, W=42, d=16, mode="sl4", room=room, n_loop_frames=(1 if i == 1 else 0),
                          bad_fraction=(0.004 if i != 2 else 0.0), start=0.25 * i, first_frame_number=8 + 3 * i,
                          emb_kind=("painted" if i == 2 else "normal"))
        for i in range(3)
    ]
    gm = GraphMap()
    for s in subsC:
        gm.add_submap(to_ref_submap(Submap, s))
    out = pack_inputs(subsC)
    out["meta"] = np.array(json.dumps(meta))
    for tag, kw in (
        ("s1_dedup", dict(stride=1)),
        ("s2_dedup", dict(stride=2)),
        ("s1_nodedup", dict(stride=1, deduplicate_contributors=False)),
        ("s1_withloop", dict(stride=1, ignore_loop_closure_frames=False)),
    ):
        try:
            with np.errstate(all="ignore"):
                m = gm.build_semantic_voxel_map(0.05, use_torch=False, **kw)
        except IndexError as e:  # loop frames have no frame id upstream (map.py:240)
            out[f"{tag}_error"] = np.array(type(e).__name__)
            continue
        v = m.get_voxels()
        out[f"{tag}_centers"] = v.centers_world
        out[f"{tag}_features"] = v.features
        out[f"{tag}_contrib"] = contrib_json(v.contributors)
        out[f"{tag}_names"] = np.array(json.dumps(m.frame_name_maps))
        out[f"{tag}_recon_coords"] = m._voxel_coords
        if tag == "s1_dedup":
            mapC = m
    # ---- query / lookup on the s1_dedup map (a11-a14)
    rng = np.random.default_rng(99)
    feats = mapC.get_features()
    d = feats.shape[1]
    Q = rng.normal(size=(6, d)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    out["q"] = Q
    for k in (1, 5):
        idxs, coords, sims = [], [], []
        for p in range(Q.shape[0]):
            qi = Q[p] if p % 2 == 0 else Q[p][None, :]  # both accepted shapes
            i, c, s = mapC.query_with_embedding(qi, top_k=k)
            idxs.append(i)
            coords.append(c)
            sims.append(s)
        out[f"q_k{k}_idx"] = np.asarray(idxs, dtype=np.int64)
        out[f"q_k{k}_coords"] = np.stack(coords)
        out[f"q_k{k}_sims"] = np.asarray(sims, dtype=np.float64)
    lat = [mapC.get_latest_frame_at_voxel(int(i)) for i in range(0, feats.shape[0], 7)]
    out["latest_every7"] = np.array(json.dumps([[a, int(b), c] for a, b, c in lat]))
    centers = mapC.get_centers_world()
    probes = np.concatenate([centers[::5] + rng.uniform(-0.02, 0.02, size=centers[::5].shape).astype(np.float32),
                             rng.uniform(-3, 3, size=(40, 3)).astype(np.float32)])
    out["probe_pos"] = probes
    out["probe_idx"] = np.asarray([-1 if (i := mapC.get_index_at_position(p)) is None else i for p in probes],
                                  dtype=np.int64)
    # ---- persistence written by the reference (a15)
    with tempfile.TemporaryDirectory() as td:
        # get_latest_frame_at_voxel sorted some contributor lists in place above; rebuild a clean map
        with np.errstate(all="ignore"):
            m2 = gm.build_semantic_voxel_map(0.05, stride=2, use_torch=False)
        m2.save_to_directory(td)
        with open(os.path.join(td, "semantic_voxels.npz"), "rb") as f:
            out["saved_npz_bytes"] = np.frombuffer(f.read(), dtype=np.uint8)
        with open(os.path.join(td, "frame_names.json")) as f:
            out["saved_json"] = np.array(f.read())
    np.savez_compressed(os.path.join(HERE, "case_c_global_sl4.npz"), **out)

    # ---- case D: percentile arithmetic (a1 + bbox filter) on awkward sizes
    out = {"meta": np.array(json.dumps(meta))}
    rng = np.random.default_rng(5)
    cases = []
    for n in (1, 2, 3, 10, 101, 1000, 4097, 70001):
        x = (1.0 + rng.gamma(2.0, 2.0, size=n)).astype(np.float32)
        for q in (0.0, 0.5, 25.0, 50.0, 99.5, 100.0, 33.3):
            cases.append((n, q, np.percentile(x, q)))
        out[f"x_{n}"] = x
    out["cases"] = np.asarray(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "case_d_percentile.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
