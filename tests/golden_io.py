"""Helpers to read the committed golden fixtures (tests/golden/*.npz) back
into the input objects the oracle and the product API consume."""
from __future__ import annotations

import json
import os

import numpy as np

from vsm import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def inputs(z, prefix="in"):
    """list[synth.SynthSubmap] stored by make_golden.pack_inputs."""
    subs = []
    for i in range(int(z[f"{prefix}_n"])):
        pts = z[f"{prefix}{i}_points"]
        bits = z[f"{prefix}{i}_emb_bf16"]
        if bits.size:
            emb = (bits.astype(np.uint32) << 16).view(np.float32).reshape(pts.shape[:3] + (-1,))
        else:
            emb = z[f"{prefix}{i}_emb"]
        subs.append(synth.SynthSubmap(
            submap_id=int(z[f"{prefix}{i}_id"]), points=pts, conf=z[f"{prefix}{i}_conf"],
            colors=z[f"{prefix}{i}_colors"], emb=np.ascontiguousarray(emb), H_world_map=z[f"{prefix}{i}_H"],
            frame_paths=json.loads(str(z[f"{prefix}{i}_paths"])),
            last_non_loop_frame_index=int(z[f"{prefix}{i}_last"]), conf_percentile=float(z[f"{prefix}{i}_pct"])))
    return subs


def contributors(z, key):
    return [[(int(a), str(b)) for a, b in c] for c in json.loads(str(z[key]))]


def to_oracle_submap(s: synth.SynthSubmap):
    from oracle import voxel_oracle as vo

    fids = [vo.frame_id_from_name(p) for p in s.frame_paths]
    names = {str(f): os.path.basename(p) for f, p in zip(fids, s.frame_paths)}
    return vo.OracleSubmap(s.submap_id, s.points, s.conf, vo.conf_threshold(s.conf, s.conf_percentile), s.emb,
                           s.H_world_map, fids, names, s.last_non_loop_frame_index)
