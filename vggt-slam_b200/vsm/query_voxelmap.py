"""CLI: query a saved semantic voxel map with a text prompt (vggt_slam/query_voxelmap.py:15-71).

    python -m vsm.query_voxelmap --voxel_map_dir DIR --query_prompt "a red chair" [--top_k 5] [--embedding q.npy]

Loads the map (semantic_voxels.npz + frame_names.json), embeds the prompt with the upstream CLIP text tower (or
takes a precomputed, L2-normalised embedding from --embedding), scores all voxels on the GPU and prints the top-k
voxels with the latest frame that contributed to each.  The upstream script's PNG and viser output is UI and is not
reproduced; retrieval results are written as JSON instead."""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

from .semantic_voxel import SemanticVoxelMap
from .voxel_evaluators import clip_text_encoder


def query(voxel_map: SemanticVoxelMap, text_embedding: np.ndarray, top_k: int = 1):
    """The scoring call path of the upstream script: normalised (1,d) embedding -> top-k voxels -> latest frames."""
    q = np.asarray(text_embedding, dtype=np.float32).reshape(1, -1)
    idx, coords, sims = voxel_map.query_with_embedding(q, top_k=top_k)
    out = []
    for i, (vi, vc, s) in enumerate(zip(idx, coords, sims)):
        frame_name, submap_id, frame_id = voxel_map.get_latest_frame_at_voxel(vi)
        out.append({"rank": i + 1, "voxel_index": int(vi), "voxel_coord": [int(x) for x in vc], "similarity": float(s),
                    "frame_name": frame_name, "submap_id": int(submap_id), "frame_id": str(frame_id)})
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="Query semantic voxel map with a clip embedding vector.")
    ap.add_argument("--voxel_map_dir", type=str, required=True)
    ap.add_argument("--query_prompt", type=str, required=True)
    ap.add_argument("--output_dir", type=str, default="retrieval_results")
    ap.add_argument("--top_k", type=int, default=1)
    ap.add_argument("--embedding", type=str, default=None, help=".npy with a precomputed text embedding (skips CLIP)")
    args = ap.parse_args(argv)
    voxel_map = SemanticVoxelMap.load_from_directory(args.voxel_map_dir)
    if args.embedding:
        emb = np.load(args.embedding).astype(np.float32).reshape(-1)
    else:
        emb = clip_text_encoder()([args.query_prompt])[0]
    emb = emb / np.linalg.norm(emb)  # query_voxelmap.py:33
    results = query(voxel_map, emb, args.top_k)
    print("topk retrieval results for query prompt: ", args.query_prompt)
    for r in results:
        print(f"Top {r['rank']} retrieval result: Voxel index: {r['voxel_index']}, voxel coordinate: "
              f"{r['voxel_coord']}, similarity: {r['similarity']}")
        print(f"submap id: {r['submap_id']}, frame id: {r['frame_id']}")
        print(f"frame name: {r['frame_name']}")
    save_dir = os.path.join(args.output_dir, "_".join(args.query_prompt.split(" ")))
    os.makedirs(save_dir, exist_ok=True)
    with open(os.path.join(save_dir, "retrieval_results.json"), "w") as f:
        json.dump({"query_prompt": args.query_prompt, "results": results}, f, indent=2)
    return results


if __name__ == "__main__":
    main()
