"""Evaluators over a semantic voxel map, with the reference's names, config keys and result fields
(vggt_slam/voxel_evaluators.py:12-166).  Differences that keep the results identical:
  * all prompts of a step are scored in ONE batched device query (query_with_embeddings) instead of one CPU
    matmul per prompt (voxel_evaluators.py:61-73);
  * the text encoder is injectable (``text_encoder(list[str]) -> (P,d) array``); the default loads the same
    HuggingFace CLIP model as upstream (openai/clip-vit-base-patch32) lazily and fails with a clear message
    when the weights are not available offline;
  * PerformanceEvaluator reports real query latencies instead of the upstream placeholder.
"""
from __future__ import annotations

import json
import os
import re
from typing import Callable, List, Optional, Sequence

import numpy as np


def get_ts(f):
    """Timestamp in a file name: the digits between '_' and '.' (voxel_evaluators.py:8-10)."""
    m = re.search(r"_(\d+)\.", str(f))
    return int(m.group(1)) if m else None


class BaseEvaluator:
    def ingest_chapter(self, chapter_metadata):
        pass

    def evaluate(self, voxel_map, step_info):
        raise NotImplementedError


def clip_text_encoder(model_name: str = "openai/clip-vit-base-patch32") -> Callable[[Sequence[str]], np.ndarray]:
    """The upstream text tower (voxel_evaluators.py:22-24, 63-65).  Needs the HF weights in the local cache."""
    import torch
    from transformers import CLIPModel, CLIPProcessor

    try:
        device = "cuda" if torch.cuda.is_available() else "cpu"
        model = CLIPModel.from_pretrained(model_name).to(device)
        proc = CLIPProcessor.from_pretrained(model_name)
    except Exception as e:  # offline box without cached weights
        raise RuntimeError(f"cannot load {model_name} ({e}); pass text_encoder=... to SearchValidityEvaluator") from e

    def encode(texts: Sequence[str]) -> np.ndarray:
        inputs = proc(text=list(texts), return_tensors="pt", padding=True).to(device)
        with torch.no_grad():
            return model.get_text_features(**inputs).float().cpu().numpy()

    return encode


class SearchValidityEvaluator(BaseEvaluator):
    """text -> embedding -> top-k voxels -> latest contributing frame -> is its timestamp within
    ``time_tolerance_ns`` of an annotation whose label contains the query?  (voxel_evaluators.py:20-119)"""

    def __init__(self, annotation_path, time_tolerance_ns=5e7, text_encoder: Optional[Callable] = None):
        self.time_tolerance = time_tolerance_ns
        self._encoder = text_encoder
        with open(annotation_path, "r") as f:
            data = json.load(f)
        if isinstance(data, dict) and "images" in data:
            annotations = data["images"]
        elif isinstance(data, list):
            annotations = data
        else:
            annotations = []

        def ann_ts(item):
            if "timestamp" in item:
                return item["timestamp"]
            name = item.get("file", item.get("file_name", ""))
            return get_ts(name) if name else None

        annotations.sort(key=lambda x: ann_ts(x) if ann_ts(x) is not None else 0)
        self.annotations = annotations
        self.timestamps = [ann_ts(x) for x in annotations]

    def _embed(self, queries: List[str]) -> np.ndarray:
        if self._encoder is None:
            self._encoder = clip_text_encoder()
        q = np.asarray(self._encoder(queries), dtype=np.float32)
        norms = np.linalg.norm(q, axis=1, keepdims=True)
        return np.where(norms > 0, q / np.where(norms > 0, norms, 1.0), q).astype(np.float32)  # voxel_evaluators.py:66-68

    def evaluate(self, voxel_map, step_info):
        queries = step_info.get("queries") or step_info.get("query")
        if not queries:
            return None
        if isinstance(queries, str):
            queries = [queries]
        top_k = int(step_info.get("top_k", 1))
        centers = voxel_map.get_centers_world()
        if centers is None or len(centers) == 0:
            return [{"query": q, "found": False, "reason": "Voxel map empty"} for q in queries]
        idx, coords, sims = voxel_map.query_with_embeddings(self._embed(list(queries)), top_k=top_k)
        results = []
        for p, user_query in enumerate(queries):
            if idx.shape[1] == 0:
                results.append({"query": user_query, "found": False, "reason": "No voxels returned"})
                continue
            voxel_index, voxel_coord, score = int(idx[p, 0]), coords[p, 0], float(sims[p, 0])
            frame_name, submap_id, frame_id = voxel_map.get_latest_frame_at_voxel(voxel_index)
            retrieved_ts = get_ts(frame_name) if frame_name else None
            is_valid, closest_gt, min_dt = False, None, float("inf")
            needle = user_query.lower()
            for i, ann in enumerate(self.annotations):
                if needle not in str(ann.get("label", "")).lower():
                    continue
                gt_ts = self.timestamps[i]
                if gt_ts is None or retrieved_ts is None:
                    continue
                dt = abs(gt_ts - retrieved_ts)
                if dt < min_dt:
                    min_dt, closest_gt = dt, ann
                if dt <= self.time_tolerance:
                    is_valid = True
            results.append({
                "query": user_query,
                "found": True,
                "valid": is_valid,
                "score": score,
                "retrieved_ts": retrieved_ts,
                "retrieved_voxel_index": voxel_index,
                "retrieved_voxel_coord": [int(x) for x in voxel_coord],
                "closest_gt_label": closest_gt["label"] if closest_gt and "label" in closest_gt else "None",
                "time_diff_ns": float(min_dt) if min_dt != float("inf") else None,
                "retrieved_img": os.path.basename(frame_name) if frame_name else None,
                "retrieved_submap_id": int(submap_id) if submap_id is not None else None,
                "retrieved_frame_id": str(frame_id) if frame_id is not None else None,
            })
        return results


class VoxelCountEvaluator(BaseEvaluator):
    def evaluate(self, voxel_map, step_info):
        centers = voxel_map.get_centers_world()
        n = int(centers.shape[0]) if centers is not None else 0
        dm = getattr(voxel_map, "_dm", None)
        if dm is not None and n > 0:
            feature_dim = int(dm.dim)  # no need to pull the features off the device for their width
        else:
            feats = voxel_map.get_features()
            feature_dim = int(feats.shape[1]) if feats is not None and feats.ndim == 2 else 0
        return {"num_voxels": n, "feature_dim": feature_dim, "voxel_size": float(voxel_map.get_voxel_size())}


class PerformanceEvaluator(BaseEvaluator):
    """Query latency of the map under evaluation (upstream returns a 'not_available' placeholder,
    voxel_evaluators.py:136-150)."""

    def __init__(self, prompts=(1, 8), top_k=10, repeats=10):
        self.prompts, self.top_k, self.repeats = tuple(prompts), int(top_k), int(repeats)

    def evaluate(self, voxel_map, step_info):
        dm = getattr(voxel_map, "_dm", None)
        if dm is None or dm.num_voxels == 0:
            return {"status": "not_available", "reason": "Voxel map empty"}
        import torch

        V, d = dm.num_voxels, dm.dim
        k = min(self.top_k, V)
        out = {"status": "ok", "num_voxels": V, "feature_dim": d, "top_k": k, "queries": {}}
        rng = np.random.default_rng(0)
        for P in self.prompts:
            q = rng.normal(size=(P, d)).astype(np.float32)
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            qt = torch.from_numpy(q).to(dm.device)
            for _ in range(2):
                dm.query(qt, top_k=k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(self.repeats):
                dm.query(qt, top_k=k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / self.repeats
            out["queries"][str(P)] = {"ms": ms, "GBps": V * d * 4 / (ms * 1e-3) * 1e-9}
        return out


def get_evaluator(name, config):
    """Same names as upstream (voxel_evaluators.py:152-166); ``config['text_encoder']`` is an extra, optional key."""
    if name == "search_validity_metric":
        return SearchValidityEvaluator(annotation_path=config["annotation_path"], text_encoder=config.get("text_encoder"))
    if name in {"node_count_metric", "voxel_count_metric"}:
        return VoxelCountEvaluator()
    if name in {"navigability_metric", "localization_metric"}:
        print(f"Warning: Evaluator '{name}' is not supported for voxel maps.")
        return BaseEvaluator()
    if name == "performance_metric":
        return PerformanceEvaluator()
    print(f"Warning: Unknown evaluator '{name}'")
    return BaseEvaluator()
