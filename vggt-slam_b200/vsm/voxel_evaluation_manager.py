"""Evaluation manager: expands a JSON config into jobs (datasets x generators x detectors x hyperparameters),
loads the saved voxel map of each job, runs the requested evaluators and writes results/evals/<name>.json --
the reference's schema and output layout (vggt_slam/voxel_evaluation_manager.py:13-130).

Worker processes are spawned (never forked): every worker owns its CUDA context and its own device map."""
from __future__ import annotations

import argparse
import json
import multiprocessing
import os
import time
from itertools import product

from . import voxel_evaluators as evaluators
from .semantic_voxel import SemanticVoxelMap


def run_experiment(config_packet):
    dataset_path = config_packet["dataset_path"]
    annotation_path = config_packet["annotation_path"]
    voxel_dir = config_packet.get("voxel_dir")
    params = config_packet["params"]
    if not voxel_dir:
        return {"status": "failed", "error": "Missing voxel_dir in config packet."}
    if not os.path.exists(voxel_dir):
        return {"status": "failed", "error": f"voxel_dir not found: {voxel_dir}"}
    voxel_map = SemanticVoxelMap.load_from_directory(voxel_dir)

    active = []
    for eval_name in config_packet["eval_funcs"]:
        try:
            ev = evaluators.get_evaluator(eval_name, {**config_packet})
            if hasattr(ev, "ingest_chapter"):
                ev.ingest_chapter([])
            active.append((eval_name, ev))
        except Exception as e:
            print(f"Warning: Failed to initialize evaluator {eval_name}: {e}")

    queries = params.get("queries", None)
    if isinstance(queries, str):
        queries = [queries]
    step_info = {
        "timestamp": time.time(),
        "experiment_name": config_packet.get("experiment_name", "default_eval"),
        "dataset_path": dataset_path,
        "annotation_path": annotation_path,
        "voxel_dir": voxel_dir,
        "top_k": params.get("top_k", 1),
    }
    if queries:
        step_info["queries"] = queries
        if len(queries) == 1:
            step_info["query"] = queries[0]

    metrics = {}
    for name, ev in active:
        try:
            metrics[name] = ev.evaluate(voxel_map, step_info)
        except Exception as e:  # upstream records the error and carries on (voxel_evaluation_manager.py:59-63)
            metrics[name] = {"error": str(e)}
    packet = {k: v for k, v in config_packet.items() if k != "text_encoder"}  # callables do not go into JSON
    return {"status": "success", "config": packet,
            "history": [{"voxel_map": os.path.basename(voxel_dir), "metrics": metrics}]}


def expand_jobs(config):
    jobs = []
    combos = product(config["datasets"], config.get("generators", [None]), config.get("detectors", [None]),
                     config.get("hyperparameters", [{}]))
    for ds, gen, det, params in combos:
        jobs.append({
            "dataset_path": ds["path"],
            "annotation_path": ds.get("annotation_file", os.path.join(ds["path"], "annotations.json")),
            "pcd_path": ds.get("pcd_file", "").strip(),
            "colmap_path": ds.get("colmap_file", "").strip(),
            "voxel_dir": ds.get("voxel_dir", ds.get("voxel_map_dir", "")).strip(),
            "generator": gen,
            "detector": det,
            "params": params,
            "eval_funcs": config["eval_functions"],
            "experiment_name": config.get("experiment_name", "default_eval"),
        })
    return jobs


def run(config, workers: int = 1, out_dir: str = "results/evals", text_encoder=None):
    jobs = expand_jobs(config)
    if text_encoder is not None:
        if workers != 1:
            raise ValueError("an injected text_encoder cannot cross process boundaries: use workers=1")
        for j in jobs:
            j["text_encoder"] = text_encoder
    print(f"Generated {len(jobs)} jobs. Running with {workers} workers...")
    if workers == 1:
        results = [run_experiment(j) for j in jobs]
    else:
        ctx = multiprocessing.get_context("spawn")
        with ctx.Pool(processes=workers) as pool:
            results = list(pool.imap(run_experiment, jobs))
    os.makedirs(out_dir, exist_ok=True)
    exp_name = config.get("experiment_name", "default_eval")
    safe = "".join(c for c in exp_name if c.isalpha() or c.isdigit() or c in "_-").rstrip()
    out_filename = os.path.join(out_dir, f"{safe}.json")
    with open(out_filename, "w") as f:
        json.dump({"experiment_meta": config, "timestamp": time.time(), "runs": results}, f, indent=4)
    print(f"\n--- Complete. Saved to {out_filename} ---")
    return out_filename, results


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", default="eval_config.json")
    parser.add_argument("--workers", type=int, default=1)
    args = parser.parse_args(argv)
    with open(args.config, "r") as f:
        config = json.load(f)
    print(f"--- Starting Evaluation Manager: {config.get('experiment_name', 'Unnamed')} ---")
    run(config, workers=args.workers)


if __name__ == "__main__":
    main()
