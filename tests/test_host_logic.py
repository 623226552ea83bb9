"""CPU tests of the host-side mirror of the reference interface (no CUDA calls): argument contracts of the Submap
setters (vggt_slam/submap.py:41-65, 109-131), frame-id parsing against the golden name maps produced by the
reference, the lazy containers behind SemanticVoxel, and bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vsm_mod():
    import vsm  # needs the built library (import fails loudly without it), not a GPU

    return vsm


def _bare_submap(vsm, S=2, H=4, W=6):
    sm = vsm.Submap(7)
    sm.pointclouds = np.zeros((S, H, W, 3), dtype=np.float32)   # add_all_points would compute the threshold on the GPU
    sm.conf = np.ones((S, H, W), dtype=np.float32)
    return sm


def test_frame_ids_match_the_reference_name_maps(vsm_mod):
    """set_frame_ids: first number of the basename as float, str(float) -> filename (submap.py:109-131); the golden
    files hold the maps the reference built from the same paths."""
    z = gio.load("case_c_global_sl4.npz")
    want = json.loads(str(z["s1_dedup_names"]))
    for s in gio.inputs(z):
        sm = vsm_mod.Submap(s.submap_id)
        sm.set_frame_ids(s.frame_paths)
        assert sm.frame_id_to_name == want[str(s.submap_id)]
        assert sm.frame_ids == [float(k) for k in sm.frame_id_to_name]
    sm = vsm_mod.Submap(0)
    sm.set_frame_ids(["/data/left_1768798238702970000.png", "img12.5_b.jpg", "a/b/0007.png"])
    assert sm.frame_ids == [1.76879823870297e18, 12.5, 7.0]
    assert sm.frame_id_to_name == {"1.76879823870297e+18": "left_1768798238702970000.png", "12.5": "img12.5_b.jpg",
                                   "7.0": "0007.png"}
    with pytest.raises(ValueError, match="No number found"):
        sm.set_frame_ids(["nodigits.png"])


def test_embedding_argument_contract(vsm_mod):
    """TypeError / ValueError exactly where the reference raises them (submap.py:52-63)."""
    sm = _bare_submap(vsm_mod)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings([[1.0]])
    with pytest.raises(ValueError, match="4 dims"):
        sm.add_all_semantic_embeddings(np.zeros((2, 4, 6), dtype=np.float32))
    with pytest.raises(ValueError, match="spatial dims must match"):
        sm.add_all_semantic_embeddings(np.zeros((2, 4, 5, 8), dtype=np.float32))
    sm.add_all_semantic_embeddings(np.zeros((2, 4, 6, 8), dtype=np.float32))
    assert sm.semantic_index is None and sm.semantic_embeddings.shape == (2, 4, 6, 8)
    sm.add_all_semantic_embeddings(None)
    assert sm.semantic_embeddings is None
    # indexed form: same exception types
    ids, table = np.zeros((2, 4, 6), dtype=np.int16), np.zeros((3, 8), dtype=np.float32)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings_indexed(ids.astype(np.float32), table)
    with pytest.raises(TypeError):
        sm.add_all_semantic_embeddings_indexed(ids.tolist(), table)
    with pytest.raises(ValueError, match="3 dims"):
        sm.add_all_semantic_embeddings_indexed(ids[0], table)
    with pytest.raises(ValueError, match="2 dims"):
        sm.add_all_semantic_embeddings_indexed(ids, table[0])
    with pytest.raises(ValueError, match="spatial dims must match"):
        sm.add_all_semantic_embeddings_indexed(ids[:, :, :5], table)
    sm.add_all_semantic_embeddings_indexed(ids, table)
    assert sm.semantic_index is ids and sm.semantic_embeddings is table
    table[1] = 2.0
    ids[1, 2, 3] = 1
    dense = sm.dense_semantic_embeddings()
    assert dense.shape == (2, 4, 6, 8) and dense[1, 2, 3, 0] == 2.0 and dense.sum() == 16.0
    sm.add_all_semantic_embeddings(dense)           # a dense call clears the index again
    assert sm.semantic_index is None


def test_lazy_containers():
    from vsm.semantic_voxel import LazyContributors, SemanticVoxel

    calls = []

    def maker(i):
        calls.append(i)
        return [(0, "2.0"), (1, "10.0"), (1, "9.0")] if i == 1 else [(0, str(float(i)))]

    lc = LazyContributors(3, maker)
    assert len(lc) == 3 and calls == []
    assert lc[1] == [(0, "2.0"), (1, "10.0"), (1, "9.0")] and calls == [1]
    lc[1].sort(reverse=True)                         # get_latest_frame_at_voxel sorts in place (semantic_voxel.py:124)
    assert lc[1][0] == (1, "9.0") and calls == [1]   # string order: '9.0' > '10.0'; the mutation persists, no rebuild
    assert lc[-1] == [(0, "2.0")] and lc[0:2] == [[(0, "0.0")], lc[1]]
    assert lc == [[(0, "0.0")], lc[1], [(0, "2.0")]] and lc.tolist() == list(lc)
    with pytest.raises(IndexError):
        lc[3]
    fetched = []

    def centers():
        fetched.append(1)
        return np.arange(9, dtype=np.float32).reshape(3, 3)

    v = SemanticVoxel.lazy(0.05, centers, lambda: np.ones((3, 8), dtype=np.float32), lc)
    assert len(v) == 3 and fetched == [] and "n_voxels=3" in repr(v)
    assert v.centers_world.shape == (3, 3) and v.centers_world is v.centers_world and fetched == [1]
    assert v.features.shape == (3, 8)
    v.centers_world = np.zeros((2, 3), dtype=np.float32)
    assert len(v) == 2
    w = SemanticVoxel(0.1, np.zeros((0, 3), dtype=np.float32), np.zeros((0, 0), dtype=np.float32), [])
    assert len(w) == 0 and w.features.shape == (0, 0)


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` times the oracle port and prints ONE json line with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--frames", "2", "--cpu-frames", "2", "--height", "28", "--width", "42", "--dim", "16", "--voxel-size", "0.5", "--cpu-procs", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "points fused/sec" and j["unit"] == "points/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 1
    par = j["cpu_baseline"]["parallel"]              # independent copies of the port, one per core (an extra figure)
    assert par["procs"] >= 1 and par["value"] > 0 and par["unit"] == "points/s"
    assert j["e2e"] == {"value": j["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the record says what actually ran: one submap of `frames` frames per step, not the GPU arm's workload
    assert j["config"]["frames"] == 2 and j["config"]["submaps_per_gpu"] == 1 and j["config"]["same_config_as_gpu_arm"] is False
    assert "1 submap x 2 frames" in j["config"]["workload"] and "1 submap x 2 frames" in j["cpu_baseline"]["sample"]


def test_clock_sampler_without_nvml_reports_instead_of_failing():
    sys.path.insert(0, ROOT)
    import bench

    s = bench.ClockSampler(0)
    s.start()            # no GPU / no NVML here: the error is recorded
    s.mark_begin()
    s.mark_end()
    r = s.stop()
    assert r["samples"] == 0 and r["sm_mhz"] is None and r["reasons"] == []
