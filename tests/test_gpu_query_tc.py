"""Tensor-core query engine (tcgen05, engine 2) against the exact fp32 engine (engine 1) and the oracle: identical
top-k indices, scores within 1e-3 relative (they are the exactly re-scored values, so they agree to fp32 rounding)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_map(V, d, seed, device="cuda"):
    import torch
    from vsm import _native as N
    from vsm import voxel_map as vm

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    feats = torch.randn((V, d), dtype=torch.float32, device=device, generator=g)
    feats = feats / feats.norm(dim=1, keepdim=True) * (0.3 + 0.7 * torch.rand((V, 1), device=device, generator=g))
    centers = torch.rand((V, 3), dtype=torch.float32, device=device, generator=g) * 50.0
    dm = vm.DeviceVoxelMap(0.05, d, N.F32, capacity=V)
    dm.load_dense(centers, feats)
    return dm, feats


@pytest.mark.parametrize("V,d,P,k", [(70000, 512, 16, 10), (300001, 512, 64, 5), (200000, 64, 100, 20), (131072, 256, 9, 1),
                                     (150001, 512, 256, 10), (100000, 256, 200, 3), (90000, 512, 128, 4)])
@pytest.mark.parametrize("normalize", [False, True])
def test_engine2_equals_engine1(V, d, P, k, normalize):
    import torch

    dm, feats = _random_map(V, d, 1000 + V % 97)
    rng = np.random.default_rng(V + P)
    q = rng.normal(size=(P, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    i1, s1 = dm.query(q, top_k=k, normalize=normalize, engine=1)
    i2, s2 = dm.query(q, top_k=k, normalize=normalize, engine=2)
    torch.cuda.synchronize()
    st = dm.query_stats()
    assert st["fallbacks"] == 0, st          # the tensor-core path answered, not the fallback
    assert k <= st["last_candidates"] < 65536, st
    assert torch.equal(i1, i2)
    torch.testing.assert_close(s1, s2, rtol=1e-3, atol=1e-7)
    # and against a float64 torch restatement for a few prompts
    f = feats.double()
    if normalize:
        f = f / f.norm(dim=1, keepdim=True).clamp_min(1e-12)
    ref = torch.topk(f @ torch.from_numpy(q[:3]).double().cuda().T, k, dim=0).indices.T
    assert torch.equal(ref, i2[:3])
    ia, _ = dm.query(q, top_k=k, normalize=normalize, engine=0)   # auto picks the tensor cores for P > 8
    assert torch.equal(ia, i1)


def test_engine2_on_fused_map_and_overflow_fallback():
    """On a map built by fusion (sums + counts, ids != ranks); then a map of identical rows, whose ties overflow the
    candidate lists and must fall back to the exact engine."""
    import torch
    from test_gpu_parity import graph_from
    import vsm
    from vsm import synth
    from vsm import _native as N
    from vsm import voxel_map as vm

    subs = [synth.make_submap(91, i, S=6, H=112, W=168, d=64, mode="sl4", room=(6.0, 4.0, 3.0), start=0.3 * i)
            for i in range(3)]
    m = graph_from(vsm, subs, device_inputs=True).build_semantic_voxel_map(0.02)
    assert m._dm.num_voxels > 20000
    rng = np.random.default_rng(2)
    q = rng.normal(size=(24, 64)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    i1, _, s1 = m.query_with_embeddings(q, top_k=7, engine=1)
    i2, _, s2 = m.query_with_embeddings(q, top_k=7, engine=2)
    assert m._dm.query_stats()["fallbacks"] == 0
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_allclose(s1, s2, rtol=1e-3, atol=1e-7)

    V, d = 100000, 64
    feats = torch.ones((V, d), device="cuda") / 8.0
    dm = vm.DeviceVoxelMap(0.05, d, N.F32, capacity=V)
    dm.load_dense(torch.rand((V, 3), device="cuda"), feats)
    i1, s1 = dm.query(q[:12], top_k=3, engine=1)
    i2, s2 = dm.query(q[:12], top_k=3, engine=2)
    assert dm.query_stats()["fallbacks"] == 1
    assert torch.equal(i1, i2) and i1[0].tolist() == [0, 1, 2]      # ties: lower index first
