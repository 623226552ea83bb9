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
    # engine 3: the same selection on a bf16 shadow of the sums (tcgen05 kind::f16), exact re-scoring unchanged
    if d % 64 == 0:
        f0 = dm.query_stats()["fallbacks"]
        i3, s3 = dm.query(q, top_k=k, normalize=normalize, engine=3)
        torch.cuda.synchronize()
        st3 = dm.query_stats()
        assert st3["fallbacks"] == f0 and k <= st3["last_candidates"] < 65536, st3
        assert torch.equal(i1, i3)
        torch.testing.assert_close(s1, s3, rtol=1e-3, atol=1e-7)


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


@pytest.mark.parametrize("P", [4, 200])
def test_engine2_margin_covers_worst_case_tf32_truncation(P):
    """Pins the selection margin (kTf32Margin).  Every fp32 value below has its 13 low mantissa bits set, so the
    tensor core (which reads 10 mantissa bits) understates each product by 2^-10 .. 2^-9 relative, and the
    voxels are parallel to the prompts, where Cauchy-Schwarz is tight.  70 000 voxels parallel to prompt p have norms
    a relative 2e-6 apart, so their exact ranking is decided far below the TF32 error: a margin smaller than the
    error would drop true top-k voxels; the engine must still agree with the exact engine."""
    import torch
    from vsm import _native as N
    from vsm import voxel_map as vm

    V, d, k = 140000, 128, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(5)

    def low_bits_set(x):  # same sign and exponent, mantissa low 13 bits all ones
        return (x.view(torch.int32) | 0x1FFF).view(torch.float32)

    q = torch.randn((P, d), device="cuda", generator=g).abs() + 0.5       # all products positive
    q = low_bits_set(q / q.norm(dim=1, keepdim=True))
    owner = torch.arange(V, device="cuda") % min(P, 2)                      # half the voxels parallel to prompt 0, half to 1
    scale = 1.0 + 2e-6 * torch.arange(V, device="cuda", dtype=torch.float32)
    feats = low_bits_set(q[owner] * scale[:, None]).contiguous()
    dm = vm.DeviceVoxelMap(0.05, d, N.F32, capacity=V)
    dm.load_dense(torch.rand((V, 3), device="cuda", generator=g), feats)
    i1, s1 = dm.query(q, top_k=k, engine=1)
    i2, s2 = dm.query(q, top_k=k, engine=2)
    assert dm.query_stats()["fallbacks"] == 0
    assert torch.equal(i1, i2)
    torch.testing.assert_close(s1, s2, rtol=1e-6, atol=0)
    # the tensor-core scores really are that far off: float64 vs 10-bit operands
    tf = lambda x: (x.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
    exact = (feats[:1000].double() * q[owner[:1000]].double()).sum(1)
    trunc = (tf(feats[:1000]) * tf(q[owner[:1000]])).sum(1)
    rel = ((exact - trunc) / exact).min().item()
    assert rel > 2.0 ** -10   # per operand between 2^-11 (mantissa near 2) and 2^-10 (near 1): products lose 2^-10 .. 2^-9
    dm.close()


@pytest.mark.parametrize("P", [4, 200])
def test_engine3_margin_covers_worst_case_bf16_rounding(P):
    """Pins kBf16Margin.  bf16 keeps 8 significant bits, rounded to nearest: a value whose mantissa sits just below a
    rounding midpoint loses almost 2^-9 relative.  Voxels and prompts here have every mantissa at 1.xxxxxxx0111...1
    (bit 15 clear, the 15 bits below it set): both operands round DOWN by ~2^-9, the products lose ~2^-8, and the
    voxels are parallel to the prompts where Cauchy-Schwarz is tight.  Norms a relative 2e-6 apart decide the exact
    ranking far below that error; the engine must agree with the exact one, without falling back."""
    import torch
    from vsm import _native as N
    from vsm import voxel_map as vm

    V, d, k = 140000, 128, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(6)

    def just_below_midpoint(x):  # keep sign, exponent and the top 7 mantissa bits; then 0 followed by fifteen 1s
        return ((x.view(torch.int32) & ~0xFFFF) | 0x7FFF).view(torch.float32)

    q = torch.randn((P, d), device="cuda", generator=g).abs() + 0.5
    q = just_below_midpoint(q / q.norm(dim=1, keepdim=True))
    owner = torch.arange(V, device="cuda") % min(P, 2)
    scale = 1.0 + 2e-6 * torch.arange(V, device="cuda", dtype=torch.float32)
    feats = (q[owner] * scale[:, None]).contiguous()     # (the scaling perturbs the low bits: rounding stays ~2^-9 off on average)
    dm = vm.DeviceVoxelMap(0.05, d, N.F32, capacity=V)
    dm.load_dense(torch.rand((V, 3), device="cuda", generator=g), feats)
    i1, s1 = dm.query(q, top_k=k, engine=1)
    i3, s3 = dm.query(q, top_k=k, engine=3)
    assert dm.query_stats()["fallbacks"] == 0
    assert torch.equal(i1, i3)
    torch.testing.assert_close(s1, s3, rtol=1e-6, atol=0)
    # the shadow scores really are that far off
    exact = (feats[:1000].double() * q[owner[:1000]].double()).sum(1)
    shadow = (feats[:1000].bfloat16().double() * q[owner[:1000]].bfloat16().double()).sum(1)
    assert ((exact - shadow).abs() / exact).max().item() > 2.0 ** -10
    dm.close()


def test_engine3_shadow_follows_the_map():
    """The shadow is rebuilt after the sums change (a second fuse + finalise) and can be released."""
    import torch
    from test_gpu_parity import graph_from
    import vsm
    from vsm import synth
    from vsm import _native as N

    subs = [synth.make_submap(93, i, S=6, H=112, W=168, d=64, mode="sl4", room=(6.0, 4.0, 3.0), start=0.3 * i)
            for i in range(4)]
    rng = np.random.default_rng(3)
    q = rng.normal(size=(24, 64)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for n in (2, 4):
        m = graph_from(vsm, subs[:n], device_inputs=True).build_semantic_voxel_map(0.02)
        i1, _, s1 = m.query_with_embeddings(q, top_k=7, engine=1)
        i3, _, s3 = m.query_with_embeddings(q, top_k=7, engine=3)
        np.testing.assert_array_equal(i1, i3)
        np.testing.assert_allclose(s1, s3, rtol=1e-3, atol=1e-7)
        N.check(N.lib.vsm_query_shadow_release(m._dm._h))
        i3b, _, _ = m.query_with_embeddings(q, top_k=7, engine=3)
        np.testing.assert_array_equal(i1, i3b)
    N.set_option("query_shadow", 1)
    try:
        ia, _, _ = m.query_with_embeddings(q, top_k=7)
    finally:
        N.set_option("query_shadow", 0)
    np.testing.assert_array_equal(i1, ia)
