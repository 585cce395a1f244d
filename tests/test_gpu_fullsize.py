"""GPU parity at BASELINE.json's full size (configs[1]: 8192 rays, ~50 % occupancy -> ~2 M samples/step,
L=16 T=2^19 F=2 fp16 table, the stock MLP): direct comparison with the (OpenMP) oracle where it finishes in
seconds, plus the size-independent properties of the domain — sample ordering / partition, adjointness and
linearity of the hash backward, weight sums of the compositor, Morton round trip over the whole grid."""
import numpy as np
import pytest
import torch

from taichi_nerfs_b200.layout import make_hash_layout

pytestmark = pytest.mark.gpu

DEV = "cuda"
N_RAYS = 8192


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from taichi_nerfs_b200 import ops as o
    return o


@pytest.fixture(scope="module")
def march(oracle, rays_factory):
    """One full-size marching result (oracle side), shared by the tests below."""
    rng = np.random.default_rng(101)
    o, d = rays_factory(N_RAYS, seed=101)
    bits = rng.integers(0, 256, 128 ** 3 // 8, dtype=np.uint8)      # ~50 % occupied, as bench.py's workload
    hits = oracle.ray_aabb_intersect(o, d, 0.5)
    noise = rng.random(N_RAYS, dtype=np.float32)
    ra, xyzs, dirs, deltas, ts, S = oracle.raymarching_train(o, d, hits, bits, noise, 1, 0.5, 0.0, 128, 1024)
    assert S > 1_500_000
    return dict(o=o, d=d, bits=bits, hits=hits, noise=noise, rays_a=ra, xyzs=xyzs, dirs=dirs, deltas=deltas, ts=ts, S=S)


def test_fullsize_march_bit_exact_and_ordered(ops, march):
    from modules.ray_march import raymarching_train
    m = march
    g_ra, g_xyzs, g_dirs, g_deltas, g_ts, g_total = raymarching_train(
        T(m["o"]), T(m["d"]), T(m["hits"]), T(m["bits"]), 1, 0.5, 0.0, 128, 1024, noise=T(m["noise"]))
    assert int(g_total) == m["S"]
    assert np.array_equal(N(g_ra), m["rays_a"])
    for a, b in ((g_xyzs, m["xyzs"]), (g_dirs, m["dirs"]), (g_deltas, m["deltas"]), (g_ts, m["ts"])):
        assert np.array_equal(N(a).view(np.uint32), b.view(np.uint32))
    # size-independent properties: rays_a partitions [0, S); t strictly increases inside a ray; deltas > 0
    ra = N(g_ra).astype(np.int64)
    order = np.argsort(ra[:, 1], kind="stable")
    starts, counts = ra[order, 1], ra[order, 2]
    assert starts[0] == 0 and np.array_equal(starts[1:], np.cumsum(counts)[:-1]) and counts.sum() == m["S"]
    ts = N(g_ts)
    inside = np.ones(m["S"], bool)
    inside[starts[counts > 0]] = False                      # first sample of every ray
    assert (np.diff(ts)[inside[1:]] > 0).all()
    assert (N(g_deltas) > 0).all()
    # xyz = o + t * d, recomputed without FMA (separate torch kernels)
    ray_of = torch.repeat_interleave(g_ra[:, 0].long(), g_ra[:, 2].long())
    assert torch.equal(g_xyzs, T(m["o"])[ray_of] + g_ts[:, None] * T(m["d"])[ray_of])


def test_fullsize_single_pass_march_is_a_permutation_of_blocks(ops, march):
    """ngp_raymarching_frame (atomic row reservation): every ray's block equals the oracle's block for that ray."""
    m = march
    counter = torch.zeros(2, device=DEV, dtype=torch.int32)
    cap = m["S"] + 4096
    rays_a = torch.zeros(N_RAYS, 3, device=DEV, dtype=torch.int32)
    xyzs, dirs = torch.zeros(cap, 3, device=DEV), torch.zeros(cap, 3, device=DEV)
    deltas, ts = torch.zeros(cap, device=DEV), torch.zeros(cap, device=DEV)
    ops.raymarching_frame(T(m["o"]), T(m["d"]), T(m["hits"]), T(m["bits"]), 1, 0.5, 0.0, 128, 1024, counter, rays_a,
                          xyzs, dirs, deltas, ts, noise=T(m["noise"]))
    assert int(counter[0]) == m["S"] and int(counter[1]) == 0
    ra = N(rays_a).astype(np.int64)
    ref = m["rays_a"].astype(np.int64)
    assert np.array_equal(ra[:, 0], ref[:, 0]) and np.array_equal(ra[:, 2], ref[:, 2])
    # gather the GPU rows into the oracle's order and compare bit-exactly
    idx = np.concatenate([np.arange(s, s + c) for s, c in ra[:, 1:3]])
    ref_idx = np.concatenate([np.arange(s, s + c) for s, c in ref[:, 1:3]])
    assert np.array_equal(np.sort(idx), np.arange(m["S"]))                      # rows reserved exactly once
    for a, b in ((ts, m["ts"]), (deltas, m["deltas"]), (xyzs, m["xyzs"])):
        assert np.array_equal(N(a)[idx].view(np.uint32), b[ref_idx].view(np.uint32))


def test_fullsize_hash_fwd_bit_exact_bwd_adjoint_and_linear(ops, oracle, march):
    rng = np.random.default_rng(102)
    lay = make_hash_layout(2 ** 19, 16, 16, 1024, 2)
    xn = (march["xyzs"] + 0.5).astype(np.float32)                              # (x - xyz_min) / (xyz_max - xyz_min)
    S = xn.shape[0]
    table16 = ((rng.random(lay.total_param_size, dtype=np.float32) * 2 - 1) * 1e-1).astype(np.float16)
    ref = oracle.hash_encode_fwd(xn, table16, lay)
    got = ops.hash_encode_fwd(T(xn), T(table16), lay.as_ctypes(), lay.out_dim)
    assert np.array_equal(N(got).view(np.uint16), ref.view(np.uint16))        # all ~2 M x 32 outputs
    # in-kernel AABB normalisation (the graph step's variant) is the same function
    aabb = (-0.5, -0.5, -0.5, 1.0, 1.0, 1.0)                                   # xyz_min, xyz_max - xyz_min
    got2 = ops.hash_encode_fwd(T(march["xyzs"]), T(table16), lay.as_ctypes(), lay.out_dim, aabb=aabb)
    assert torch.equal(got, got2)

    # adjointness with an fp32 table: <fwd(x; W), dY> == <W, bwd(x; dY)>  (the encoder is linear in W)
    W = rng.standard_normal(lay.total_param_size).astype(np.float32)
    dY = rng.standard_normal((S, 32)).astype(np.float32)
    y = ops.hash_encode_fwd(T(xn), T(W), lay.as_ctypes(), lay.out_dim)
    g = torch.zeros(lay.total_param_size, device=DEV)
    ops.hash_encode_bwd(T(xn), T(dY), lay.as_ctypes(), g)
    lhs = float((y.double() * T(dY).double()).sum())
    rhs = float((T(W).double() * g.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), float((y.double() * T(dY).double()).abs().sum()) * 1e-3)
    # linearity in dY: bwd(2.5 dY1 + dY2) == 2.5 bwd(dY1) + bwd(dY2)  (fp32 atomics, arbitrary order: 1e-3 of the scale)
    dY2 = rng.standard_normal((S, 32)).astype(np.float32)
    g2 = torch.zeros_like(g)
    ops.hash_encode_bwd(T(xn), T(dY2), lay.as_ctypes(), g2)
    g3 = torch.zeros_like(g)
    ops.hash_encode_bwd(T(xn), T(2.5 * dY + dY2), lay.as_ctypes(), g3)
    err = (g3 - (2.5 * g + g2)).abs().max()
    assert float(err) <= 1e-3 * float(g3.abs().max())


def test_fullsize_mlp_matches_oracle(ops, oracle, march):
    rng = np.random.default_rng(103)
    S = march["S"]
    emb = rng.standard_normal((S, 32)).astype(np.float16)
    shapes = [(64, 32), (16, 64), (64, 32), (64, 64), (3, 64)]
    ws = [(rng.uniform(-1, 1, s) * np.sqrt(6 / (s[0] + s[1]))).astype(np.float32) for s in shapes]
    sig_ref, rgb_ref = oracle.mlp_fwd(emb, march["dirs"], ws)
    sig, rgb = ops.mlp_fwd(T(emb), T(march["dirs"]), [T(w) for w in ws])
    sig, rgb = N(sig), N(rgb).astype(np.float32)
    # same tolerances as the small-size test (tests/test_gpu_parity.py::test_mlp_fwd_tcgen05_matches_oracle)
    np.testing.assert_allclose(sig, sig_ref, rtol=8e-3)
    assert np.abs(rgb - rgb_ref.astype(np.float32)).max() <= 2e-3
    assert np.median(np.abs(sig - sig_ref) / sig_ref) < 1e-3
    # backward, full size: weight gradients are sums over ~2 M samples -> compare relative to their scale
    dsig = (rng.standard_normal(S) * 1e-2).astype(np.float32)
    drgb = (rng.standard_normal((S, 3)) * 1e-2).astype(np.float16)
    demb_ref, gw_ref = oracle.mlp_bwd(emb, march["dirs"], ws, dsig, drgb)
    _, _, save = ops.mlp_fwd(T(emb), T(march["dirs"]), [T(w) for w in ws], with_save=True)
    demb, gw = ops.mlp_bwd(T(emb), T(march["dirs"]), [T(w) for w in ws], T(dsig), T(drgb), save=save)
    demb2, gw2 = ops.mlp_bwd(T(emb), T(march["dirs"]), [T(w) for w in ws], T(dsig), T(drgb))     # recompute variant
    assert float((demb.float() - demb2.float()).abs().max()) <= 2e-3 * float(demb2.float().abs().max())
    assert float((gw - gw2).abs().max()) <= 2e-3 * float(gw2.abs().max())
    demb, gw = N(demb).astype(np.float32), N(gw)
    scale = np.abs(demb_ref.astype(np.float32)).max()
    assert np.abs(demb - demb_ref.astype(np.float32)).max() <= 5e-3 * scale
    offs = np.cumsum([0, 2048, 1024, 2048, 4096, 192])                        # w1 | w2 | w3 | w4 | w5
    for a, b in zip(offs[:-1], offs[1:]):
        assert np.abs(gw[a:b] - gw_ref[a:b]).max() <= 5e-3 * np.abs(gw_ref[a:b]).max(), (a, b)


def test_fullsize_composite_matches_oracle_and_weights_sum_to_opacity(ops, oracle, march):
    rng = np.random.default_rng(104)
    m, S = march, march["S"]
    sig = (rng.random(S) * 8.0).astype(np.float32)
    rgbs = rng.random((S, 3)).astype(np.float16)
    tot, op, dep, rgb, ws = oracle.composite_train_fwd(sig, rgbs, m["deltas"], m["ts"], m["rays_a"], 1e-4)
    g_tot, g_op, g_dep, g_rgb, g_ws = ops.composite_train_fwd(T(sig), T(rgbs), T(m["deltas"]), T(m["ts"]),
                                                              T(m["rays_a"]), 1e-4)
    np.testing.assert_allclose(N(g_op), op, atol=2e-5)
    np.testing.assert_allclose(N(g_rgb), rgb, atol=2e-5)
    np.testing.assert_allclose(N(g_dep), dep, atol=5e-5)
    # properties: weights of a ray sum to its opacity, opacity in [0, 1], colours bounded by opacity
    ray_of = torch.repeat_interleave(T(m["rays_a"][:, 0]).long(), T(m["rays_a"][:, 2]).long())
    wsum = torch.zeros(N_RAYS, device=DEV, dtype=torch.float64).index_add_(0, ray_of, g_ws.double())
    assert float((wsum - g_op.double()).abs().max()) < 1e-5
    assert float(g_op.min()) >= 0.0 and float(g_op.max()) <= 1.0 + 1e-6
    assert bool((g_rgb <= g_op[:, None] + 1e-5).all())


def test_fullsize_adam_and_grid_helpers(ops, oracle):
    """Fused Adam over the whole parameter vector vs torch.optim.Adam (eps 1e-15, train.py:137-142); Morton round
    trip and packbits over the whole 128^3 grid."""
    P = 11_420_064 + 9408
    g = torch.Generator(device=DEV).manual_seed(5)
    p0 = torch.randn(P, device=DEV, generator=g) * 1e-2
    grad = torch.randn(P, device=DEV, generator=g) * 1e-3
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15)
    p, m, v = p0.clone(), torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
    for step in (1, 2):
        ref.grad = grad.clone()
        opt.step()
        gbuf = (grad * 1024.0).clone()                                       # scaled gradient, unscaled in-kernel
        ops.adam_step(p, gbuf, m, v, lr=1e-2, step=step, inv_scale=1 / 1024.0, zero_grad=True)
        assert float(gbuf.abs().max()) == 0.0                                # zero_grad fused
    assert float((p - ref.detach()).abs().max()) <= 1e-6 + 1e-5 * 1e-2 * 2   # lr-scaled: two steps of <= lr each

    coords = torch.stack(torch.meshgrid(*[torch.arange(128, device=DEV, dtype=torch.int32)] * 3, indexing="ij"),
                         -1).reshape(-1, 3)
    idx = ops.morton3d(coords)
    assert torch.equal(torch.sort(idx.long()).values, torch.arange(128 ** 3, device=DEV))   # a bijection
    assert torch.equal(ops.morton3d_invert(idx), coords)
    grid = torch.rand(128 ** 3, device=DEV, generator=g)
    bits = torch.zeros(128 ** 3 // 8, device=DEV, dtype=torch.uint8)
    ops.packbits(grid, 0.37, bits)
    assert np.array_equal(N(bits), oracle.packbits(N(grid), 0.37))
    assert np.array_equal(np.unpackbits(N(bits), bitorder="little").astype(bool), N(grid) > 0.37)
