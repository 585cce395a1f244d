"""Multi-GPU optimizer step over peer memory (csrc/p2p.cu) on ONE device: the "peers" are separate buffers on the same GPU,
so the arithmetic (fixed-order peer sums, owned slice, replicated MLP range, fp16 shadow broadcast, GradScaler skip) and
the barrier's flag protocol (epochs, parity slots, OR of the inf bits) are checked without a second process.  The
two-process run over CUDA IPC + NVLink is scripts/check_dist_overlap.py (torchrun, N >= 2)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def L():
    from taichi_nerfs_b200 import _lib
    return _lib.load()


def check(rc):
    from taichi_nerfs_b200 import _lib
    _lib.check(rc)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def table(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_adam_step_p2p_equals_allreduce_plus_adam(world, oracle):
    P, M = 4096 * world, 512                       # table elements (world equal shards), replicated tail
    total = P + M
    rng = np.random.default_rng(5 + world)
    p0 = rng.standard_normal(total).astype(np.float32)
    grads_np = [(rng.standard_normal(total) * 300).astype(np.float32) for _ in range(world)]
    # reference: fp32 sum in rank order, then the oracle's Adam on everything
    gsum = grads_np[0].copy()
    for g in grads_np[1:]:
        gsum = (gsum + g).astype(np.float32)
    p_ref, m_ref, v_ref = p0.copy(), np.zeros(total, np.float32), np.zeros(total, np.float32)
    inv = 1.0 / (1024.0 * world)
    oracle.adam_step(p_ref, gsum.copy(), m_ref, v_ref, 1e-2, 1, inv_scale=inv)

    grads = [torch.from_numpy(g).to(DEV) for g in grads_np]
    shadows = [torch.zeros(total, device=DEV, dtype=torch.float16) for _ in range(world)]
    found = torch.zeros(1, device=DEV, dtype=torch.int32)
    bc1, bc2 = 1 - 0.9, 1 - 0.999
    hyper = torch.tensor([1e-2 / bc1, bc2 ** 0.5, inv, 0.0], device=DEV, dtype=torch.float32)
    shard = P // world
    params = []
    for r in range(world):     # every "rank" runs the kernel on its own master copy / moments
        p = torch.from_numpy(p0).to(DEV)
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        check(L().ngp_adam_step_p2p(ptr(p), table(grads), ptr(m), ptr(v), table(shadows), r, world, ptr(found),
                                    ptr(hyper), 0.9, 0.999, 1e-15, r * shard, (r + 1) * shard, P, total, None))
        params.append(p)
    torch.cuda.synchronize()
    for r in range(world):
        pr = params[r].cpu().numpy()
        own = slice(r * shard, (r + 1) * shard)
        np.testing.assert_allclose(pr[own], p_ref[own], rtol=2e-6, atol=1e-7)          # owned shard: updated
        np.testing.assert_allclose(pr[P:], p_ref[P:], rtol=2e-6, atol=1e-7)            # replicated range: updated
        other = np.ones(total, bool)
        other[own] = False
        other[P:] = False
        assert np.array_equal(pr[other], p0[other])                                    # the rest: untouched (stale master)
        # every rank's shadow: the whole table from the owners' stores + its own replicated range
        sh = shadows[r].cpu().numpy()
        full = np.concatenate([params[q].cpu().numpy()[q * shard:(q + 1) * shard] for q in range(world)])
        assert np.array_equal(sh[:P], full.astype(np.float16))
        assert np.array_equal(sh[P:], pr[P:].astype(np.float16))
    # the replicated range is bit-identical on all ranks (same sums in the same order)
    for r in range(1, world):
        assert torch.equal(params[r][P:], params[0][P:])
    # GradScaler skip: nothing is touched
    found.fill_(1)
    before = params[0].clone()
    sh_before = shadows[0].clone()
    m, v = torch.zeros_like(before), torch.zeros_like(before)
    check(L().ngp_adam_step_p2p(ptr(params[0]), table(grads), ptr(m), ptr(v), table(shadows), 0, world, ptr(found),
                                ptr(hyper), 0.9, 0.999, 1e-15, 0, shard, P, total, None))
    torch.cuda.synchronize()
    assert torch.equal(params[0], before) and torch.equal(shadows[0], sh_before) and not m.any()


def test_p2p_barrier_protocol_two_ranks_on_one_device():
    """Two 'ranks' = two streams of one GPU, each running the barrier kernel against the other's flag block."""
    nwords = int(L().ngp_p2p_flag_bytes()) // 4
    flags = [torch.zeros(nwords, device=DEV, dtype=torch.int32) for _ in range(2)]
    epochs = [torch.zeros(1, device=DEV, dtype=torch.int32) for _ in range(2)]
    infs = [torch.zeros(1, device=DEV, dtype=torch.int32) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    tab = table(flags)
    torch.cuda.synchronize()
    for it in range(6):
        infs[0].fill_(1 if it == 3 else 0)      # rank 0 sees an inf in iteration 3: both must learn it
        infs[1].fill_(0)
        torch.cuda.synchronize()
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                check(L().ngp_p2p_barrier(tab, r, 2, ptr(epochs[r]), ptr(infs[r]), C.c_void_p(streams[r].cuda_stream)))
                check(L().ngp_p2p_barrier(tab, r, 2, ptr(epochs[r]), None, C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        assert int(epochs[0]) == int(epochs[1]) == 2 * (it + 1)
        assert int(infs[0]) == int(infs[1]) == (1 if it == 3 else 0)
    assert int(flags[0][16]) == 0 and int(flags[1][16]) == 0      # no timeout


def test_p2p_barrier_single_rank_and_ipc_roundtrip():
    nwords = int(L().ngp_p2p_flag_bytes()) // 4
    flags = torch.zeros(nwords, device=DEV, dtype=torch.int32)
    epoch = torch.zeros(1, device=DEV, dtype=torch.int32)
    inf = torch.ones(1, device=DEV, dtype=torch.int32)
    for _ in range(3):
        check(L().ngp_p2p_barrier(table([flags]), 0, 1, ptr(epoch), ptr(inf), None))
    torch.cuda.synchronize()
    assert int(epoch) == 3 and int(inf) == 1
    # IPC export of a fresh allocation (opening it needs a second process: scripts/check_dist_overlap.py)
    p, h = C.c_void_p(), (C.c_uint8 * 64)()
    check(L().ngp_p2p_alloc(1 << 20, C.byref(p), h))
    assert p.value and any(bytes(h))
    from taichi_nerfs_b200.p2p import _Raw
    t = torch.as_tensor(_Raw(p.value, 1 << 18, "<f4"), device=DEV)
    assert t.data_ptr() == p.value and not t.any()
    t.fill_(2.0)
    torch.cuda.synchronize()
    del t
    check(L().ngp_p2p_free(p))
