"""Isolated timing of the tcgen05 MLP kernels (CUDA events, L2 flushed between repeats): forward with / without the
save buffer, backward restarting from the saved activations vs recomputing everything."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from taichi_nerfs_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_190_000
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
emb = torch.randn(n, 32, device=dev, generator=g).half()
dirs = torch.randn(n, 3, device=dev, generator=g)
ws = [torch.randn(s, device=dev, generator=g) * 0.2 for s in ((64, 32), (16, 64), (64, 32), (64, 64), (3, 64))]
dsig = torch.randn(n, device=dev, generator=g) * 1e-2
drgb = (torch.randn(n, 3, device=dev, generator=g) * 1e-2).half()
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
_, _, save = ops.mlp_fwd(emb, dirs, ws, with_save=True)


def t(fn, reps=7):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out)


for name, fn in [("fwd", lambda: ops.mlp_fwd(emb, dirs, ws)), ("fwd+save", lambda: ops.mlp_fwd(emb, dirs, ws, with_save=True)),
                 ("bwd recompute", lambda: ops.mlp_bwd(emb, dirs, ws, dsig, drgb)),
                 ("bwd saved", lambda: ops.mlp_bwd(emb, dirs, ws, dsig, drgb, save=save))]:
    fn()
    print(f"{name:14s} {t(fn) * 1e3:8.1f} us   (n = {n})")
