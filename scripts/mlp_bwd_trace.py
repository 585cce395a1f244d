"""Per-round clock trace of the tcgen05 MLP backward v2 (development tool, not part of the product path).

  python scripts/mlp_bwd_trace.py --build    # here: compiles scripts/trace/libngp_trace.so with -DNGP_MLP_TRACE
  python scripts/mlp_bwd_trace.py            # on the GPU: one launch at full size, prints the timeline of CTA 0

Stamps (cycles relative to the first one): issuer per (slot, round): rdy wait done -> MMAs issued + committed;
row-0 thread of every slot per round: accumulator wait done -> epilogue done -> published (fence + arrive)."""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
SO = os.path.join(HERE, "trace", "libngp_trace.so")


def build():
    from taichi_nerfs_b200 import build as b
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = [b._nvcc(), "-ccbin", "/usr/bin/g++"] + b.NVCC_FLAGS + ["-DNGP_MLP_TRACE", "-shared", "-o", SO] + b.sources()
    subprocess.run(cmd, check=True)
    print(SO)


def main():
    import numpy as np
    import torch
    lib = C.CDLL(SO)
    vp, i64 = C.c_void_p, C.c_int64

    class W(C.Structure):
        _fields_ = [(n, vp) for n in ("w1", "w2", "w3", "w4", "w5")]
    n = 148 * 4 * 128 * 6
    dev = "cuda"
    emb = torch.randn(n, 32, device=dev).half()
    dirs = torch.randn(n, 3, device=dev)
    ws = [torch.randn(s, device=dev) * 0.2 for s in ((64, 32), (16, 64), (64, 32), (64, 64), (3, 64))]
    w = W(*[t.data_ptr() for t in ws])
    sig = torch.empty(n, device=dev)
    rgb = torch.empty(n, 3, device=dev, dtype=torch.float16)
    save = torch.zeros(n * 40, device=dev, dtype=torch.uint8)
    dsig = torch.randn(n, device=dev) * 1e-2
    drgb = (torch.randn(n, 3, device=dev) * 1e-2).half()
    demb = torch.empty(n, 32, device=dev, dtype=torch.float16)
    gw = torch.zeros(9408, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.ngp_mlp_fwd_dyn.argtypes = [vp, C.c_int, vp, C.POINTER(W), vp, vp, vp, i64, vp, vp]
    lib.ngp_mlp_bwd_dyn.argtypes = [vp, C.c_int, vp, C.POINTER(W), vp, vp, vp, vp, vp, i64, vp, vp, vp]
    assert lib.ngp_mlp_fwd_dyn(emb.data_ptr(), 1, dirs.data_ptr(), C.byref(w), sig.data_ptr(), rgb.data_ptr(),
                               save.data_ptr(), n, None, st) == 0
    for _ in range(2):
        assert lib.ngp_mlp_bwd_dyn(emb.data_ptr(), 1, dirs.data_ptr(), C.byref(w), save.data_ptr(), dsig.data_ptr(),
                                   drgb.data_ptr(), demb.data_ptr(), gw.data_ptr(), n, None, None, st) == 0
    torch.cuda.synchronize()
    iss, wrk = (C.c_longlong * 128)(), (C.c_longlong * 192)()
    assert lib.ngp_debug_mlp_bwd_trace(iss, wrk) == 0
    I = np.array(iss[:], dtype=np.int64).reshape(4, 2, 8, 2)
    Wk = np.array(wrk[:], dtype=np.int64).reshape(4, 2, 8, 3)
    t0 = Wk[Wk > 0].min()
    names = ["L3", "L4", "R1", "R2", "R3", "L1", "R4", "R5"]
    print("per slot and round: [dX issued @t] -> D ready after +a, epilogue (incl. waiting for a weight-gradient MMA) +e;"
          " dW: operands seen by the issue warp @t, issued in +d")
    for j in range(2):
        print(f"=== tile j = {j + 1} (cycles since the first stamp)")
        for l in range(8):
            line = [f"{names[l]}:"]
            for s in range(3):
                w0, w1, w2 = Wk[s, j, l] - t0
                dw = ""
                if I[s, j, l, 0] > 0:
                    a, b = I[s, j, l] - t0
                    dw = f" dW@{a:6d}+{b - a:4d}"
                line.append(f"s{s} iss@{w2:6d} D+{w0 - w2:4d} epi+{w1 - w0:5d}{dw}")
            print("  ".join(line))


if __name__ == "__main__":
    build() if "--build" in sys.argv else main()
