"""Per-round clock trace of the tcgen05 MLP forward kernel (development tool, not part of the product path).

  python scripts/mlp_round_trace.py --build     # here: compiles scripts/trace/libngp_trace.so with -DNGP_MLP_TRACE
  python scripts/mlp_round_trace.py             # on the GPU: runs the kernel under full occupancy, prints the segments

Timeline slots per round r (thread 0 = MMA issuer / hh0, thread 160 = hh1), cycles of CTA 0:
  6r+0 barrier passed   6r+1 MMAs + commit issued   6r+2 mbarrier wait done   6r+3 epilogue (tcgen05.ld + math + st.shared) done
  6r+4 fences done      6r+5 __syncthreads passed
"""
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
    n = 148 * 4 * 128 * 8
    dev = "cuda"
    emb = torch.randn(n, 32, device=dev).half()
    dirs = torch.randn(n, 3, device=dev)
    ws = [torch.randn(s, device=dev) * 0.2 for s in ((64, 32), (16, 64), (64, 32), (64, 64), (3, 64))]
    w = W(*[t.data_ptr() for t in ws])
    sig = torch.empty(n, device=dev)
    rgb = torch.empty(n, 3, device=dev, dtype=torch.float16)
    lib.ngp_mlp_fwd_dyn.argtypes = [vp, C.c_int, vp, C.POINTER(W), vp, vp, vp, i64, vp, vp]
    for _ in range(3):
        rc = lib.ngp_mlp_fwd_dyn(emb.data_ptr(), 1, dirs.data_ptr(), C.byref(w), sig.data_ptr(), rgb.data_ptr(), None, n, None,
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    torch.cuda.synchronize()
    out = (C.c_longlong * 256)()
    assert lib.ngp_debug_mlp_trace(out) == 0
    t = np.array(out[:], dtype=np.int64).reshape(2, 4, 32)
    names = ["issue", "mma+commit->wake", "epilogue", "fences", "syncthreads"]
    for who, label in ((0, "thread 0 (issuer, hh0)"), (1, "thread 160 (hh1)")):
        print(label)
        for tile in range(1, 4):
            row = t[who, tile]
            segs = []
            for r in range(5):
                b = row[6 * r:6 * r + 6]
                if r < 4:
                    segs.append([int(b[k + 1] - b[k]) for k in range(5)])
                else:
                    segs.append([int(b[1] - b[0]), int(b[2] - b[1]), int(b[3] - b[2]), 0, 0])
            total = int(row[27] - row[31]) if row[27] else 0
            print(f"  tile {tile}: tile-start->layer1 issue {int(row[0] - row[31])}  total {total}")
            for r, s in enumerate(segs):
                print(f"    round {r + 1}: " + "  ".join(f"{nm} {v}" for nm, v in zip(names, s)) + f"   = {sum(s)}")


if __name__ == "__main__":
    build() if "--build" in sys.argv else main()
