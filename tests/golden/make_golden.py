"""Generates the small golden fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (committed):
  lego_bitfield.npz      — the trained Lego occupancy bitfield shipped with the reference's mobile
                           demo (deployment/InstantNGP/taichi_ngp/compiled/density_bitfield.bin,
                           128^3 bits, 3.94 % occupied), zlib-compressed.  Used as the "occupancy (A)"
                           workload of BASELINE.md §4 and as a real-world marching fixture.
  layout_constants.json  — hash-layout constants printed by the reference itself
                           (notebooks/pipeline.ipynb cell 1; deployment/InstantNGP/utils/app_fp32.cpp:70-71).
"""
import json
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_bin(path):
    """[int32 dtype][int32 numel][payload] container (deployment/InstantNGP/taichi_ngp/taichi_ngp.py:34-65)."""
    raw = np.fromfile(path, dtype=np.uint8)
    dtype_code, numel = raw[:8].view(np.int32)
    np_dtype = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.int16, 4: np.uint32, 5: np.uint16}[int(dtype_code)]
    return raw[8:].view(np_dtype)[:numel]


def main():
    comp = os.path.join(REF, "deployment/InstantNGP/taichi_ngp/compiled")
    bits = read_bin(os.path.join(comp, "density_bitfield.bin")).view(np.uint8)
    assert bits.size == 128 ** 3 // 8
    np.savez_compressed(os.path.join(HERE, "lego_bitfield.npz"), bitfield=bits)
    pose = read_bin(os.path.join(comp, "pose.bin")).reshape(3, 4)

    # constants the reference prints / hard-codes
    consts = {
        "source": {
            "lego_16_1024": "notebooks/pipeline.ipynb cell 1 (per_level_scale, offset_, total_hash_size)",
            "deployment": "deployment/InstantNGP/utils/app_fp32.cpp:70-71, taichi_ngp/kernels.py (offsets)",
        },
        "lego_16_1024": {"per_level_scale": 1.3195079107728942, "total_entries": 5710032,
                         "total_params": 11420064},
        "deployment": {"total_params": 11176096, "offsets_entries": [0, 32768, 165424, 696872]},
        "deployment_pose": pose.tolist(),
        "bitfield_occupied_fraction": float(np.unpackbits(bits).mean()),
    }
    with open(os.path.join(HERE, "layout_constants.json"), "w") as f:
        json.dump(consts, f, indent=1)
    print("bitfield occupied:", consts["bitfield_occupied_fraction"])


if __name__ == "__main__":
    main()
