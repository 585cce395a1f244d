"""Test-time incremental compositing — mirrors modules/volume_render_test.py."""
from taichi_nerfs_b200 import ops


def composite_test(sigmas, rgbs, deltas, ts, pack_info, alive_indices, T_threshold, opacity, depth, rgb):
    """In-place accumulation into opacity/depth/rgb; converged rays get alive_indices[n] = -1
    (reference kernel: volume_render_test.py:4-54)."""
    ops.composite_test(sigmas.float().contiguous(), rgbs.contiguous(), deltas.contiguous(), ts.contiguous(),
                       pack_info.contiguous(), alive_indices, T_threshold, opacity, depth, rgb)
