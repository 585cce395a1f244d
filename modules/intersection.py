"""Ray / AABB intersection — mirrors modules/intersection.py of the reference."""
from taichi_nerfs_b200 import ops

from .utils import NEAR_DISTANCE  # noqa: F401  (re-exported like the reference)


def ray_aabb_intersection(rays_o, rays_d, scale):
    """hits_t[r] = (max(t_near, 0.01), t_far) against the cube [-scale, scale]^3, or (-1, -1)
    (reference: ray_aabb_intersect kernel, modules/intersection.py:8-37)."""
    return ops.ray_aabb_intersect(rays_o, rays_d, scale)
