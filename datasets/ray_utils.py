"""Camera-ray helpers — mirrors get_ray_directions / get_rays of the reference's
datasets/ray_utils.py (:8-48, :51-80) without the kornia dependency.  fp32 even under autocast."""
from __future__ import annotations

import torch


@torch.amp.autocast('cuda', enabled=False)
def get_ray_directions(H, W, K, device='cpu', random=False, return_uv=False, flatten=True):
    """Per-pixel directions in the camera frame [right, down, front]: ((u-cx+.5)/fx, (v-cy+.5)/fy, 1)."""
    v, u = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                          torch.arange(W, device=device, dtype=torch.float32), indexing='ij')
    grid = torch.stack([u, v], dim=-1)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    if random:
        du, dv = torch.rand_like(u), torch.rand_like(v)
    else:
        du = dv = 0.5
    directions = torch.stack([(u - cx + du) / fx, (v - cy + dv) / fy, torch.ones_like(u)], -1)
    if flatten:
        directions = directions.reshape(-1, 3)
        grid = grid.reshape(-1, 2)
    if return_uv:
        return directions, grid
    return directions


@torch.amp.autocast('cuda', enabled=False)
def get_rays(directions, c2w):
    """World-space origins / (un-normalised) directions for camera-frame ``directions`` (N,3) and a
    camera-to-world matrix (3,4) or per-ray matrices (N,3,4)."""
    if c2w.ndim == 2:
        rays_d = directions @ c2w[:, :3].T
    else:
        rays_d = torch.einsum('nc,nac->na', directions, c2w[..., :3])
    rays_o = c2w[..., 3].expand_as(rays_d)
    return rays_o, rays_d
