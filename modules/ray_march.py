"""Occupancy-grid ray marching — mirrors modules/ray_march.py of the reference
(raymarching_train :126-194, raymarching_test :270-334)."""
from __future__ import annotations

import torch

from taichi_nerfs_b200 import ops


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor,
                      grid_size, max_samples, noise=None):
    """Returns (rays_a, xyzs, dirs, deltas, ts, total_samples) like the reference.

    Differences (all documented in DESIGN.md): ``noise`` may be passed explicitly (the reference
    draws it inside, ray_march.py:138); rays_a rows are in ray order with an exclusive-scan start
    (the reference's order/starts come from atomics, :76-81); sample buffers are sized from the
    device counter instead of n_rays*max_samples rows (:149-168).
    """
    rays_o = rays_o.float().contiguous()
    rays_d = rays_d.float().contiguous()
    hits_t = hits_t.contiguous()
    if noise is None:
        noise = torch.rand_like(rays_o[:, 0])
    noise = noise.contiguous()
    counter, rays_a = ops.raymarching_train_count(rays_o, rays_d, hits_t, density_bitfield, noise, cascades, scale,
                                                  exp_step_factor, grid_size, max_samples)
    total = int(counter[0].item())  # same host read-back the reference does at ray_march.py:187-192
    dev = rays_o.device
    xyzs = torch.empty(total, 3, device=dev, dtype=torch.float32)
    dirs = torch.empty(total, 3, device=dev, dtype=torch.float32)
    deltas = torch.empty(total, device=dev, dtype=torch.float32)
    ts = torch.empty(total, device=dev, dtype=torch.float32)
    if total > 0:
        ops.raymarching_train_write(rays_o, rays_d, hits_t, density_bitfield, noise, cascades, scale,
                                    exp_step_factor, grid_size, counter, rays_a, xyzs, dirs, deltas, ts)
    return rays_a, xyzs, dirs, deltas, ts, counter[0]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale,
                     exp_step_factor, grid_size, max_samples):
    """Returns (packed_info, ray_indices, deltas, ts); advances hits_t[:, 0] in place."""
    n_alive = alive_indices.size(0)
    dev = rays_o.device
    m = int(max_samples)
    ray_indices = torch.empty(n_alive * m, device=dev, dtype=torch.long)
    valid_mask = torch.zeros(n_alive * m, device=dev, dtype=torch.uint8)
    deltas = torch.empty(n_alive * m, device=dev, dtype=torch.float32)
    ts = torch.empty(n_alive * m, device=dev, dtype=torch.float32)
    samples_counter = torch.empty(n_alive, device=dev, dtype=torch.int32)
    assert hits_t.is_contiguous() and hits_t.dtype == torch.float32, "hits_t is updated in place"
    ops.raymarching_test(rays_o.float().contiguous(), rays_d.float().contiguous(), hits_t,
                         alive_indices.contiguous(), density_bitfield, cascades, scale, exp_step_factor,
                         grid_size, m, ray_indices, valid_mask, deltas, ts, samples_counter)
    keep = valid_mask.bool()
    ends = torch.cumsum(samples_counter, 0)
    packed_info = torch.stack([ends - samples_counter, samples_counter], dim=-1)
    return packed_info, ray_indices[keep], deltas[keep], ts[keep]
