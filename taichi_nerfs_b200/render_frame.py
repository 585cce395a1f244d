"""Test-time frame rendering without the host-driven while-loop.

The reference renders a frame (gui.py:115-145 -> modules/rendering.py:61-158) by repeatedly marching
<= N samples per alive ray, compacting with boolean masks (host syncs), running the network and
compositing incrementally, until every ray has terminated — hundreds of tiny launches and >= 3 host
syncs per iteration.  Compositing is sequential per ray and stops at the first sample where the
transmittance falls to <= T_threshold, so the result does not depend on how the samples are chunked.
Here the frame is rendered in a few big launches per ray block: march ALL samples of the block
(warp-per-ray march, deterministic layout), encode + MLP them in one pass each, composite with the
warp-per-ray kernel that applies the same early termination.  Samples behind an opaque surface are
evaluated but ignored (the price of no per-iteration sync); the rendered rgb/depth/opacity equal the
loop's up to fp rounding.
"""
from __future__ import annotations

import torch

from . import ops
from .fused_mlp import mlp_weights


@torch.no_grad()
def render_frame(model, rays_o, rays_d, exp_step_factor=0.0, T_threshold=1e-4, max_samples=1024,
                 block_rays=1 << 18):
    """Returns the dict of rendering.render(test_time=True): rgb [N,3], depth [N], opacity [N], total_samples."""
    dev = rays_o.device
    n = rays_o.shape[0]
    rays_o = rays_o.float().contiguous()
    rays_d = rays_d.float().contiguous()
    rgb = torch.empty(n, 3, device=dev, dtype=torch.float32)
    depth = torch.empty(n, device=dev, dtype=torch.float32)
    opacity = torch.empty(n, device=dev, dtype=torch.float32)
    enc = model.pos_encoder
    half = hasattr(enc, "table_f16")
    fused = bool(model._fusable(rays_o))     # stock architecture: hash + tcgen05 MLP kernels on the raw sample rows
    if fused:
        table = enc.table_f16() if half else enc.hash_table.detach().contiguous()
        W = [w.detach() for w in mlp_weights(model)]
    aabb = model.xyz_min.flatten().tolist() + (model.xyz_max - model.xyz_min).flatten().tolist()
    total = 0
    zeros = torch.zeros(min(block_rays, n), device=dev, dtype=torch.float32)  # test-time march has no jitter
    caps = model.__dict__.setdefault('_frame_capacity', {})  # learned per ray block from the previous frame
    for b in range(0, n, block_rays):
        e = min(b + block_rays, n)
        o, d = rays_o[b:e], rays_d[b:e]
        hits = ops.ray_aabb_intersect(o, d, model.scale)
        key = (b, e - b)
        done = False
        if key in caps and max_samples <= 1024:
            # single-pass march: each ray reserves its rows with one atomic; capacity from the last frame
            cap = caps[key]
            counter = torch.zeros(2, device=dev, dtype=torch.int32)
            rays_a = torch.empty(e - b, 3, device=dev, dtype=torch.int32)
            xyzs = torch.empty(cap, 3, device=dev, dtype=torch.float32)
            dirs = torch.empty(cap, 3, device=dev, dtype=torch.float32)
            deltas = torch.empty(cap, device=dev, dtype=torch.float32)
            ts = torch.empty(cap, device=dev, dtype=torch.float32)
            ops.raymarching_frame(o, d, hits, model.density_bitfield, model.cascades, model.scale, exp_step_factor,
                                  model.grid_size, max_samples, counter, rays_a, xyzs, dirs, deltas, ts)
            S, dropped = counter.tolist()  # one host read per ray block
            caps[key] = max(int(S * 1.25) + 4096, 1 << 16)
            if dropped == 0:
                done = True
                xyzs, dirs, deltas, ts = xyzs[:S], dirs[:S], deltas[:S], ts[:S]
        if not done:
            # exact two-pass march (count -> scan -> write); also used for the first frame to learn S
            noise = zeros[: e - b]
            counter, rays_a = ops.raymarching_train_count(o, d, hits, model.density_bitfield, noise, model.cascades,
                                                          model.scale, exp_step_factor, model.grid_size, max_samples)
            S = int(counter[0].item())
            caps[key] = max(int(S * 1.25) + 4096, 1 << 16)
            if S > 0:
                xyzs = torch.empty(S, 3, device=dev, dtype=torch.float32)
                dirs = torch.empty(S, 3, device=dev, dtype=torch.float32)
                deltas = torch.empty(S, device=dev, dtype=torch.float32)
                ts = torch.empty(S, device=dev, dtype=torch.float32)
                ops.raymarching_train_write(o, d, hits, model.density_bitfield, noise, model.cascades, model.scale,
                                            exp_step_factor, model.grid_size, counter, rays_a, xyzs, dirs, deltas, ts)
        total += S
        if S == 0:
            opacity[b:e] = 0
            depth[b:e] = 0
            rgb[b:e] = 0
            continue
        if fused:
            emb = ops.hash_encode_fwd(xyzs, table, enc._clayout, enc.out_dim, aabb=aabb)  # normalisation in-kernel
            sigmas, rgbs = ops.mlp_fwd(emb, dirs, W)
        else:   # any other NGP configuration (e.g. the reference's L=4 F=4 deployment model): module forward
            with torch.autocast('cuda', dtype=torch.float16):
                outs = [model(xyzs[i:i + (1 << 21)], dirs[i:i + (1 << 21)]) for i in range(0, S, 1 << 21)]
            sigmas = torch.cat([o[0] for o in outs]).float().contiguous()
            rgbs = torch.cat([o[1] for o in outs]).contiguous()
        _, op_b, dp_b, rgb_b, _ = ops.composite_train_fwd(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
        opacity[b:e] = op_b
        depth[b:e] = dp_b
        rgb[b:e] = rgb_b
    bg = 1.0 if exp_step_factor == 0 else 0.0  # rendering.py:152-156
    if bg:
        rgb += bg * (1 - opacity)[:, None]
    return {'opacity': opacity, 'depth': depth, 'rgb': rgb, 'total_samples': torch.tensor(total, device=dev)}
