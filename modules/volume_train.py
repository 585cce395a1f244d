"""Differentiable volume-rendering compositing — mirrors modules/volume_train.py."""
import torch

from taichi_nerfs_b200 import ops


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        total, opacity, depth, rgb, ws = ops.composite_train_fwd(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
        ctx.T_threshold = T_threshold
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a)
        ctx.mark_non_differentiable(total)
        return total, opacity, depth, rgb, ws

    @staticmethod
    def backward(ctx, _g_total, g_opacity, g_depth, g_rgb, g_ws):
        sigmas, rgbs, deltas, ts, rays_a = ctx.saved_tensors
        n = rays_a.shape[0]
        dev = sigmas.device

        def dense(g, shape):
            return torch.zeros(shape, device=dev, dtype=torch.float32) if g is None else g.float().contiguous()

        dsig, drgbs = ops.composite_train_bwd(dense(g_opacity, (n,)), dense(g_depth, (n,)), dense(g_rgb, (n, 3)),
                                              None if g_ws is None else g_ws.float().contiguous(),
                                              sigmas, rgbs, deltas, ts, rays_a, ctx.T_threshold)
        return dsig, drgbs, None, None, None, None


class VolumeRenderer(torch.nn.Module):
    """forward(sigmas, rgbs, deltas, ts, rays_a, T_threshold) ->
    (vr_samples, opacity, depth, rgb, ws)   (reference: VolumeRenderer, volume_train.py:52-195)."""

    def forward(self, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        total, opacity, depth, rgb, ws = _Composite.apply(
            sigmas.float().contiguous(), rgbs.contiguous(), deltas.contiguous(), ts.contiguous(),
            rays_a.contiguous(), T_threshold)
        return total.sum(), opacity, depth, rgb, ws
