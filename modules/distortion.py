"""Mip-NeRF-360 distortion loss — mirrors modules/distortion.py of the reference
(distortion_loss :8-12, DistortionLoss :122-194); off by default (opt.py --distortion_loss_w 0)."""
import torch

from taichi_nerfs_b200 import ops


class DistortionLoss(torch.autograd.Function):
    """forward(ws [S], deltas [S], ts [S], rays_a [N,3]) -> loss [N] (indexed by ray id)."""

    @staticmethod
    def forward(ctx, ws, interval, tmid, packed_info):
        ws, interval, tmid = ws.float().contiguous(), interval.contiguous(), tmid.contiguous()
        packed_info = packed_info.contiguous()
        loss = ops.distortion_fwd(ws, interval, tmid, packed_info)
        ctx.save_for_backward(ws, interval, tmid, packed_info)
        return loss

    @staticmethod
    def backward(ctx, dL_dloss):
        ws, interval, tmid, packed_info = ctx.saved_tensors
        return ops.distortion_bwd(dL_dloss.float().contiguous(), ws, interval, tmid, packed_info), None, None, None


def distortion_loss(results):
    return DistortionLoss.apply(results['ws'], results['deltas'], results['ts'], results['rays_a'])
