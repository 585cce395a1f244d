"""Degree-4 spherical-harmonics direction encoder — mirrors modules/spherical_harmonics.py."""
import torch

from taichi_nerfs_b200 import ops


class _DirEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs):
        return ops.dir_encode(dirs)

    @staticmethod
    def backward(ctx, grad_out):
        # view directions never require grad on the NGP path; like the reference
        # (spherical_harmonics.py:88-98) nothing upstream consumes this.
        return None


class DirEncoder(torch.nn.Module):
    """forward(dirs [N,3]) -> [N,16] fp32 (reference: dir_encoder kernel :7-42)."""

    def __init__(self):
        super().__init__()
        self.out_dim = 16

    def forward(self, dirs):
        return _DirEncode.apply(dirs.contiguous())
