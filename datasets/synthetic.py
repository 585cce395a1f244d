"""Synthetic "Lego-shape" dataset: the geometry of the reference's Synthetic-NeRF/NSVF Lego setup
(800x800 pinhole, fx=fy=1111.111 — datasets/nsvf.py:37-44; cameras on the upper hemisphere looking at
the origin, scene in [-0.5,0.5]^3) with random target colours.  No image files are available offline,
so this stands in for datasets/{nsvf,nerf}.py on the benchmark path.

Follows the BaseDataset protocol of the reference (datasets/base.py:6-61): ``dataset[i]`` in a train
split returns {'img_idxs','pix_idxs','pose','direction','rgb'} for ``batch_size`` random rays.
"""
from __future__ import annotations

import math

import torch

from .ray_utils import get_ray_directions


def hemisphere_poses(n_poses: int, radius: float = 1.4, seed: int = 23) -> torch.Tensor:
    """(n,3,4) camera-to-world matrices, columns = [right, down, front, position]."""
    g = torch.Generator().manual_seed(seed)
    theta = torch.rand(n_poses, generator=g) * 2 * math.pi
    phi = torch.acos(torch.rand(n_poses, generator=g) * 0.9 + 0.05)  # elevation away from the horizon/pole
    pos = torch.stack([torch.sin(phi) * torch.cos(theta), torch.sin(phi) * torch.sin(theta), torch.cos(phi)], -1) * radius
    front = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(front)
    right = torch.cross(front, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(front, right, dim=-1)
    return torch.stack([right, down, front, pos], dim=-1).float()


class SyntheticLego:
    def __init__(self, n_images: int = 100, img_wh=(800, 800), focal: float = 1111.111, radius: float = 1.4,
                 split: str = 'train', batch_size: int = 8192, seed: int = 23, with_rgb: bool = True):
        w, h = img_wh
        self.img_wh = (w, h)
        self.split = split
        self.batch_size = batch_size
        self.ray_sampling_strategy = 'all_images'
        self.K = torch.tensor([[focal, 0, w / 2], [0, focal, h / 2], [0, 0, 1]], dtype=torch.float32)
        self.directions = get_ray_directions(h, w, self.K)
        self.poses = hemisphere_poses(n_images, radius, seed)
        self._gen = None
        self._seed = seed
        self.with_rgb = with_rgb

    def __len__(self):
        return len(self.poses)

    def to(self, device):
        self.K = self.K.to(device)
        self.directions = self.directions.to(device)
        self.poses = self.poses.to(device)
        return self

    def __getitem__(self, idx):
        dev = self.poses.device
        if self._gen is None or self._gen.device != dev:
            self._gen = torch.Generator(device=dev).manual_seed(self._seed)
        n = self.batch_size
        if self.split.startswith('train'):
            img = torch.randint(0, len(self.poses), (n,), device=dev, generator=self._gen)
            pix = torch.randint(0, self.img_wh[0] * self.img_wh[1], (n,), device=dev, generator=self._gen)
            sample = {'img_idxs': img, 'pix_idxs': pix, 'pose': self.poses[img], 'direction': self.directions[pix]}
            if self.with_rgb:
                sample['rgb'] = torch.rand(n, 3, device=dev, generator=self._gen)
            return sample
        return {'pose': self.poses[idx], 'img_idxs': idx}
