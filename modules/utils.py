"""Constants and occupancy-grid helpers — mirrors the public names of the reference's
modules/utils.py (constants :12-16, layout helpers :19-42, morton3D[_invert] :120-154,
packbits :157-169, save_deployment_model :230-253)."""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from taichi_nerfs_b200 import ops

torch_type = torch.float32

MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.01
SQRT3 = 1.7320508075688772
SQRT3_MAX_SAMPLES = SQRT3 / 1024
SQRT3_2 = 1.7320508075688772 * 2


def res_in_level_np(level_i, base_res, log_per_level_scale):
    """Resolution of hash level ``level_i`` (f64, as used for table sizing)."""
    return float(math.ceil(float(base_res) * math.exp(float(level_i) * log_per_level_scale) - 1.0) + 1)


def scale_in_level_np(base_res, max_res, levels):
    """log of the per-level growth factor b."""
    return math.log(float(max_res) / float(base_res)) / float(levels - 1)


def align_to(x, y):
    return int((x + y - 1) / y) * y


def morton3D(coords1):
    """(N,3) int32 grid coordinates -> (N,) int32 Morton codes.  No host sync (the reference
    calls ti.sync(), utils.py:153)."""
    return ops.morton3d(coords1.to(torch.int32))


def morton3D_invert(indices):
    """(N,) int32 Morton codes -> (N,3) int32 grid coordinates."""
    return ops.morton3d_invert(indices.to(torch.int32))


def packbits(density_grid, density_threshold, density_bitfield):
    """density_bitfield[n] bit i = density_grid[8n+i] > density_threshold (in place)."""
    ops.packbits(density_grid, density_threshold, density_bitfield)


def depth2img(depth):
    """Turbo-ish colour map of a depth image without the cv2 dependency of the reference."""
    d = (depth - depth.min()) / max(float(depth.max() - depth.min()), 1e-12)
    x = np.clip(d, 0.0, 1.0)[..., None]
    r = np.clip(1.5 - np.abs(4.0 * x - 3.0), 0, 1)
    g = np.clip(1.5 - np.abs(4.0 * x - 2.0), 0, 1)
    b = np.clip(1.5 - np.abs(4.0 * x - 1.0), 0, 1)
    return (np.concatenate([b, g, r], -1) * 255).astype(np.uint8)


def save_deployment_model(model, dataset, save_dir):
    """Write ``deployment.npy`` in the layout the reference's mobile demo loads
    (modules/utils.py:230-253 <-> deployment/InstantNGP/taichi_ngp/kernels.py:385-518)."""
    w_out = model.rgb_net.output_layer.weight.detach().cpu()
    w_out = torch.cat([w_out, torch.zeros(13, w_out.shape[1])], dim=0)  # pad 3 -> 16 rows
    blob = {
        'poses': dataset.poses.cpu().numpy(),
        'model.density_bitfield': model.density_bitfield.cpu().numpy(),
        'model.hash_encoder.params': model.pos_encoder.hash_table.detach().cpu().numpy(),
        'model.per_level_scale': model.pos_encoder.log_b,
        'model.xyz_encoder.params': torch.cat([
            model.xyz_encoder.hidden_layers[0].weight.detach().cpu().reshape(-1),
            model.xyz_encoder.output_layer.weight.detach().cpu().reshape(-1)]).numpy(),
        'model.rgb_net.params': torch.cat([
            model.rgb_net.hidden_layers[0].weight.detach().cpu().reshape(-1),
            w_out.reshape(-1)]).numpy(),
    }
    np.save(os.path.join(f'{save_dir}', 'deployment.npy'), blob)
