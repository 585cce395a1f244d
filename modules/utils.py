"""Constants and occupancy-grid helpers — mirrors the public names of the reference's
modules/utils.py (constants :12-16, layout helpers :19-42, morton3D[_invert] :120-154,
packbits :157-169, save_deployment_model :230-253) plus the deployment containers
(.bin / deployment.npy readers and writers of deployment/InstantNGP/)."""
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


# ---- deployment containers ---------------------------------------------------
_AOT_DTYPES = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.int16, 4: np.uint32, 5: np.uint16}


def write_aot_array(folder, arr, name):
    """``<name>.bin`` = [int32 dtype code][int32 numel][flat payload] — the container the mobile demos read
    (deployment/InstantNGP/taichi_ngp/taichi_ngp.py:34-65 writes it, deployment/InstantNGP/utils/utils.cpp:100-175
    parses it).  Codes: 0 f32, 1 f16, 2 i32, 3 i16, 4 u32, 5 u16."""
    arr = np.ascontiguousarray(arr)
    codes = {np.dtype(v): k for k, v in _AOT_DTYPES.items()}
    if arr.dtype not in codes:
        raise TypeError(f"dtype {arr.dtype} has no code in the .bin container")
    if arr.size >= 2 ** 31:
        raise ValueError("the .bin header stores numel as int32")
    path = os.path.join(folder, name + '.bin')
    with open(path, 'wb') as f:
        f.write(np.array([codes[arr.dtype], arr.size], dtype='<i4').tobytes())
        f.write(arr.reshape(-1).tobytes())
    return path


def read_aot_array(path):
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 8:
        raise ValueError(f"{path}: truncated header")
    code, numel = (int(v) for v in raw[:8].view('<i4'))
    if code not in _AOT_DTYPES:
        raise ValueError(f"{path}: invalid buffer dtype code {code}")   # utils.cpp:151-154
    dt = np.dtype(_AOT_DTYPES[code])
    if raw.size != 8 + numel * dt.itemsize:
        raise ValueError(f"{path}: invalid buffer size ({raw.size} bytes for {numel} x {dt})")  # utils.cpp:159-160
    return raw[8:].view(dt)


def export_aot_weights(blob, folder, directions=None, pose_index=20):
    """The six weight files of the mobile demo from a ``deployment.npy`` dict (taichi_ngp.py:66-86)."""
    os.makedirs(folder, exist_ok=True)
    f32 = lambda a: np.asarray(a).astype(np.float32)  # noqa: E731
    write_aot_array(folder, f32(blob['model.hash_encoder.params']), 'hash_embedding')
    write_aot_array(folder, f32(blob['model.xyz_encoder.params']), 'sigma_weights')
    write_aot_array(folder, f32(blob['model.rgb_net.params']), 'rgb_weights')
    write_aot_array(folder, np.ascontiguousarray(blob['model.density_bitfield']).view(np.uint32), 'density_bitfield')
    poses = np.asarray(blob['poses'])
    write_aot_array(folder, f32(poses[min(pose_index, len(poses) - 1)]).reshape(3, 4), 'pose')
    if directions is not None or 'model.directions' in blob:
        write_aot_array(folder, f32(directions if directions is not None else blob['model.directions']), 'directions')


def load_deployment_model(model, source):
    """Inverse of save_deployment_model: fill an ``NGP`` built with the deployment config (train.py:88-99) from a
    ``deployment.npy`` file / dict or from a folder of ``.bin`` files.  Returns the extra entries (poses, pose,
    directions) that are not model state."""
    if isinstance(source, dict):
        blob = source
    elif os.path.isdir(source):
        names = {'hash_embedding': 'model.hash_encoder.params', 'sigma_weights': 'model.xyz_encoder.params',
                 'rgb_weights': 'model.rgb_net.params', 'density_bitfield': 'model.density_bitfield',
                 'pose': 'pose', 'directions': 'model.directions'}
        blob = {key: read_aot_array(os.path.join(source, n + '.bin')) for n, key in names.items()
                if os.path.exists(os.path.join(source, n + '.bin'))}
    else:
        blob = np.load(source, allow_pickle=True).item()
    xyz, rgb = model.xyz_encoder, model.rgb_net
    if len(xyz.hidden_layers) != 1 or len(rgb.hidden_layers) != 1:
        raise ValueError("deployment files hold one hidden layer per network (train.py:88-99 config)")

    def fill(param, flat, what):
        flat = np.asarray(flat, dtype=np.float32).reshape(-1)
        if flat.size != param.numel():
            raise ValueError(f"{what}: {flat.size} values for a parameter of shape {tuple(param.shape)}")
        with torch.no_grad():
            param.copy_(torch.from_numpy(flat.copy()).view_as(param))

    fill(model.pos_encoder.hash_table, blob['model.hash_encoder.params'], 'hash table')
    sw = np.asarray(blob['model.xyz_encoder.params'], dtype=np.float32).reshape(-1)
    n1 = xyz.hidden_layers[0].weight.numel()
    fill(xyz.hidden_layers[0].weight, sw[:n1], 'sigma net layer 1')
    fill(xyz.output_layer.weight, sw[n1:], 'sigma net layer 2')
    rw = np.asarray(blob['model.rgb_net.params'], dtype=np.float32).reshape(-1)
    n1 = rgb.hidden_layers[0].weight.numel()
    fill(rgb.hidden_layers[0].weight, rw[:n1], 'rgb net layer 1')
    out_w = rgb.output_layer.weight
    fill(out_w, rw[n1:].reshape(-1, out_w.shape[1])[:out_w.shape[0]], 'rgb net output')  # rows 3..15 are padding
    bits = np.ascontiguousarray(blob['model.density_bitfield']).view(np.uint8).reshape(-1)
    if bits.size != model.density_bitfield.numel():
        raise ValueError(f"density bitfield: {bits.size} bytes, model expects {model.density_bitfield.numel()}")
    with torch.no_grad():
        model.density_bitfield.copy_(torch.from_numpy(bits.copy()))
    return {k: blob[k] for k in ('poses', 'pose', 'model.directions') if k in blob}
