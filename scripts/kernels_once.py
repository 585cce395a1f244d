"""Launches each major kernel of the training step a few times on the tensors of one real step
(for ncu captures: `ncu --set full -k regex:<name> ... python scripts/kernels_once.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from datasets.ray_utils import get_rays
from datasets.synthetic import SyntheticLego
from modules.intersection import ray_aabb_intersection
from modules.networks import NGP
from modules.ray_march import raymarching_train
from taichi_nerfs_b200 import ops
from taichi_nerfs_b200.fused_mlp import mlp_weights

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
lay, table, ws = bench.init_weights_numpy(bench.SEED)
model = NGP(scale=0.5, max_res=1024, half_opt=True).to(dev)
with torch.no_grad():
    model.pos_encoder.hash_table.copy_(torch.from_numpy(table))
ds = SyntheticLego(batch_size=bench.BATCH).to(dev)
model.mark_invisible_cells(ds.K, ds.poses, ds.img_wh)
with torch.autocast("cuda", dtype=torch.float16):
    model.update_density_grid(bench.DENSITY_THRESHOLD, warmup=True)
b = ds[0]
rays_o, rays_d = get_rays(b["direction"], b["pose"])
enc = model.pos_encoder
W = [w.detach() for w in mlp_weights(model)]
for _ in range(reps):
    hits = ray_aabb_intersection(rays_o, rays_d, model.scale)
    rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(rays_o, rays_d, hits, model.density_bitfield, 1, 0.5, 0.0, 128, 1024)
    xn = ((xyzs - model.xyz_min) / (model.xyz_max - model.xyz_min)).contiguous()
    emb = ops.hash_encode_fwd(xn, enc.table_f16(), enc._clayout, 32)
    sig, rgbs = ops.mlp_fwd(emb, dirs, W)
    tot, op, dep, rgb, wsamp = ops.composite_train_fwd(sig, rgbs, deltas, ts, rays_a, 1e-4)
    n = rays_a.shape[0]
    g_rgb = (torch.randn(n, 3, device=dev) * 8).contiguous()
    dsig, drgbs = ops.composite_train_bwd(-g_rgb.sum(1).contiguous(), torch.zeros(n, device=dev), g_rgb, None, sig, rgbs, deltas, ts, rays_a, 1e-4)
    demb, gw = ops.mlp_bwd(emb, dirs, W, dsig, drgbs)
    grad = torch.zeros(enc.total_param_size, device=dev)
    ops.hash_encode_bwd(xn, demb, enc._clayout, grad)
torch.cuda.synchronize()
print("samples", int(total))
