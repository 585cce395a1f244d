"""Training step of the NGP hot path (the body of the reference's loop, train.py:168-201):
get_rays -> render -> MSE -> backward -> [gradient all-reduce] -> optimizer step.

``NGPTrainer.step`` keeps the reference's numerics (torch.autocast(fp16), loss scaling as
GradScaler(2**16 | 2**19), Adam(eps=1e-15), cosine LR to lr/30) but replaces
optimizer.zero_grad + GradScaler.unscale_/inf-check + Adam + the fp16 table re-cast by ONE fused pass
per parameter (csrc/optim.cu) and never synchronises the host for the inf check.

Multi-GPU: rays are sharded across ranks (each rank renders its own batch); the only collective is
one all-reduce (sum) of the flat gradient buffer per step, folded into the fused Adam as inv_scale /
world_size.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ops, parallel


class NGPTrainer:
    def __init__(self, model, lr: float = 1e-2, max_steps: int = 20000, loss_scale: float | None = None,
                 betas=(0.9, 0.999), eps: float = 1e-15, process_group=None):
        self.model = model
        self.lr0 = lr
        self.max_steps = max_steps
        self.betas = betas
        self.eps = eps
        # train.py:137-141: GradScaler(2**16) with --half_opt, else 2**19
        self.loss_scale = float(loss_scale if loss_scale is not None else (2 ** 16 if model.half_opt else 2 ** 19))
        self.step_count = 0
        self.pg = process_group
        self.world_size = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world_size = torch.distributed.get_world_size(process_group)

        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        # one flat gradient buffer so the multi-GPU path is a single all-reduce (hash grad | MLP grads)
        sizes = [p.numel() for p in self.params]
        pad = [(-s) % 4 for s in sizes]  # keep every slice 16-byte aligned for the float4 Adam kernel
        total = sum(s + q for s, q in zip(sizes, pad))
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        # parameters live in one flat buffer as well (each nn.Parameter becomes a view of it, state_dict is
        # unchanged), so the fused Adam is ONE launch over [hash table | MLP weights]
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.slices = []
        off = 0
        for p, s, q in zip(self.params, sizes, pad):
            self.slices.append((off, s))
            p.grad = self.flat_grad[off:off + s].view_as(p)  # autograd accumulates straight into the flat buffer
            with torch.no_grad():
                self.flat_param[off:off + s].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + s].view_as(p)
            off += s + q
            if p is getattr(model.pos_encoder, 'hash_table', None):
                model.pos_encoder.grad_sink = self.flat_grad[off - s - q:off - q]
        self.found_inf = torch.zeros(1, device=dev, dtype=torch.int32)
        self._shadow = self._shadow_full = None
        enc = model.pos_encoder
        if hasattr(enc, "adopt_shadow"):
            assert self.params[0] is enc.hash_table, "the hash table must be the first parameter"
            # fp16 copy of the whole flat buffer, rewritten by the Adam pass; the encoder reads its first slice
            self._shadow_full = self.flat_param.to(torch.float16)
            self._shadow = self._shadow_full[:enc.hash_table.numel()]
            enc.adopt_shadow(self._shadow)

    # cosine annealing to lr/30 (train.py:159-163, CosineAnnealingLR(T_max=max_steps, eta_min=lr/30))
    def lr_at(self, step: int) -> float:
        eta_min = self.lr0 / 30
        return eta_min + (self.lr0 - eta_min) * (1 + math.cos(math.pi * min(step, self.max_steps) / self.max_steps)) / 2

    def forward_backward(self, rays_o, rays_d, rgb_gt, exp_step_factor=0.0, extra_loss=None):
        from modules.rendering import render
        with torch.autocast(device_type='cuda', dtype=torch.float16):
            results = render(self.model, rays_o, rays_d, exp_step_factor=exp_step_factor)
            loss = F.mse_loss(results['rgb'], rgb_gt)
            if extra_loss is not None:  # e.g. distortion loss (train.py:194-195)
                loss = loss + extra_loss(results)
        (loss * self.loss_scale).backward()
        return loss, results

    def optimizer_step(self):
        self.step_count += 1
        parallel.allreduce_gradients(self.flat_grad, self.pg)
        self.found_inf.zero_()
        ops.check_finite(self.flat_grad, self.found_inf)  # after the sum: identical on every rank
        lr = self.lr_at(self.step_count - 1)
        inv = parallel.inv_grad_scale(self.loss_scale, self.world_size)
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, lr, self.step_count,
                      self.betas[0], self.betas[1], self.eps, inv, param_f16=self._shadow_full,
                      found_inf=self.found_inf, zero_grad=True)
        if self._shadow is not None:
            self.model.pos_encoder.adopt_shadow(self._shadow)

    def step(self, rays_o, rays_d, rgb_gt, exp_step_factor=0.0, extra_loss=None):
        loss, results = self.forward_backward(rays_o, rays_d, rgb_gt, exp_step_factor, extra_loss)
        self.optimizer_step()
        return loss, results
