#!/usr/bin/env python
"""bench.py — rays/s of the Instant-NGP training hot path on B200 (driver contract: see task statement).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference restatement on host cores

A "step" = one pass of the hot path over one batch of synthetic Lego-shape rays:
get_rays -> ray/AABB -> occupancy march -> hash encode -> MLP(+SH) -> composite -> MSE -> backward ->
[grad all-reduce] -> fused Adam, plus update_density_grid every 16th step exactly like the reference's
loop (train.py:168-201).  Workload = BASELINE.json configs[1]: Lego shape, 8192 rays/GPU, fp16 hash
encoder, random-init table + MLP, occupancy from one warm-up grid update ("occupancy B").
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 8192
SEED = 23
UPDATE_INTERVAL = 16          # train.py:57-58
PREWARM = 20                  # extra untimed steps before the W warm-up steps
DENSITY_THRESHOLD = 0.01 * 1024 / 3 ** 0.5  # train.py:180
# algorithmic bytes per sample, fp16 encoder
BYTES_PER_SAMPLE = {"hash_fwd": 588, "hash_bwd": 1100, "mlp_fwd": 86, "mlp_bwd": 150,
                    "composite_fwd": 22, "composite_bwd": 32, "march": 32}


# DRAM bytes per launch of each kernel from the committed `ncu --set full` capture (profiles/), same workload
NCU_DRAM_BYTES_PER_LAUNCH = {"hash_bwd": 215.3e6, "hash_fwd": 146.0e6, "mlp_bwd": 288.2e6, "mlp_fwd": 181.2e6}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def init_weights_numpy(seed):
    """Random-init parameters shared by both arms: table U(-1e-4,1e-4) (hash_encoder_half.py:299),
    xavier-uniform MLP (networks.py:306-312)."""
    from taichi_nerfs_b200.layout import make_hash_layout
    rng = np.random.default_rng(seed)
    lay = make_hash_layout(2 ** 19, 16, 16, 1024, 2)
    table = ((rng.random((lay.total_entries, 2), dtype=np.float32) * 2 - 1) * 1e-4).astype(np.float32)
    shapes = [(64, 32), (16, 64), (64, 32), (64, 64), (3, 64)]
    ws = [(rng.uniform(-1, 1, s) * math.sqrt(6.0 / (s[0] + s[1]))).astype(np.float32) for s in shapes]
    return lay, table, ws


def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    return world, rank, local


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.nn.functional as F
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    from datasets.ray_utils import get_rays
    from datasets.synthetic import SyntheticLego
    from modules.networks import NGP
    from taichi_nerfs_b200 import _lib, ops
    from taichi_nerfs_b200.trainer import NGPTrainer

    torch.manual_seed(SEED + rank)
    lay, table, ws = init_weights_numpy(SEED)
    model = NGP(scale=0.5, max_res=1024, half_opt=True).to(dev)
    with torch.no_grad():
        model.pos_encoder.hash_table.copy_(torch.from_numpy(table))
        for p, w in zip([model.xyz_encoder.hidden_layers[0].weight, model.xyz_encoder.output_layer.weight,
                         model.rgb_net.hidden_layers[0].weight, model.rgb_net.hidden_layers[1].weight,
                         model.rgb_net.output_layer.weight], ws):
            p.copy_(torch.from_numpy(w))
    ds = SyntheticLego(batch_size=BATCH, seed=SEED + rank).to(dev)
    model.mark_invisible_cells(ds.K, ds.poses, ds.img_wh)
    with torch.autocast("cuda", dtype=torch.float16):
        model.update_density_grid(DENSITY_THRESHOLD, warmup=True)
    occupied = float(np.unpackbits(model.density_bitfield.cpu().numpy()).mean())
    trainer = NGPTrainer(model, lr=1e-2, max_steps=20000)

    n_total = args.steps + args.warmup
    batches = [ds[0] for _ in range(n_total)]                      # device-resident inputs
    host_batches = [{k: v.cpu().pin_memory() for k, v in b.items() if k in ("direction", "pose", "rgb")}
                    for b in batches]

    sample_counts = []
    from taichi_nerfs_b200.fast_step import StaticTrainStep
    fast = None if args.path == "modules" else StaticTrainStep(trainer, BATCH, samples_per_ray_capacity=384,
                                                               overlap_optimizer=not args.no_overlap)
    if fast is not None:
        # the training set stays resident in HBM (train.py: `train_dataset.to(device)`), and the step draws its own
        # batch on the device (datasets/base.py:34-61 + get_rays as the first node of the graph)
        ds.build_image_bank()
        fast.attach_ray_source(ds.rays, ds.poses, ds.directions, seed=SEED + rank)

    def one_step(step_idx, b):
        with torch.autocast("cuda", dtype=torch.float16):
            if step_idx % UPDATE_INTERVAL == 0:
                if fast is not None:
                    fast.flush()   # overlap mode: the grid update must see the parameters of the last step
                model.update_density_grid(DENSITY_THRESHOLD, warmup=step_idx < 256)
        if fast is not None and b is None:   # batch sampling + the whole step = one CUDA-graph replay, no host sync
            loss = fast.step_sampled()
            sample_counts.append(fast.counter[0].clone())
            return loss
        rays_o, rays_d = get_rays(b["direction"], b["pose"])
        if fast is not None:   # caller-provided batch (host buffers in the e2e arm)
            loss = fast.step(rays_o, rays_d, b["rgb"])
            sample_counts.append(fast.counter[0].clone())
        else:                  # reference-shaped module API: render() + autograd + fused Adam
            loss, results = trainer.step(rays_o, rays_d, b["rgb"])
            sample_counts.append(results["rm_samples"])
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms)

    if fast is not None:
        batches = [None] * n_total   # "value" arm: batches are drawn on the device inside the graph
    if args.ncu_window > 0:
        for s in range(PREWARM):
            one_step(1 + s % 8, batches[s % len(batches)])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for s in range(args.ncu_window):
            one_step(1 + s, batches[s % len(batches)])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"ncu_window_steps": args.ncu_window,
                              "samples_per_step": [int(c) for c in sample_counts[-args.ncu_window:]]}))
        return

    # ---- device-resident arm ("value") -------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # untimed pre-warm beyond --warmup: the caching allocator must have seen the range of per-step
    # sample counts (every new size is a cudaMalloc) and the clocks must have ramped up
    for s in range(PREWARM):
        one_step(1 + s % 8, batches[s % len(batches)])
    for s in range(args.warmup):
        one_step(s, batches[s])
    launches0 = _lib.launch_count()
    def timed_steps():
        for k in range(args.steps):
            one_step(args.warmup + k, batches[args.warmup + k])
        if fast is not None:
            fast.flush()   # every one of the K updates is applied inside the timed region
    graph0 = fast.graph_kernel_launches if fast is not None else 0
    ms_total = timed(timed_steps)
    clock_info = clocks.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0   # eager launches of libngp_b200 kernels
    if fast is not None:                          # + kernel nodes executed by CUDA-graph replays
        launches += fast.graph_kernel_launches - graph0
    ms_step = ms_total / args.steps
    value = world * BATCH / (ms_step * 1e-3)
    spr = float(torch.stack([c.float() for c in sample_counts[-args.steps:]]).mean()) / BATCH

    # ---- end-to-end arm: host buffers, H2D inputs + D2H loss every step ----------------------------------
    def e2e_step(step_idx, hb):
        b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        loss = one_step(step_idx, b)
        return float(loss.detach().float().cpu())   # device->host read of the step's result

    base = args.warmup + args.steps
    for s in range(min(3, args.warmup)):
        e2e_step(base + s, host_batches[s])
    def timed_e2e():
        for k in range(args.steps):
            e2e_step(base + 3 + k, host_batches[args.warmup + k])
        if fast is not None:
            fast.flush()
    ms_e2e = timed(timed_e2e)
    e2e_value = world * BATCH / (ms_e2e / args.steps * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())

    # ---- roofline of the dominant kernel, timed live with CUDA events --------------------------------------
    roof = kernel_roofline(torch, ops, model, trainer, ds, get_rays, dev)

    # ---- amortised density-grid update --------------------------------------------------------------------
    def upd():
        with torch.autocast("cuda", dtype=torch.float16):
            model.update_density_grid(DENSITY_THRESHOLD, warmup=True)
    upd()
    upd_ms = timed(upd)

    if rank != 0:
        return
    line = {
        "metric": "rays/sec (8192-ray batch, Lego shape)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: Synthetic-NeRF Lego shape, batch 8192 rays/GPU, fp16 hash encoder "
                               "L=16 T=2^19 F=2, random-init table+MLP, occupancy B (one warm-up grid update)",
                   "rays_per_gpu": BATCH, "global_batch": world * BATCH, "samples_per_ray": spr,
                   "occupied_fraction": occupied, "parallelism": f"ray-sharded dp{world}, 1 NCCL all-reduce/step",
                   "l2": "no flush: per-step working set (~%d MB of per-sample tensors) exceeds the 126 MB L2; "
                         "new rays every step" % int(spr * BATCH * 2010 / 1e6),
                   "density_grid_update": f"inside timed loop every {UPDATE_INTERVAL} steps (warm-up mode); "
                                          f"{upd_ms:.3f} ms each",
                   "mlp": "torch.nn.Linear (cuBLAS) under autocast" if not _fused_mlp() else "fused tcgen05 kernel",
                   "step_path": "StaticTrainStep: batch sampling (resident 100x800x800 training set) + whole step = one "
                                "CUDA-graph replay, sample count stays on the device; e2e arm: host batches -> get_rays -> "
                                "the same graph without the sampler node"
                                + ("" if args.no_overlap else "; optimizer of step k runs on a parallel graph branch "
                                   "beside ray_aabb + marching of step k+1 (flushed before every grid update and at "
                                   "the end of the timed region)")
                                if fast is not None else "modules API: render() + torch.autograd + fused Adam"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": roof,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(budget_s=args.cpu_budget)
    print(json.dumps(line))


def _fused_mlp():
    try:
        from taichi_nerfs_b200 import fused_mlp
        return fused_mlp.available()
    except ImportError:
        return False


def kernel_roofline(torch, ops, model, trainer, ds, get_rays, dev):
    """Times each of this library's major kernels alone (CUDA events on the launching stream, L2 flushed
    between repeats) on the tensors of one real step and reports the roofline of the slowest one."""
    from modules.intersection import ray_aabb_intersection
    from modules.ray_march import raymarching_train
    peak, peak_src = measured_peaks()
    b = ds[0]
    rays_o, rays_d = get_rays(b["direction"], b["pose"])
    hits = ray_aabb_intersection(rays_o, rays_d, model.scale)
    rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(rays_o, rays_d, hits, model.density_bitfield,
                                                              model.cascades, model.scale, 0.0, model.grid_size, 1024)
    S = int(total)
    enc = model.pos_encoder
    xn = ((xyzs - model.xyz_min) / (model.xyz_max - model.xyz_min)).contiguous()
    table = enc.table_f16()
    emb = ops.hash_encode_fwd(xn, table, enc._clayout, enc.out_dim)
    dout = (torch.randn_like(emb.float()) * 1e-3).half()
    grad = torch.zeros(enc.total_param_size, device=dev)
    ws = [model.xyz_encoder.hidden_layers[0].weight, model.xyz_encoder.output_layer.weight,
          model.rgb_net.hidden_layers[0].weight, model.rgb_net.hidden_layers[1].weight, model.rgb_net.output_layer.weight]
    sig, rgbs = ops.mlp_fwd(emb, dirs, ws)
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)

    def t(fn, reps=5):
        out = []
        for _ in range(reps):
            flush.fill_(1)                      # evict L2 (126 MB) between repeats
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return statistics.median(out)

    times = {
        "hash_fwd": t(lambda: ops.hash_encode_fwd(xn, table, enc._clayout, enc.out_dim)),
        "hash_bwd": t(lambda: ops.hash_encode_bwd(xn, dout, enc._clayout, grad)),
        "mlp_fwd": t(lambda: ops.mlp_fwd(emb, dirs, ws)),
        "composite_fwd": t(lambda: ops.composite_train_fwd(sig, rgbs, deltas, ts, rays_a, 1e-4)),
    }
    top = max(times, key=times.get)
    alg_bytes = BYTES_PER_SAMPLE[top] * S
    achieved = alg_bytes / (times[top] * 1e-3) / 1e9
    return {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(top), "traffic_unit": "bytes/launch",
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, "
                              "profiles/r1_kernels_ncu_table_final.md (one graph step, S=2.19 M samples)",
            "peak_source": peak_src, "samples": S,
            "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE[top], "kernel_ms": times,
            "note": "fp16 table (21.8 MiB) + fp32 grad (43.6 MiB) fit the 126 MB L2: gathers/atomics are L2-bound, "
                    "so algorithmic GB/s over the HBM peak can exceed 1 (BASELINE.md §5)"}


# --------------------------------------------------------------------------------------------------
def oracle_workload(n_rays, seed):
    """Builds the CPU-arm model (same random init, occupancy B computed by the oracle) and ray batches."""
    from oracle import oracle as O
    from oracle import train_step as TS
    O.build()
    lay, table, ws = init_weights_numpy(SEED)
    rng = np.random.default_rng(seed)
    # occupancy B: density at a jittered point of every cell, threshold = min(mean, 5.91) (networks.py:255-290)
    g = 128
    coords = np.stack(np.meshgrid(np.arange(g), np.arange(g), np.arange(g), indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
    s, half = 0.5, 0.5 / g
    xyz = (coords / (g - 1) * 2 - 1) * (s - half) + (rng.random((g ** 3, 3)) * 2 - 1) * half
    xn = ((xyz + 0.5) / 1.0).astype(np.float32)
    emb = O.hash_encode_fwd(xn, table.astype(np.float16).reshape(-1), lay)
    dens, _ = O.mlp_fwd(emb, np.tile(np.array([[0, 0, 1]], np.float32), (g ** 3, 1)), ws)
    grid = np.zeros(g ** 3, np.float32)
    grid[O.morton3d(coords).astype(np.int64)] = dens
    thr = min(float(grid[grid > 0].mean()), DENSITY_THRESHOLD)
    bitfield = O.packbits(grid, thr)
    model = TS.OracleModel(lay, table, ws, bitfield, scale=0.5, cascades=1, half=True)

    make_rays = TS.make_rays
    def batch(i):
        o, d = make_rays(n_rays, seed=seed * 1000 + i)
        r = np.random.default_rng(seed * 1000 + i)
        return o, d, r.random((n_rays, 3), dtype=np.float32), r.random(n_rays, dtype=np.float32)
    return TS, model, batch


def cpu_baseline(budget_s=20.0, n_rays=2048):
    TS, model, batch = oracle_workload(n_rays, SEED)
    cores = len(os.sched_getaffinity(0))
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    o, d, gt, nz = batch(0)
    TS.train_step(model, o, d, gt, nz)      # warm-up
    t0, steps, samples = time.perf_counter(), 0, 0
    while time.perf_counter() - t0 < budget_s and steps < 64:
        o, d, gt, nz = batch(steps + 1)
        _, cache = TS.train_step(model, o, d, gt, nz)
        samples += cache["S"]
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": steps * n_rays / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{steps} full train steps of {n_rays} rays (same Lego-shape workload, occupancy B, "
                      f"{samples / max(steps * n_rays, 1):.0f} samples/ray) through oracle/ngp_oracle.c (OpenMP, {cores} threads)"}


def run_reference(args):
    """Reference arm: the reference's algorithm on the host cores (the Taichi reference itself cannot be
    installed offline — see DESIGN.md — so this is the strict-fp32 C/OpenMP restatement, kind=port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    n_rays = args.ref_rays
    if n_rays <= 0:
        # as much of the 8192-ray batch per step as fits a ~2 minute run on this host (per-step fixed costs -
        # Adam over 11.4 M parameters, gradient zeroing - are amortised as in the real workload)
        TS, model, batch = oracle_workload(256, SEED)
        o, d, gt, nz = batch(0)
        TS.train_step(model, o, d, gt, nz)
        t0 = time.perf_counter()
        TS.train_step(model, *batch(1))
        per_ray = (time.perf_counter() - t0) / 256
        n_rays = 256
        while n_rays < BATCH and (args.steps + args.warmup) * (2 * n_rays) * per_ray <= 120.0:
            n_rays *= 2
    TS, model, batch = oracle_workload(n_rays, SEED)
    for s in range(args.warmup):
        o, d, gt, nz = batch(s)
        TS.train_step(model, o, d, gt, nz)
    t0 = time.perf_counter()
    samples = 0
    for k in range(args.steps):
        o, d, gt, nz = batch(args.warmup + k)
        _, cache = TS.train_step(model, o, d, gt, nz)
        samples += cache["S"]
    dt = time.perf_counter() - t0
    value = args.steps * n_rays / dt
    sample = (f"{args.steps} steps x {n_rays} rays per step (bounded sample of the 8192-ray workload, "
              f"{samples / (args.steps * n_rays):.0f} samples/ray)")
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (8192-ray batch, Lego shape)", "value": value, "unit": "rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] restated on CPU: Lego shape, fp16 hash encoder semantics, "
                               "random-init table+MLP, occupancy B; " + sample},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-rays", type=int, default=0,
                    help="rays per step of the reference arm's bounded sample (0 = as many of the 8192 as fit ~2 min)")
    ap.add_argument("--path", default="graph", choices=["graph", "modules"],
                    help="graph: StaticTrainStep (one CUDA graph per step, sync-free); "
                         "modules: render()+autograd through the reference-shaped module API")
    ap.add_argument("--no-overlap", action="store_true",
                    help="graph path: run the optimizer at the end of its own step instead of next to the next "
                         "step's marching")
    ap.add_argument("--ncu-window", type=int, default=0,
                    help="profiling aid: wrap this many extra steps in cudaProfilerStart/Stop "
                         "(use with `ncu --profile-from-start off`); numbers printed under ncu are not bench values")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
