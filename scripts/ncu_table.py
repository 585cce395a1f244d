"""Turns an `ncu --page raw --csv` dump of scripts/kernels_once.py into the per-kernel roofline table of
profiles/ (achieved algorithmic GB/s vs the measured HBM peak, DRAM traffic, L2 %, tensor-pipe %)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALG = {  # algorithmic bytes per sample (SURVEY.md §8d, fp16 encoder) / per ray / per parameter
    "hash_fwd": ("sample", 588), "hash_bwd": ("sample", 1100), "mlp_fwd": ("sample", 86), "mlp_bwd": ("sample", 150),
    "composite_train_fwd": ("sample", 22), "composite_train_bwd": ("sample", 32), "march_train_warp_kernel<1>": ("sample", 32),
    "march_train_warp_kernel<0>": ("ray", 48), "adam_kernel": ("param", 34),
    "march_train_warp_kernel<2>": ("sample", 32), "ray_head_fused_kernel": ("sample", 24),
    "sample_ray_batch_kernel": ("ray", 72), "check_finite_kernel": ("param", 4),
}
rows = list(csv.reader(open(sys.argv[1])))
S = float(sys.argv[2])
n_rays, n_param = 8192.0, 11429472.0
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(f"| kernel | time (us) | alg. bytes/launch | achieved GB/s | frac of HBM peak ({peak:.0f}) | DRAM rd+wr (MB) | lts % | l1tex % | sm % | tensor pipe % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    key = next((k for k in ALG if k in name), None)
    if key is None or key in seen:
        continue
    if key == "adam_kernel" and float(r[idx['launch__grid_size']]) < 500:
        continue  # take the big (hash table) launch
    seen.add(key)
    g = lambda m: float(r[idx[m]].replace(',', ''))
    t_us = g('gpu__time_duration.sum')
    unit = rows[1][idx['gpu__time_duration.sum']]
    t_us = t_us * (1e3 if unit == 'ms' else 1e-3 if unit == 'ns' else 1.0)
    kind, per = ALG[key]
    units = {"sample": S, "ray": n_rays, "param": n_param}[kind]
    alg = per * units
    def mb(m):
        v = g(m); u = rows[1][idx[m]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
    dram = mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum')
    ach = alg / (t_us * 1e-6) / 1e9
    print(f"| {key} | {t_us:.1f} | {alg / 1e6:.1f} MB | {ach:.0f} | {ach / peak:.2f} | {dram:.1f} | "
          f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {int(g('launch__registers_per_thread'))} |")
