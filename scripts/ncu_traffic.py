"""Condenses `ncu -i <rep> --page raw --csv` of `bench.py --ncu-window N` (graph-node profiling) into
  * a per-kernel markdown table (time, algorithmic bytes -> achieved GB/s vs the measured HBM peak, DRAM traffic,
    L2 / L1TEX / tensor-pipe utilisation), and
  * profiles/r2_traffic.json: DRAM bytes per launch of each kernel, keyed by bench config, which bench.py attaches
    to `roofline.traffic` when the live sample count matches the captured one.

    python scripts/ncu_traffic.py <raw.csv> <config> <samples per step> [--json profiles/r2_traffic.json] > table.md
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = [("hash_fwd_kernel", "hash_fwd"), ("hash_bwd_kernel", "hash_bwd"), ("mlp_fwd", "mlp_fwd"),
        ("mlp_bwd", "mlp_bwd"), ("ray_head_fused", "ray_head"), ("adam_kernel", "adam"),
        ("march_train_warp_kernel", "march"), ("composite_round", "composite_round"),
        ("sample_ray_batch", "sampler"), ("check_finite", "check_finite"), ("ray_aabb", "ray_aabb")]


def main():
    raw, config, S = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {h: i for i, h in enumerate(rows[hdr])}
    units = rows[hdr + 1]
    cfg = bench.CONFIGS[config]
    col = 0 if cfg["half"] else 1
    hbm, tf, _ = bench.measured_peaks()
    agg = {}
    for r in rows[hdr + 2:]:
        if len(r) <= idx["Kernel Name"]:
            continue
        name = r[idx["Kernel Name"]]
        key = next((k for pat, k in KEYS if pat in name), None)
        if key is None:
            continue

        def g(m, r=r):
            v = float(r[idx[m]].replace(",", "") or 0)
            u = units[idx[m]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
                        "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        e = agg.setdefault(key, {"n": 0, "us": 0.0, "dram": 0.0, "lts": 0.0, "l1": 0.0, "tensor": 0.0, "warps": 0.0,
                                 "regs": 0, "name": name.split("(")[0]})
        if key == "adam" and float(r[idx["launch__grid_size"]]) < 500:
            continue
        e["n"] += 1
        e["us"] += g("gpu__time_duration.sum")
        e["dram"] += g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
        e["lts"] += g("lts__throughput.avg.pct_of_peak_sustained_elapsed")
        e["l1"] += g("l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
        e["tensor"] += g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        e["warps"] += g("sm__warps_active.avg.pct_of_peak_sustained_active")
        e["regs"] = int(g("launch__registers_per_thread"))
    print(f"| kernel | launches | time/launch (us) | alg. MB/launch | achieved GB/s | frac of HBM peak ({hbm:.0f}) | "
          f"DRAM rd+wr MB/launch | lts % | l1tex % | tensor pipe % | warps active % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for key, e in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = max(e["n"], 1)
        us = e["us"] / n
        if key in bench.BYTES_PER_SAMPLE:
            alg = bench.BYTES_PER_SAMPLE[key][col] * S
        elif key == "adam":
            alg = bench.ADAM_BYTES_PER_PARAM * 11429472
        else:
            alg = 0.0
        ach = alg / (us * 1e-6) / 1e9 if us > 0 else 0.0
        traffic[key] = {"dram_bytes": e["dram"] / n, "time_us": us, "launches": e["n"],
                        "lts_pct": e["lts"] / n, "l1tex_pct": e["l1"] / n, "tensor_pct": e["tensor"] / n}
        print(f"| {key} | {e['n']} | {us:.1f} | {alg / 1e6:.1f} | {ach:.0f} | {ach / hbm:.2f} | {e['dram'] / n / 1e6:.1f} | "
              f"{e['lts'] / n:.0f} | {e['l1'] / n:.0f} | {e['tensor'] / n:.1f} | {e['warps'] / n:.0f} | {e['regs']} |")
    if out_json:
        d = {}
        if os.path.exists(out_json):
            d = json.load(open(out_json))
        d[config] = {"samples": S, "kernels": traffic,
                     "source": f"ncu --set full --graph-profiling node of `python bench.py --config {config} --ncu-window 1` "
                               f"(profiles/, {os.path.basename(raw)}): dram__bytes_read.sum + dram__bytes_write.sum per launch"}
        json.dump(d, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
