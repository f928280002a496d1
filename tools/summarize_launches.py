"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch
count, total and mean device time, share of the total.  Usage: summarize_launches.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v))
agg = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    key = name.split("(")[0][:90]
    agg[key][0] += 1
    agg[key][1] += us
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total / 1e3:.3f} ms of device time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':92s} {'n':>6s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:92s} {n:6d} {us:12.1f} {us / n:10.2f} {100 * us / total:6.2f}%")
ours = {k: v for k, v in agg.items() if any(s in k for s in ("gemv_kernel", "snapshot_u8", "gemm_tc", "gemm_simt", "iadb_step", "pack_kernel", "epilogue_kernel", "combine_kernel", "groupnorm_nhwc", "add_bias_nhwc", "attention_small", "linear_tc", "shortcut_tc", "conv_in3x3", "upsample2x", "tile_L", "tri_check", "ddim_step", "to_u8", "white128"))}
print("-- kernels of libbndm_b200.so")
for k, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:92s} {n:6d} {us:12.1f} {us / n:10.2f} {100 * us / total:6.2f}%")
