"""Per-launch summary of an `ncu --set full` report: duration, DRAM traffic, DRAM / tensor-pipe /
L2 utilisation, registers, grid.  Usage: python tools/ncu_summary.py prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

COLS = [("Kernel Name", "kernel", 44), ("launch__grid_size", "grid", 6), ("launch__registers_per_thread", "regs", 5),
        ("gpu__time_duration.sum", "dur_us", 8), ("dram__bytes_read.sum", "dram_rd_MB", 10),
        ("dram__bytes_write.sum", "dram_wr_MB", 10), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 7),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 8),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 6), ("lts__t_sector_hit_rate.pct", "l2_hit_%", 8),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB", 11), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 6)]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; per-launch, cold cache)")
print(" ".join(f"{name:>{w}s}" if i else f"{name:{w}s}" for i, (_, name, w) in enumerate(COLS)))
for r in rows[2:]:
    out = []
    for i, (key, name, w) in enumerate(COLS):
        if key not in hdr:
            out.append(f"{'n/a':>{w}s}")
            continue
        v, u = r[hdr.index(key)], units[hdr.index(key)]
        if i == 0:
            out.append(f"{v.split('(')[0][:w]:{w}s}")
            continue
        try:
            f = float(v.replace(",", ""))
            if u in ("byte", "Kbyte", "Gbyte"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Gbyte": 1e3}[u]
            if u in ("ns", "ms", "second"):
                f *= {"ns": 1e-3, "ms": 1e3, "second": 1e6}[u]
            out.append(f"{f:>{w}.2f}" if f != int(f) or f < 1000 else f"{int(f):>{w}d}")
        except ValueError:
            out.append(f"{v[:w]:>{w}s}")
    print(" ".join(out))
