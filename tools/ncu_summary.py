"""Per-kernel means of an `ncu --set full` capture: python tools/ncu_summary.py raw.csv  (raw.csv = ncu -i x.ncu-rep --page raw --csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}
WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB"), ("dram__bytes_write.sum", "MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def val(r, name):
    try:
        return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)
    except (KeyError, ValueError):
        return float("nan")


groups = {}
for r in rows[2:]:
    groups.setdefault(r[col["Kernel Name"]].split("(")[0][:70], []).append(r)
print(f"{'kernel':72s} {'n':>3s} " + " ".join(f"{u:>8s}" for _, u in WANT))
for k, rs in sorted(groups.items(), key=lambda kv: -sum(val(r, "gpu__time_duration.sum") for r in kv[1])):
    cells = []
    for name, u in WANT:
        m = sum(val(r, name) for r in rs) / len(rs)
        cells.append(f"{m / 1e6 if u == 'MB' else m:8.2f}")
    print(f"{k:72s} {len(rs):3d} " + " ".join(cells))
