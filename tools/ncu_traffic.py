"""profiles/ncu_traffic.json (read by bench.py for roofline.traffic) and a text summary from an `ncu --set full` capture of the windowed forward:
   ncu -i gpurun_out/prof_window_r02.ncu-rep --page raw --csv > /tmp/win_raw.csv;  python tools/ncu_traffic.py /tmp/win_raw.csv r02"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)


groups = {}
for r in rows[2:]:
    k = "recurrent_step" if "<1, 1," in r[col["Kernel Name"]] or "<0, 1," in r[col["Kernel Name"]] else "fused_window"
    groups.setdefault(k, []).append(r)
out, lines = {}, []
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size"]
for k, rs in groups.items():
    n = len(rs)
    avg = {w: sum(val(r, w) for r in rs) / n for w in WANT if w in col}
    out[k] = {"launches": n, "kernel": rs[0][col["Kernel Name"]].split("(")[0], "dram_bytes_read": avg["dram__bytes_read.sum"],
              "dram_bytes_write": avg["dram__bytes_write.sum"], "duration_us": avg["gpu__time_duration.sum"],
              "tensor_pipe_pct": avg.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")}
    lines.append(f"== {k}: {n} launches of {out[k]['kernel']} (averages per launch)")
    lines += [f"   {w:75s} {v:16.3f}" for w, v in avg.items()]
d = {"source": f"profiles/{tag}_ncu_window_fwd.txt (ncu --set full --clock-control none, tools/window_once.py)",
     "dram_bytes_read": out["recurrent_step"]["dram_bytes_read"], "dram_bytes_write": out["recurrent_step"]["dram_bytes_write"],
     "kernels": out,
     "note": "top-level bytes = the roofline kernel of bench.py (recurrent fused cell step), per launch; writes that still sit in the 126 MB L2 when "
             "the kernel ends are not in dram__bytes_write"}
json.dump(d, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_window_fwd.txt"), "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:lif_conv_fwd_tc -s 50 -c 25 python tools/window_once.py\n"
    "one inference window (B 8, 128x128, T 10) of the bench model through LIFFireNet.forward_window, direct launches\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
