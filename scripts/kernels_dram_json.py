#!/usr/bin/env python
"""profiles/r02_kernels_dram.json (read by bench.py's roofline_by_kernel) + a per-kernel text summary from ncu reports:
   python scripts/kernels_dram_json.py gpurun_out/r02_kernels.ncu-rep [more.ncu-rep ...]
For every kernel name the LAST captured launch is used (the earlier ones are warm-ups of the same call)."""
import csv, io, json, os, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
out, lines = {}, []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("svo::", "")
        key = name.split("<")[0]
        if key == "sia_kernel":
            key = name  # keep the template arguments: full-batch and cluster geometries are different kernels
        d = {}
        for w in WANT:
            if w in h:
                i = h.index(w)
                try:
                    d[w] = float(r[i]) * SCALE.get(units[i], 1)
                except ValueError:
                    d[w] = r[i]
        d["dram_bytes"] = int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
        d["duration_us_under_ncu"] = d.get("gpu__time_duration.sum")
        d["source"] = os.path.basename(rep)
        out[key] = d
for k, d in out.items():
    lines.append(f"{k}: {d['duration_us_under_ncu']:.1f} us under ncu, grid {int(d.get('launch__grid_size', 0))} x {int(d.get('launch__block_size', 0))}, "
                 f"{int(d.get('launch__registers_per_thread', 0))} regs, DRAM {d['dram_bytes'] / 1e6:.3f} MB "
                 f"({d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.2f} % of peak), warps active "
                 f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} %, issue active "
                 f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} %, L2 bytes {d.get('lts__t_bytes.sum', 0) / 1e6:.2f} MB")
json.dump(out, open("profiles/r02_kernels_dram.json", "w"), indent=1)
open("profiles/r02_kernels_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
