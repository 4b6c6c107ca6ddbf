#!/usr/bin/env python
"""profiles/sia_kernel_dram.json (read by bench.py for roofline.traffic) from an ncu report of one sia_kernel launch."""
import csv, io, json, subprocess, sys

rep, summary = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, vals = rows[0], rows[1], rows[2]
def val(name):
    i = h.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return int(round(float(vals[i]) * scale))
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"source": f"ncu --set full --clock-control none -k regex:sia_kernel -c 1 ({summary}), one launch of bench.py's default window",
       "kernel": vals[h.index("Kernel Name")], "pairs_per_launch": int(float(vals[h.index("launch__grid_size")])),
       "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr}
json.dump(out, open("profiles/sia_kernel_dram.json", "w"), indent=1)
print(out)
