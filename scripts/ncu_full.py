#!/usr/bin/env python
"""Everything ncu knows about one kernel that bears on "what is it bound by": unit throughputs, pipe utilisation, L1
wavefronts, the complete warp-stall taxonomy, and the stall samples bucketed by source-line ranges.

   python scripts/ncu_full.py gpurun_out/<tag>_sia.ncu-rep [file.cu:lo-hi=name ...]
"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
buckets = []
for a in sys.argv[2:]:
    m = re.match(r"([^:]+):(\d+)-(\d+)=(\S+)", a)
    if m: buckets.append((m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
pat = re.compile(r"(throughput|pipe_.*(active|pct)|issue_stalled.*per_warp_active|issue_stalled.*ratio|wavefronts|"
                 r"l1tex__t_(requests|sectors|bytes)|l1tex__data_bank|lsu_mem|inst_executed(\.sum|_pipe_(lsu|fp64|fma|alu|xu|uniform))|"
                 r"warps_(active|eligible)|issue_active|cycles_elapsed\.max|duration|registers|local_(load|store)|lts__t_(bytes|sectors)\.sum|"
                 r"dram__bytes|smsp__inst_executed_op_(shared|global|local)|hit_rate|one_or_more_eligible|no_instruction|occupancy)")
for r in rows[2:]:
    print("== kernel:", r[h.index("Kernel Name")])
    seen = []
    for i, name in enumerate(h):
        if pat.search(name) and r[i] not in ("", "n/a"):
            seen.append((name, units[i], r[i]))
    for name, u, v in sorted(seen):
        print(f"{name:100s} {u:14s} {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, cols = None, None
per_line = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": cols = r; continue
    if cols is None or len(r) < 8 or r[2] != "-": continue
    try: ln = int(r[0])
    except ValueError: continue
    for j, name in enumerate(cols):
        if name in ("# Samples", "Instructions Executed") or name.startswith("stall_"):
            try: per_line[(cur, ln)][name] += int(r[j])
            except (ValueError, IndexError): pass
tot = sum(v["# Samples"] for v in per_line.values()) or 1
agg = collections.Counter()
for v in per_line.values():
    for n, c in v.items(): agg[n] += c
print("\n== stall samples (all reasons), total", tot)
for n, c in sorted(agg.items(), key=lambda kv: -kv[1]):
    if n.startswith("stall_"): print(f"  {n:28s} {100*c/tot:5.1f}%")
print("instructions executed:", agg["Instructions Executed"])
if buckets:
    print("\n== by source range")
    b = collections.defaultdict(collections.Counter)
    for (f, ln), v in per_line.items():
        name = "other"
        for bf, lo, hi, nm in buckets:
            if f == bf and lo <= ln <= hi: name = nm; break
        for n, c in v.items(): b[name][n] += c
    for name, v in sorted(b.items(), key=lambda kv: -kv[1]["# Samples"]):
        top = sorted(((n, c) for n, c in v.items() if n.startswith("stall_")), key=lambda kv: -kv[1])[:5]
        print(f"  {name:16s} samples {100*v['# Samples']/tot:5.1f}%  inst {100*v['Instructions Executed']/max(1,agg['Instructions Executed']):5.1f}%  " +
              " ".join(f"{n[6:]}={100*c/tot:.1f}%" for n, c in top))
