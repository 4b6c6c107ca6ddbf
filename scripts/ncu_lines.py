#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + stall samples per CUDA source line (needs -lineinfo)."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor','sm__cycles_elapsed.max','launch__shared_mem_per_block_dynamic','lts__t_bytes.sum','l1tex__t_bytes.sum',
        'smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fp64.sum','launch__grid_size','launch__block_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in h:
        i = h.index(w); print(f"{w}: {[r[i] for r in rows[1:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out = collections.defaultdict(lambda: collections.Counter()); cur = None; cols = None; seen_first = False
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No': cols = r; continue
    if cols is None or len(r) < 8 or r[2] != '-': continue
    try: ln = int(r[0])
    except ValueError: continue
    k = (cur.split('/')[-1], ln, r[1].strip()[:88])
    for name in ('# Samples', 'Instructions Executed', 'stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_lg'):
        try: out[k][name] += int(r[cols.index(name)])
        except (ValueError, IndexError): pass
tot = sum(v['# Samples'] for v in out.values()) or 1
agg = collections.Counter()
for v in out.values():
    for n, c in v.items(): agg[n] += c
print("total samples", tot, {n: f"{100*c/tot:.1f}%" for n, c in agg.items() if n.startswith('stall')}, "inst", agg['Instructions Executed'])
for k, v in sorted(out.items(), key=lambda kv: -kv[1]['# Samples'])[:topn]:
    print(f"{v['# Samples']:7d} {100*v['# Samples']/tot:5.1f}% inst {v['Instructions Executed']:9d} bar {v['stall_barrier']:6d} lsb {v['stall_long_sb']:5d} ssb {v['stall_short_sb']:5d} wait {v['stall_wait']:5d} | {k[0]}:{k[1]} {k[2]}")
