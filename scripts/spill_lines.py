#!/usr/bin/env python
"""List local-memory (spill) instructions and barriers of one kernel by CUDA source line:
   python scripts/spill_lines.py build/sparse_align.o 'sia_kernelILi1ELb0ELi320ELi2ELi1E'   (needs -lineinfo)"""
import collections, re, subprocess, sys, tempfile, os
obj, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
cur_fn, line, cnt = None, None, collections.Counter()
total = collections.Counter()
for l in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: cur_fn = m.group(1); continue
    if cur_fn is None or pat not in cur_fn: continue
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.search(r"\b(STL|LDL|BAR\.SYNC|LDS|STS|LDG|DFMA|DMUL|DADD|SHFL|MUFU|F2F|BSSY|CALL)[\.\w]*", l)
    if re.search(r"/\*[0-9a-f]{4}\*/", l): total[line] += 1
    if m and m.group(1) in ("STL", "LDL", "BAR.SYNC", "CALL"): cnt[(line, m.group(1))] += 1
for (ln, op), c in sorted(cnt.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(ln, op, c)
print("instructions:", sum(total.values()))
