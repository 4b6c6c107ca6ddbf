#!/usr/bin/env python
"""Static SASS instruction histogram of one kernel by CUDA source line (needs -lineinfo):
   python scripts/sass_hist.py build/sparse_align.o 'sia_kernelILi1ELb0ELi320ELi2ELi1E' [topn]"""
import collections, re, subprocess, sys, tempfile, os
obj, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
cur_fn, line = None, None
total = collections.Counter()
for l in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: cur_fn = m.group(1); continue
    if cur_fn is None or pat not in cur_fn: continue
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,5}\*/", l): total[line] += 1
print("instructions:", sum(total.values()))
byfile = collections.Counter()
for (f, ln), c in total.items(): byfile[f] += c
print(dict(byfile))
for (f, ln), c in sorted(total.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{c:6d}  {f}:{ln}")
