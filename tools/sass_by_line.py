#!/usr/bin/env python
"""Static SASS instruction count per source file:line of one kernel (needs -lineinfo): sass_by_line.py <lib.so> <kernel substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt = collections.Counter(); total = 0
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            infn = kern in ln; continue
        if ln.startswith("\t.section"): infn = False
        if not infn: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s", ln): cnt[cur] += 1; total += 1
print("total SASS", total, "=", total * 16 // 1024, "KB")
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
print("by file:", dict(byfile))
if os.environ.get("RANGES"):  # "name:lo-hi,..." line ranges of sampler_kernel.cuh
    for part in os.environ["RANGES"].split(","):
        name, r = part.split(":"); lo, hi = map(int, r.split("-"))
        print(f"  {name:10s} lines {lo}-{hi}: {sum(c for (f, l), c in cnt.items() if f.startswith('sampler_kernel') and lo <= l <= hi)}")
for (f, l), c in cnt.most_common(top): print(f"{f}:{l}  {c}")
