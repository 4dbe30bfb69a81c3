#!/usr/bin/env python
"""Dynamic warp-instruction count and samples per PHASE of tg_sampler_kernel from an ncu --set full report.
usage: ncu_phases.py <report.ncu-rep> <lib.so> [kernel-substring]
SASS instructions are attributed to the sampler_kernel.cuh function whose line range holds their source line; instructions
inlined from other headers inherit the phase of the nearest preceding instruction that has a sampler_kernel.cuh line."""
import csv, io, os, re, subprocess, sys, tempfile
rep, so = sys.argv[1:3]
kern = sys.argv[3] if len(sys.argv) > 3 else "tg_sampler_kernelILi3ELb0"
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc-in-tonga_b200", "csrc", "sampler_kernel.cuh")
# phase = function name containing the line
funcs = []
for i, ln in enumerate(open(SRC).read().splitlines(), 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:static )?__(?:device|global)__ .*?\b(\w+)\(", ln)
    if m and not ln.startswith(" "):
        funcs.append((i, m.group(1)))
def phase_of(line):
    name = "?"
    for l0, n in funcs:
        if l0 <= line: name = n
    if name == "tg_sampler_kernel":
        src = open(SRC).read().splitlines()
        # sub-phases of the main loop by marker comments
        marks = [(i, t) for i, t in enumerate(src, 1) if "// ====" in t and i > funcs[-1][0]]
        sub = "prologue"
        for l0, t in marks:
            if l0 <= line: sub = t.split("=")[-1].strip()[:28] or sub
        return "main:" + sub
    return name
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []
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
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln): lines.append(cur)
    if lines: break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
col = {h: i for i, h in enumerate(hdr)}
num = lambda x: float(x) if x not in ("", "-") else 0.0
agg, cur = {}, "?"
for (fl, r) in zip(lines, data):
    if fl[0].startswith("sampler_kernel"): cur = phase_of(fl[1])
    a = agg.setdefault(cur, [0.0, 0.0, 0])
    a[0] += num(r[col["# Samples"]]); a[1] += num(r[col["Instructions Executed"]]); a[2] += 1
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"total warp instructions {ti:.4g}, samples {ts:.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} inst {100*a[1]/ti:5.1f}%  samples {100*a[0]/ts:5.1f}%  static {a[2]:5d}")
