#!/usr/bin/env python
"""Attribute ncu SASS-level samples / instruction counts to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-name-substring> [top]
Needs the library compiled with -lineinfo.  (ncu's own `--print-source cuda` CSV carries no metrics.)
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []  # (file, line) per SASS instruction of the kernel, in order
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            infn = kern in ln
            continue
        if ln.startswith("\t.section"):
            infn = False
        if not infn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    if lines:
        break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
print(f"SASS instructions: disasm {len(lines)}, ncu {len(data)}")
col = {h: i for i, h in enumerate(hdr)}
num = lambda x: float(x) if x not in ("", "-") else 0.0
agg = {}
keys = ["# Samples", "Instructions Executed", "stall_no_inst", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_branch_resolving"]
for (fl, r) in zip(lines, data):
    a = agg.setdefault(fl, [0.0] * len(keys) + [0])
    for k, key in enumerate(keys):
        a[k] += num(r[col[key]])
    a[-1] += 1
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts:.0f}, warp instructions {ti:.3g}")
src = {}
def srcline(f, l):
    if f not in src:
        for root in ("mcmc-in-tonga_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src[f] = open(p).read().splitlines(); break
        else:
            src[f] = []
    return src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
print("file:line            samp%  inst%  nSASS | no_inst barrier long_sb short_sb wait math branch | source")
for fl, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get("SORT") == "inst" else 0])[:top]:
    print(f"{fl[0][:14]}:{fl[1]:<5} {100*a[0]/ts:5.1f}% {100*a[1]/ti:5.1f}% {a[-1]:5d} | " + " ".join(f"{100*v/ts:4.1f}" for v in a[2:9]) + " | " + srcline(*fl))
