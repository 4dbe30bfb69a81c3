#!/usr/bin/env python
"""LSU / L1 load of tg_sampler_kernel per source line and per phase from an ncu --set full report (source page):
shared-memory wavefronts (and the excess from bank conflicts), global L1 tag requests, global L2 sectors.
usage: ncu_lsu.py <report.ncu-rep> <lib.so> [kernel-substring] [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so = sys.argv[1:3]
kern = sys.argv[3] if len(sys.argv) > 3 else "tg_sampler_kernelILi3ELb0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc-in-tonga_b200", "csrc", "sampler_kernel.cuh")
src = open(SRC).read().splitlines()
funcs = []
for i, ln in enumerate(src, 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:static )?__(?:device|global)__ .*?\b(\w+)\(", ln)
    if m and not ln.startswith(" "): funcs.append((i, m.group(1)))
def phase_of(line):
    name = "?"
    for l0, n in funcs:
        if l0 <= line: name = n
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
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
col = {h: i for i, h in enumerate(hdr)}
num = lambda x: float(x) if x not in ("", "-") else 0.0
C = ["Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L1 Tag Requests Global", "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Local"]
byline, byphase, cur = collections.defaultdict(lambda: [0.0] * len(C)), collections.defaultdict(lambda: [0.0] * len(C)), "?"
curline = ("?", 0)
for fl, r in zip(lines, data):
    if fl[0].startswith("sampler_kernel"): cur, curline = phase_of(fl[1]), fl
    v = [num(r[col[c]]) for c in C]
    for k in range(len(C)): byline[(curline, r[col["Source"]].split()[0] if False else "")][k] += v[k]; byphase[cur][k] += v[k]
tot = [sum(a[k] for a in byphase.values()) for k in range(len(C))]
print("totals:", {c: f"{t:.4g}" for c, t in zip(C, tot)})
print(f"{'phase':28s} " + " ".join(f"{c[:22]:>22s}" for c in C))
for ph, a in sorted(byphase.items(), key=lambda kv: -(kv[1][1] + kv[1][3])):
    print(f"{ph:28s} " + " ".join(f"{100 * a[k] / max(tot[k], 1):21.1f}%" for k in range(len(C))))
print("\nper nearest sampler_kernel.cuh line (smem wavefronts + global tag requests, top %d):" % top)
for (fl, _), a in sorted(byline.items(), key=lambda kv: -(kv[1][1] + kv[1][3]))[:top]:
    text = src[fl[1] - 1].strip()[:70] if fl[0].startswith("sampler_kernel") and 0 < fl[1] <= len(src) else ""
    print(f"{fl[0][:14]}:{fl[1]:<5d} inst {100*a[0]/tot[0]:4.1f}%  smem wf {100*a[1]/max(tot[1],1):4.1f}% (excess {100*a[2]/max(tot[1],1):4.1f}%)  glob tag {100*a[3]/max(tot[3],1):4.1f}%  L2 sect {100*a[4]/max(tot[4],1):4.1f}% | {text}")
