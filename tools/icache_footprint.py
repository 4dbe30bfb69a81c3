#!/usr/bin/env python
"""Dynamic instruction footprint of a kernel from an ncu --set full report + the matching library (-lineinfo):
   icache_footprint.py <rep> <lib.so> <kernel substring> <chain_iterations>
Prints, per source line, the static SASS count of the code executed about once per chain-iteration (the part that does not
stay in the instruction cache) and of the hot loops, plus no_instruction stall samples."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, so, kern, chain_iters = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."): infn = kern in ln; continue
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
assert len(lines) == len(data), (len(lines), len(data))
num = lambda x: float(x) if x not in ("", "-") else 0.0
cold = collections.Counter(); hot = collections.Counter(); noinst = collections.Counter(); never = 0
for fl, r in zip(lines, data):
    f = num(r[col["Instructions Executed"]]) / chain_iters
    noinst[fl] += num(r[col["stall_no_inst"]])
    if f == 0: never += 1
    elif f < 6: cold[fl] += 1
    else: hot[fl] += 1
print(f"static: {len(lines)} SASS; never executed {never}; cold (<6 warp-executions per chain-iteration) {sum(cold.values())} = {sum(cold.values()) * 16 / 1024:.1f} KB; hot {sum(hot.values())} = {sum(hot.values()) * 16 / 1024:.1f} KB")
tot_ni = sum(noinst.values()) or 1
def agg_file(c):
    d = collections.Counter()
    for (f, l), v in c.items(): d[f] += v
    return dict(d)
print("cold by file:", agg_file(cold)); print("hot by file:", agg_file(hot))
print("\ncold code, largest source lines:")
for (f, l), v in cold.most_common(30): print(f"  {f}:{l}  {v} SASS   no_inst samples {100 * noinst[(f, l)] / tot_ni:.1f}%")
print("\nhot code, largest source lines:")
for (f, l), v in hot.most_common(25): print(f"  {f}:{l}  {v} SASS   no_inst samples {100 * noinst[(f, l)] / tot_ni:.1f}%")
print("\nno_instruction stall samples by line:")
for (f, l), v in noinst.most_common(15): print(f"  {f}:{l}  {100 * v / tot_ni:.1f}%  (cold {cold[(f, l)]}, hot {hot[(f, l)]} SASS)")
