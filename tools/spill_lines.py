import collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt = collections.Counter()
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
        m = re.search(r"\b(LDL|STL)\b", ln)
        if m: cnt[(cur, m.group(1))] += 1
for (fl, op), c in sorted(cnt.items(), key=lambda kv: (kv[0][0][0], kv[0][0][1])): print(f"{fl[0]}:{fl[1]} {op} x{c}")
