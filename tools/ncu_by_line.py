#!/usr/bin/env python3
"""Join an ncu SASS source page (per-instruction executed counts + stall samples) with nvdisasm line info of the cubin,
and aggregate by source line / function (inlined call chain innermost line).  Usage:
   ncu_by_line.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top]"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
lines = []   # (offset, file, line, inline_chain)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((int(m.group(1), 16), cur[0], cur[1], m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
base = int(data[0][ia], 16)
byoff = {o: (f, ln, ins) for o, f, ln, ins in lines}
agg = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
src_cache = {}
for r in data:
    off = int(r[ia], 16) - base
    f, ln, ins = byoff.get(off, ("?", 0, ""))
    n = int(r[ie]); s = int(r[iss])
    agg[(f, ln)] += n; samp[(f, ln)] += s; tot += n; tots += s
def srcline(f, ln):
    for d in ("motion-planning-for-autonomous-driving-with-mpc_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:110] if 0 < ln <= len(L) else ""
    return ""
print(f"total warp instructions {tot}, stall samples {tots}")
print("--- by instructions executed")
for (f, ln), n in agg.most_common(top):
    print(f"{100*n/tot:5.1f}% inst {100*samp[(f,ln)]/max(tots,1):5.1f}% samp  {f}:{ln}  {srcline(f, ln)}")
print("--- by stall samples")
for (f, ln), s in samp.most_common(top):
    print(f"{100*s/max(tots,1):5.1f}% samp {100*agg[(f,ln)]/tot:5.1f}% inst  {f}:{ln}  {srcline(f, ln)}")
