#!/usr/bin/env python3
"""Attribute executed warp instructions / stall samples of an ncu report to the solver PHASE (the call site inside
WarpSolver::iterate() / ForcesSolver::iterate()) using nvdisasm's inline chains.
Usage: ncu_by_phase.py <report.ncu-rep> <lib.so> <kernel-substr> [core header, default warp_core.cuh]"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern = sys.argv[1:4]
core_name = sys.argv[4] if len(sys.argv) > 4 else "warp_core.cuh"
here = os.path.dirname(os.path.abspath(__file__))
core = os.path.join(here, "..", "motion-planning-for-autonomous-driving-with-mpc_b200", "csrc", core_name)
src = open(core).read().splitlines()
it0 = next(i for i, l in enumerate(src) if "void iterate(" in l) + 1
it1 = len(src)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = start = None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):       # one cubin per translation unit: take the one that holds the kernel
    d = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    st = next((i for i, l in enumerate(d) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")), None)
    if st is not None:
        dis, start = d, st
        break
if dis is None:
    sys.exit(f"kernel {kern} not found in {so}")
info = {}
chain = []; fresh = True
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."):
        break
    if "//## File" in l:
        if fresh: chain = []; fresh = False
        for m in re.finditer(r'"([^"]+)", line (\d+)', l):
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        fresh = True
        phase = "kernel (I/O, init, loop)"
        for f, ln in chain:
            if f == core_name and it0 <= ln <= it1:
                phase = f"iterate:{ln}  " + src[ln - 1].strip()[:70]
        leaf = chain[0] if chain else ("?", 0)
        info[int(m.group(1), 16)] = (phase, leaf, m.group(2).split()[0])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out)); hdr = rows[1]
ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]; base = int(data[0][ia], 16)
agg = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter); tot = tots = 0
for r in data:
    ph, leaf, op = info.get(int(r[ia], 16) - base, ("?", ("?", 0), "?"))
    n, s = int(r[ie]), int(r[iss])
    agg[ph] += n; samp[ph] += s; tot += n; tots += s
    ops[ph][op.split(".")[0]] += n
print(f"total warp instructions {tot}, stall samples {tots}")
for ph, n in agg.most_common():
    top = ", ".join(f"{o} {100*c/n:.0f}%" for o, c in ops[ph].most_common(6))
    print(f"{100*n/tot:5.1f}% inst {100*samp[ph]/max(tots,1):5.1f}% samp | {ph}\n        [{top}]")
