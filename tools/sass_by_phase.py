#!/usr/bin/env python3
"""Static SASS instruction count of a kernel per solver PHASE (call site inside iterate()) from nvdisasm's inline chains.
Usage: sass_by_phase.py <cubin> <kernel-substr> [core header, default warp_core.cuh]"""
import re, subprocess, sys, collections, os
cubin, kern = sys.argv[1:3]
core_name = sys.argv[3] if len(sys.argv) > 3 else "warp_core.cuh"
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "..", "motion-planning-for-autonomous-driving-with-mpc_b200", "csrc", core_name)).read().splitlines()
it0 = next(i for i, l in enumerate(src) if "void iterate(" in l) + 1
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
agg = collections.Counter(); ops = collections.defaultdict(collections.Counter)
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
            if f == core_name and ln >= it0:
                phase = f"iterate:{ln}  " + src[ln - 1].strip()[:70]
        agg[phase] += 1
        ops[phase][m.group(2).split()[0].split(".")[0]] += 1
tot = sum(agg.values())
print("total SASS instructions", tot)
for ph, n in agg.most_common():
    print(f"{n:6d} {100*n/tot:5.1f}% | {ph}   [" + ", ".join(f"{o} {c}" for o, c in ops[ph].most_common(6)) + "]")
