#!/usr/bin/env python3
"""Per-phase stall-reason breakdown from an ncu SOURCE page csv (`ncu -i rep --page source --csv`) + the kernel's cubin.
Usage: ncu_stall_map.py <source.csv> <cubin> <kernel-substr> [core header]  -- prints, per phase, instructions executed and the
samples of each stall reason, then the 25 instructions with the most no_instruction samples."""
import csv, re, subprocess, sys, collections, os
srccsv, cubin, kern = sys.argv[1:4]
core_name = sys.argv[4] if len(sys.argv) > 4 else "warp_core.cuh"
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "..", "motion-planning-for-autonomous-driving-with-mpc_b200", "csrc", core_name)).read().splitlines()
it0 = next(i for i, l in enumerate(src) if "void iterate(" in l) + 1
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
info = {}; chain = []; fresh = True
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
                phase = f"{ln} " + src[ln - 1].strip()[:40]
        info[int(m.group(1), 16)] = (phase, chain[0] if chain else ("?", 0), m.group(2))
rows = list(csv.reader(open(srccsv))); hdr = rows[1]; data = rows[2:]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {r: hdr.index(r) for r in reasons}
base = int(data[0][ia], 16)
agg = collections.defaultdict(collections.Counter); inst = collections.Counter(); tops = []
for r in data:
    off = int(r[ia], 16) - base
    ph, leaf, op = info.get(off, ("?", ("?", 0), "?"))
    inst[ph] += int(r[ie])
    for k, i in idx.items():
        v = int(r[i] or 0)
        agg[ph][k] += v
    tops.append((int(r[idx["stall_no_inst"]] or 0), off, ph, leaf, op, int(r[ie])))
tot = collections.Counter()
for ph in agg: tot.update(agg[ph])
allsum = sum(tot.values())
print("all samples", allsum, {k: f"{100*v/allsum:.1f}%" for k, v in tot.most_common(8)})
for ph, n in inst.most_common(12):
    s = sum(agg[ph].values())
    print(f"{100*n/sum(inst.values()):5.1f}% inst {100*s/allsum:5.1f}% samp | {ph:50s} | " + ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in agg[ph].most_common(5)))
print("top no_inst instructions:")
for v, off, ph, leaf, op, n in sorted(tops, reverse=True)[:25]:
    print(f"  {v:5d} samples  @{off:05x} exec {n:8d}  {ph[:34]:34s} {leaf[0]}:{leaf[1]}  {op[:60]}")
