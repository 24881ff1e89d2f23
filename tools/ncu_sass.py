#!/usr/bin/env python
"""Prints the SASS of one kernel from an .ncu-rep with per-instruction sample counts and top stall reason.
    python tools/ncu_sass.py gpurun_out/prof_ws.ncu-rep [kernel-substring] [min_exec]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
min_exec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
done = False
for b in blocks[1:]:
    rr = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    if sub not in rr[0][1] or done:
        continue
    done = True
    h = rr[1]; idx = {x: i for i, x in enumerate(h)}
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    print(rr[0][1])
    for r in rr[2:]:
        if len(r) < len(h):
            continue
        ex = int(r[idx["Instructions Executed"]] or 0)
        if ex < min_exec:
            continue
        n = int(r[idx["# Samples"]] or 0)
        top = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        tops = " ".join("%s:%d" % (s, v) for v, s in top if v)
        print("%5d %8d  %-70s %s" % (n, ex, r[idx["Source"]].strip()[:70], tops))
