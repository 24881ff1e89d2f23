#!/usr/bin/env python
"""SASS opcode census of libsoundbubble_sm100a.so (cuobjdump -sass): per kernel, how many tensor-core (UTCHMMA), tensor-memory
(LDTM / STTM), TMA (UBLKCP = bulk copy, UTMALDG / UTMASTG / UTMAREDG = tensor-map load / store / reduce-store), packed-FMA (FFMA2), MUFU and mbarrier
(SYNCS) instructions the build contains.  Evidence that the tcgen05 / TMA paths are what was compiled, next to the ncu captures.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sound_bubble_b200", "libsoundbubble_sm100a.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTMAREDG", "FFMA2", "FFMA", "MUFU", "SYNCS", "LDGSTS", "RED", "ATOMG", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n   # noqa: E731
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in OPS:
                if op == k or (k != "FFMA" and op.startswith(k)):
                    counts[cur][k] += 1
    tot = collections.Counter()
    print("# cuobjdump -sass %s ; instructions per kernel" % os.path.basename(LIB))
    print("%-64s %7s " % ("kernel", "instr") + " ".join("%7s" % k for k in OPS))
    for fn in order:
        c = counts[fn]
        name = re.sub(r"\(.*", "", demangle(fn))[:64]
        print("%-64s %7d " % (name, c["_total"]) + " ".join("%7d" % c[k] for k in OPS))
        tot.update(c)
    print("%-64s %7d " % ("TOTAL (%d kernels)" % len(order), tot["_total"]) + " ".join("%7d" % tot[k] for k in OPS))


if __name__ == "__main__":
    main()
