#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-specific SASS opcodes in the built library:

    python profiles/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt

UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4, LDTM = tcgen05.ld (TMEM -> registers),
UTMALDG = TMA tensor load (cp.async.bulk.tensor), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops,
UTCATOMSWS / UTCALLOC-style ops = TMEM allocation.  (cuobjdump -sass of deepsee_b200/lib/libdeepsee_b200.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "deepsee_b200", "lib", "libdeepsee_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    hist = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    hist[cur][k] += 1
                    break
            hist[cur]["_all"] += 1
    print("%-64s %6s " % ("kernel (sm_100a SASS)", "instr") + " ".join("%8s" % k for k in KEYS))
    for name, c in hist.items():
        if c["UTCHMMA"] or c["UTCQMMA"] or c["UTMALDG"] or "--all" in sys.argv:
            print("%-64s %6d " % (name[:64], c["_all"]) + " ".join("%8d" % c[k] for k in KEYS))
    tot = collections.Counter()
    for c in hist.values():
        tot.update(c)
    print("%-64s %6d " % ("TOTAL over %d kernels" % len(hist), tot["_all"]) + " ".join("%8d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    main()
