#!/usr/bin/env python
"""Per-kernel SASS opcode census of csrc/liblewin_b200.so (cuobjdump -sass): the mnemonics that prove which datapath a
kernel uses — UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), UTCBAR
(tcgen05.commit), HMMA (mma.sync), LDSM (ldmatrix), LDGSTS (cp.async), SYNCS (mbarrier).  Runs without a GPU.

    python scripts/sass_census.py > profiles/r2_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "research-and-implementation-of-image-dehazing-algorithm-based-on-vision-transformer_b200", "csrc", "liblewin_b200.so")
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "SYNCS", "MUFU", "FFMA2", "ATOMS", "RED")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    names = list(per)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
    tot = collections.Counter()
    print("# SASS opcode census of liblewin_b200.so (sm_100a), one line per kernel: instructions, then the counts of the datapath mnemonics")
    print("# " + " ".join(OPS))
    for k in names:
        c = per[k]
        d = re.sub(r"\(.*", "", demangle.get(k, k))
        d = d.replace("void ", "").replace("lewin::", "").replace("(anonymous namespace)::", "")
        cols = " ".join(f"{op}={c[op]}" for op in OPS if c[op])
        print(f"{d[:110]:110s} n={c['_total']:6d}  {cols}")
        for op in OPS:
            tot[op] += c[op]
    print("# totals: " + " ".join(f"{op}={tot[op]}" for op in OPS))


if __name__ == "__main__":
    sys.exit(main())
