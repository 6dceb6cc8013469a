#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass libsbte_b200.so): the mnemonics that prove
the sm_100a data paths -- UBLKCP (cp.async.bulk), UTMALDG (TMA tensor copy), SYNCS (mbarrier), UCGABAR (cluster
barrier), USETMAXREG, ACQBULK/fence, DFMA/DMUL/DADD (the FP64 pipe) -- and, for contrast, the absence of tensor-core
ops (the contraction is a gathered product, not a GEMM).    python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectralbte_b200", "libsbte_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "USETMAXREG", "DFMA", "DMUL", "DADD", "LDS", "LDG", "STG", "R2UR", "MUFU",
        "HMMA", "DMMA", "UTCMMA", "ATOMG", "RED", "NANOSLEEP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if fn and m:
            op = m.group(1)
            per[fn][op] += 1
            per[fn]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode counts per kernel of spectralbte_b200/libsbte_b200.so (sm_100a), static instruction counts")
    print("# %-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
    tot = collections.Counter()
    for (fn, cnt), name in zip(per.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("void sbte::", "").replace("sbte::", "")
        print("%-80s %7d " % (name[:80], cnt["_total"]) + " ".join("%7d" % cnt[k] for k in KEYS))
        tot.update(cnt)
    print("%-80s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
