#!/usr/bin/env python
"""Opcode histogram of lib/libgespmm.so's SASS, per kernel family, with the mnemonics that identify each data path
(LDGSTS = cp.async, UBLKCP / SYNCS = TMA bulk copy + mbarrier, UCGABAR = cluster barrier, LDS/STS, FFMA/FADD, SHFL ...).
CPU only (cuobjdump):  python scripts/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ge-spmm_b200", "lib", "libgespmm.so")
WATCH = ["LDGSTS", "UBLKCP", "SYNCS", "UCGABAR", "LDG", "STG", "LDS", "STS", "FFMA", "FADD", "FMUL", "SHFL", "VOTE", "REDUX", "ELECT",
         "R2UR", "LDGDEPBAR", "DEPBAR", "ATOMS", "BAR", "HMMA", "UTCHMMA", "UTMALDG"]


PREFIXED = ("UCGABAR", "SYNCS", "UBLKCP", "LDGSTS")  # mnemonics with suffixed forms, counted under their stem


def family(name):
    for key in ("WalkerBulk", "WalkerRing", "WalkerSub", "WalkerRows", "spmm_rowgroup_kernel", "6WalkerILi", "csr2csc", "max_row"):
        if key in name:
            kern = "long" if "spmm_long_kernel" in name else ("flat" if "spmm_flat_kernel" in name else "")
            return (key.replace("6WalkerILi", "Walker (register)") + (" / " + kern if kern else "")).strip()
    return "other"


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    fam, hist, kernels = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fam = family(m.group(1))
            kernels[fam] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and fam:
            op = m.group(1)
            hist[fam][next((w for w in PREFIXED if op.startswith(w)), op)] += 1
    print("# SASS opcode counts of %s (sm_100a), by kernel family; scripts/sass_histogram.py" % os.path.relpath(LIB, ROOT))
    print("# %-28s %8s %9s  %s" % ("family", "kernels", "instrs", "  ".join("%s" % w for w in WATCH)))
    for f in sorted(hist):
        h = hist[f]
        print("%-30s %8d %9d  %s" % (f, kernels[f], sum(h.values()), "  ".join("%s=%d" % (w, h[w]) for w in WATCH if h[w])))
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("%-30s %8d %9d  %s" % ("TOTAL", sum(kernels.values()), sum(tot.values()), "  ".join("%s=%d" % (w, tot[w]) for w in WATCH if tot[w])))
    print("# no HMMA / UTC*MMA / UTMALDG: the path has no dense tile (DESIGN.md 3.1); UBLKCP = cp.async.bulk (TMA 1-D), "
          "SYNCS = mbarrier arrive / try_wait, LDGSTS = cp.async, UCGABAR = thread-block-cluster barrier (long-row kernel)")


if __name__ == "__main__":
    main()
