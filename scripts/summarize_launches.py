#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
count, total device time, share.  Usage: summarize_launches.py launches.csv [> profiles/...txt]"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip("usecond") if r[ui] in ("nsecond", "usecond", "msecond", "second") else r[ui], None)
        scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        name = r[ki]
        short = name.split("(")[0][-80:] if "spmm" in name or "gespmm" in name else name[:60]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("%6s %12s %7s  kernel   (device time per ncu, cold-cache and serialised: compare shares)" % ("count", "total_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%6d %12.1f %6.1f%%  %s" % (a[0], a[1], 100 * a[1] / tot, k))
    print("%6d %12.1f %6.1f%%  TOTAL" % (sum(a[0] for a in agg.values()), tot, 100.0))


if __name__ == "__main__":
    main(sys.argv[1])
