#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0][-60:]
            v = float(d["Metric Value"].replace(",", ""))
            unit = d["Metric Unit"]
            if unit in ("nsecond", "ns"):
                v /= 1e3
            elif unit in ("msecond", "ms"):
                v *= 1e3
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time from %s (ncu: cold-cache, serialised; compare SHARES)" % path)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s n=%4d total=%9.1f us avg=%8.1f us share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
