#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (markdown table)."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    out = []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                u = d["Metric Unit"]
                v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
                out.append((d["Kernel Name"], v))
    return out


def main():
    launches = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, ms in launches:
        k = re.sub(r"\(.*", "", name).replace("void ", "")[:80]
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.2f ms total\n" % (len(launches), tot))
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    main()
