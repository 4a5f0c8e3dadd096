#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
usage: python scripts/launch_summary.py gpurun_out/launches_TAG.csv [> profiles/launches_TAG.md]"""
import collections
import csv
import statistics
import sys


def main(fn):
    hdr = None
    agg = collections.defaultdict(list)
    grids = {}
    for r in csv.reader(open(fn)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d.get("Metric Unit") == "us":
            v *= 1e3
        k = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        agg[k].append(v)
        grids[k] = (d["Grid Size"], d["Block Size"])
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total us | share | avg us | median us | max us | last grid | block |")
    print("|---|---:|---:|---:|---:|---:|---:|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f | %.3f | %.1f | %.1f | %.1f | %s | %s |" % (
            k, len(v), sum(v) / 1e3, sum(v) / tot, sum(v) / len(v) / 1e3, statistics.median(v) / 1e3, max(v) / 1e3,
            grids[k][0], grids[k][1]))
    print("\ntotal %.1f us over %d launches (gpu__time_duration.sum, cold-cache serialised replay: shares only)" % (
        tot / 1e3, sum(len(v) for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
