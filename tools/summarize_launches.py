#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel.
usage: tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys


def main(path):
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
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
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
        key = (d["Kernel Name"][:110], d["Grid Size"], d["Block Size"])
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'total ms':>10} {'n':>5} {'avg us':>10} {'share':>7}  kernel  grid block")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e6:10.3f} {v[0]:5d} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / tot:6.1f}%  {k[0]}  {k[1]} {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1])
