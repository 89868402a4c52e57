#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total device time, share.
Usage: tools/launch_shares.py launches.csv [last_n_launches]   (last_n: only the tail of the list, e.g. one PPO iteration)"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit, val = r[ci["Metric Unit"]], float(r[ci["Metric Value"]].replace(",", ""))
        us = val * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
        rows.append((name, us))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    tot = sum(u for _, u in rows)
    agg = {}
    for n, u in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += u
    print(f"total {tot:.0f} us, {len(rows)} launches")
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:64]:64s} {c:6d} {u:10.1f} us {100 * u / tot:6.2f}%  avg {u / c:8.1f}")


if __name__ == "__main__":
    main()
