#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one bench step = the launches between two
stem_fused_kernel launches (or launches of the kernel named by the second argument, e.g. im2col_stem for the training
step); per kernel name: launches, total us, share of the step.  usage: ncu_summary.py launches.csv [delimiter kernel]"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
        rows.append((r["Kernel Name"], us))
    delim = sys.argv[2] if len(sys.argv) > 2 else "stem_fused_kernel"
    stems = [i for i, (n, _) in enumerate(rows) if delim in n]
    if len(stems) >= 2:
        rows = rows[stems[-2]:stems[-1]]
    agg = collections.OrderedDict()
    for n, us in rows:
        n = re.sub(r"\(.*$", "", n)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches in the step, {total:.1f} us serialised (cold-cache: compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'us':>9s} {'share':>7s}")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:8d} {us:9.1f} {us / total:7.3f}")


if __name__ == "__main__":
    main()
