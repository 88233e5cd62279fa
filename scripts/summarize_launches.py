#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(kernel, launches, total ms, share).  usage: summarize_launches.py launches.csv "title" "command" > profiles/x.md"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(rows)
    agg = OrderedDict()
    n = 0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        elif r["Metric Unit"] == "ms":
            ns *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("# %s\n" % title)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` of `%s` (cold-cache, serialised: compare SHARES, not "
          "absolute times).  Every launch of the process.\n" % cmd)
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (name[:110], c, ns * 1e-6, 100 * ns / tot))
    print("\nTotal %.1f ms over %d launches." % (tot * 1e-6, n))


if __name__ == "__main__":
    main()
