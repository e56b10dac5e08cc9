#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST
step (steps are delimited by q_sample_kernel launches) and, with --detail, every launch of that step.

  python scripts/ncu_launch_summary.py gpurun_out/r20_launches.csv [--detail] [--md]
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("b200dm::", "").replace("__nv_bfloat16", "bf16")
    return name[:70]


def main():
    path = sys.argv[1]
    detail = "--detail" in sys.argv
    md = "--md" in sys.argv
    rows, hdr = [], None
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            ns = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif d["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        rows.append((short(d["Kernel Name"]), d["Grid Size"], d["Block Size"], ns))
    starts = [i for i, r in enumerate(rows) if r[0].startswith("q_sample_kernel")]
    if len(starts) >= 2:
        lo, hi = starts[-2], starts[-1]       # a complete warm step (the last one may be cut by -c)
        if len(rows) - starts[-1] > (starts[-1] - starts[-2]) * 0.95:
            lo, hi = starts[-1], len(rows)
    else:
        lo, hi = 0, len(rows)
    step = rows[lo:hi]
    agg = collections.OrderedDict()
    for n, g, b, ns in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# one step: {len(step)} launches, {tot / 1e3:.1f} us summed device time (serialised, cold-cache under ncu)")
    if md:
        print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print(f"| `{n}` | {c} | {ns / 1e3:.1f} | {ns / c / 1e3:.1f} | {ns / tot * 100:.1f}% |")
        else:
            print(f"{n:70s} {c:4d} {ns / 1e3:10.1f} {ns / c / 1e3:8.1f} {ns / tot * 100:5.1f}%")
    if detail:
        print()
        for i, (n, g, b, ns) in enumerate(step):
            print(f"{i:4d} {n:60s} grid {g:18s} block {b:14s} {ns / 1e3:8.1f} us")


if __name__ == "__main__":
    main()
