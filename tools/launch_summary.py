"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv <cmd>`):
launches, total and mean device time and share per kernel, for the whole run and for each of its parts. A part starts
at every big `k1_phase1` launch after a march (a world being loaded: the bench runs config 3, then 4, then 5).

    python tools/launch_summary.py profiles/r02b_launches.csv
"""
import csv
import sys
from collections import OrderedDict


def short(name):
    name = name.split("(")[0]
    for p in ("void ", "stg::", "(anonymous namespace)::", "<unnamed>::", "unnamed>::"):
        name = name.replace(p, "")
    return name.strip()[:60]


def table(rows, title):
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("\n**%s** (%d launches, %.1f ms of device time under ncu)\n" % (title, len(rows), tot / 1e3))
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot >= 0.002:
            print("| `%s` | %d | %.1f | %.1f | %.3f |" % (k, a[0], a[1], a[1] / a[0], a[1] / tot))


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v
        rows.append((short(r[ik]), us))
    table(rows, "whole run")
    parts, cur, seen_march = [], [], False
    for k, us in rows:
        if k.startswith("k1_phase1") and us > 200 and seen_march:
            parts.append(cur)
            cur, seen_march = [], False
        if k.startswith("k2_ray_march"):
            seen_march = True
        cur.append((k, us))
    parts.append(cur)
    if len(parts) > 1:
        for i, p in enumerate(parts):
            table(p, "part %d" % (i + 1))


if __name__ == "__main__":
    main()
