"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per kernel: launches,
average and total duration, share of the listed GPU time)."""
import collections
import csv
import sys


def table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name, unit, val = r[4], r[13], float(r[14].replace(",", ""))
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    out = ["| launches | avg us | total ms | share | kernel |", "|---|---|---|---|---|"]
    for name, (k, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {us / k:.2f} | {us * 1e-3:.3f} | {100 * us / total:.1f}% | `{name[:100]}` |")
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1]))
