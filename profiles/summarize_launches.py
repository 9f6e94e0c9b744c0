"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: python profiles/summarize_launches.py launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= mi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mi].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    print("%-62s %8s %12s %8s %10s" % ("kernel", "launches", "total_us", "share", "avg_us"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %8d %12.1f %7.2f%% %10.1f" % (k[:62], n, t, 100 * t / tot, t / n))
    print("%-62s %8d %12.1f" % ("total", sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
