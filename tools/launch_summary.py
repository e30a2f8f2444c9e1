"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the conv_tc sequence."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        us = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        rows.append((x["Kernel Name"].split("(")[0].replace("void ", ""), x["Grid Size"], us))
    return rows


if __name__ == "__main__":
    rows = load(sys.argv[1])
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for n, g, us in rows:
        tot[n] += us
        cnt[n] += 1
    for n in sorted(tot, key=lambda k: -tot[k]):
        print(f"{n:40s} {cnt[n]:4d} launches {tot[n] / 1e3:9.3f} ms")
    if len(sys.argv) > 2:
        seq = [us for n, g, us in rows if n.startswith(sys.argv[2])]
        print(" ".join(f"{u:.0f}" for u in seq))
