"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total us, share)."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d["Metric Unit"]
            v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
            name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:64]
            agg[name][0] += 1
            agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-66s %6s %12s %7s" % ("kernel", "n", "total us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s %6d %12.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-66s %6d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
