"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total us, share).

    python tests/tools/launch_summary.py launches.csv            # every launch in the list
    python tests/tools/launch_summary.py launches.csv --step     # ONE alternating step: the launches between the GS kernel
                                                                 # of the last-but-one step and that of the last step
GEMM share = the tcgen05 kernels (conv16_persistent, conv_gemm, conv_strip16, stem_s2d_fprop / wgrad, linear_gemm)."""
import collections
import csv
import re
import sys

GEMM = ("conv16_persistent", "conv16_pair", "conv_gemm_kernel", "conv_strip16", "stem_s2d_fprop", "stem_s2d_wgrad_kernel",
        "linear_gemm")


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, out = None, []
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
            out.append((name, v))
    return out


def main(path, step=False):
    launches = load(path)
    if step:
        gs = [i for i, (n, _) in enumerate(launches) if n.startswith("gs_project")]
        if len(gs) < 3:
            raise SystemExit("fewer than three gs_project launches in the list: cannot isolate a step")
        launches = launches[gs[-3] + 1:gs[-1] + 1]          # two modality turns = one alternating step
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v in launches:
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-66s %6s %12s %7s" % ("kernel", "n", "total us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s %6d %12.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-66s %6d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
    gemm = sum(v[1] for k, v in agg.items() if k.startswith(GEMM))
    print("tcgen05 GEMM kernels: %.1f us = %.1f %% of the serialised launches; everything else %.1f %%" % (
        gemm, 100 * gemm / tot, 100 - 100 * gemm / tot))


if __name__ == "__main__":
    main(sys.argv[1], "--step" in sys.argv)
