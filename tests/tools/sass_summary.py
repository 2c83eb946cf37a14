"""Blackwell-nativeness evidence from the shipped library: per kernel, how many tcgen05 / TMEM / TMA instructions its SASS
holds (cuobjdump -sass libmla_b200.so).   python tests/tools/sass_summary.py > profiles/r2_sass_summary.txt

  UTCHMMA / UTCQMMA...  tcgen05.mma (.2CTA = cta_group::2)      LDTM / STTM   tcgen05.ld / st (TMEM <-> registers)
  UTMALDG               TMA tensor load (.IM2COL variants)      UTMASTG       TMA tensor store      UTMAREDG  TMA reduce-add
  UBLKCP                cp.async.bulk (1-D bulk copy)           UTCBAR        tcgen05.commit -> mbarrier
  HMMA                  warp-level mma.sync (NOT tcgen05)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "multimodal-learning-with-alternating-unimodal-adaptation_b200", "libmla_b200.so")
MNEM = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.IM2COL", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    cur, k = None, -1
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            k += 1
            nm = names[k] if k < len(names) else m.group(1)
            nm = re.sub(r"\((?:int|bool|unsigned int)\)", "", nm.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
            nm = re.sub(r"\(.*", "", nm).replace("void ", "")
            cur = counts.setdefault(nm, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        if op.startswith("UTCHMMA") or op.startswith("UTCQMMA") or op.startswith("UTCIMMA") or op.startswith("UTCOMMA"):
            cur["UTCHMMA"] += 1
            if ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
        elif op.startswith("LDTM"):
            cur["LDTM"] += 1
        elif op.startswith("STTM"):
            cur["STTM"] += 1
        elif op.startswith("UTMALDG"):
            cur["UTMALDG"] += 1
            if "IM2COL" in op:
                cur["UTMALDG.IM2COL"] += 1
        elif op.startswith("UTMASTG"):
            cur["UTMASTG"] += 1
        elif op.startswith("UTMAREDG"):
            cur["UTMAREDG"] += 1
        elif op.startswith("UBLKCP"):
            cur["UBLKCP"] += 1
        elif op.startswith("UTCBAR"):
            cur["UTCBAR"] += 1
        elif op.startswith("HMMA"):
            cur["HMMA"] += 1
    print(__doc__.split("\n\n")[1] if "\n\n" in __doc__ else "")
    print("%-64s" % "kernel" + "".join("%9s" % m.replace("UTMALDG.IM2COL", ".IM2COL").replace("UTCHMMA.2CTA", ".2CTA") for m in MNEM))
    tot = collections.Counter()
    for nm, c in counts.items():
        if not any(c[m] for m in MNEM):
            continue
        print("%-64s" % nm[:64] + "".join("%9d" % c[m] for m in MNEM))
        tot.update(c)
    print("%-64s" % "TOTAL (whole library)" + "".join("%9d" % tot[m] for m in MNEM))
    print("kernels in the library: %d; kernels with tcgen05.mma: %d" % (len(counts), sum(1 for c in counts.values() if c["UTCHMMA"])))


if __name__ == "__main__":
    sys.exit(main())
