"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).  Usage: python tools/launch_shares.py file.csv [--seq]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, seq, tot = collections.OrderedDict(), [], 0.0
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"void |\(anonymous namespace\)::|rgbid::|<unnamed>::", "", name)
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        seq.append((name, v))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v; tot += v
    print("%d launches, %.1f us" % (len(seq), tot))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %3d %9.1f us %5.1f%%  avg %6.1f" % (k[:58], n, t, 100 * t / tot, t / n))
    if "--seq" in sys.argv:
        print(" ".join("%s:%.0f" % (n[:12], v) for n, v in seq))


if __name__ == "__main__":
    main()
