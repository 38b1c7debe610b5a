"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last N launches (= one bench step) grouped
by kernel with their share of the step.  python tools/summarize_launches.py launches.csv 23 > profiles/..._step.csv"""
import collections
import csv
import re
import sys


def main(path, per_step):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    step = [(re.sub(r"\(.*", "", r[ki]).strip(), float(r[vi].replace(",", ""))) for r in rows[-per_step:]]
    agg = collections.OrderedDict()
    for n, v in step:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v for _, v in step)
    print("# last %d launches of %s = one step (cold-cache, serialised); total %.3f ms" % (per_step, path, tot / 1e6))
    print("kernel,launches,time_ns,share")
    for n, (c, v) in agg.items():
        print('"%s",%d,%.0f,%.3f' % (n, c, v, v / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
