#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file <raw.csv> ...`).
usage: tools/launch_summary.py <raw.csv> <summary.csv> ["# comment line" ...]"""
import csv
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = {}, {}
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        tot[name] = tot.get(name, 0.0) + v
        cnt[name] = cnt.get(name, 0) + 1
    allt = sum(tot.values())
    with open(sys.argv[2], "w") as f:
        for c in sys.argv[3:]:
            f.write(c + "\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            f.write("%s,%d,%.1f,%.1f,%.3f\n" % (k, cnt[k], tot[k], tot[k] / cnt[k], tot[k] / allt))
        f.write("TOTAL,%d,%.1f,,1.000\n" % (sum(cnt.values()), allt))


if __name__ == "__main__":
    main()
