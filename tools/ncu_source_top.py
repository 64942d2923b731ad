#!/usr/bin/env python
"""Top instructions of an `ncu --page source --csv` dump by stall samples, plus shared-memory wavefront totals per opcode.
usage: tools/ncu_source_top.py <source.csv> [n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in data:
    for c in stall_cols: agg[c] += f(r, c)
print("stalls:", ", ".join("%s=%.1f%%" % (k[6:], 100 * v / max(1, sum(agg.values()))) for k, v in agg.most_common(8)))
op = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in data:
    o = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    o = o.split(".")[0] + ("." + ".".join(o.split(".")[1:3]) if o.startswith(("LD", "ST", "ATOM", "RED")) else "")
    op[o][0] += 1; op[o][1] += f(r, "# Samples"); op[o][2] += f(r, "L1 Wavefronts Shared"); op[o][3] += f(r, "L1 Wavefronts Shared Ideal")
print("by opcode: count samples%% shared_wavefronts ideal")
for o, v in sorted(op.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %-16s %4d %5.1f%% %12.0f %12.0f" % (o, v[0], 100 * v[1] / tot, v[2], v[3]))
print("top instructions:")
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:n]
for i in sorted(order):
    r = data[i]
    st = sorted(((f(r, c), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("  %4d %5.1f%% %-60s %s" % (i, 100 * f(r, "# Samples") / tot, r[ix["Source"]].strip()[:60], " ".join("%s=%d" % (b, a) for a, b in st)))
