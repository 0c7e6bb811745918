"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share.
usage: python tools/launch_summary.py profiles/r01o_launches.csv"""
import csv, re, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    k = re.sub(r"\(.*", "", r[ik])
    tot[k] += v
    cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common():
    print("%-44s n=%4d  total %10.1f us  avg %8.1f us  share %5.1f%%" % (k[:44], cnt[k], v, v / cnt[k], 100 * v / s))
