"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, sys, collections
tot = collections.defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as fh:
    rows = [r for r in csv.reader(fh) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    if r[iu] == "us": v *= 1e3
    elif r[iu] == "ms": v *= 1e6
    name = r[ik].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += v
allt = sum(v[1] for v in tot.values())
print("total %.3f ms over %d launches" % (allt / 1e6, sum(v[0] for v in tot.values())))
for name, (cnt, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print("%7.2f%% %10.3f ms %7d x %9.2f us  %s" % (100 * t / allt, t / 1e6, cnt, t / cnt / 1e3, name))
