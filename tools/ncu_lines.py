"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = {}
        for h, v in zip(hdr, r):
            d.setdefault(h, v)
        try:
            s = int(d["# Samples"] or 0)
        except ValueError:
            s = 0
        def g(k):
            try: return int(d.get(k, 0) or 0)
            except ValueError: return 0
        lines.append((s, cur.split("/")[-1], d["Line No"], r[1].strip()[:90], g("stall_barrier"), g("stall_short_sb"), g("stall_long_sb"), g("stall_math"), g("Instructions Executed"), g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Ideal")))
tot = sum(l[0] for l in lines)
print("total samples", tot)
byfile = {}
for l in lines:
    byfile[l[1]] = byfile.get(l[1], 0) + l[0]
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items()})
lines.sort(key=lambda x: -x[0])
print("%7s %-18s %5s  %6s %6s %6s %6s %10s %9s/%-9s  %s" % ("smp%", "file", "line", "barr", "shsb", "lgsb", "math", "inst", "shWF", "ideal", "source"))
for s, f, ln, src, b, ss, ls, m, ie, wf, wfi in lines[:top]:
    print("%6.2f%% %-18s %5s  %6d %6d %6d %6d %10d %9d/%-9d  %s" % (100 * s / tot, f, ln, b, ss, ls, m, ie, wf, wfi, src))
