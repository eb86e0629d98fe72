"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`)."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; idx = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[idx['Metric Value']].replace(',', ''))
    except Exception:
        continue
    n = r[idx['Kernel Name']].split('(')[0][:72]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:74s} {c:6d} {t/1e3:11.1f} us {100*t/tot:5.1f}%")
print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
