"""Top stall lines of an `ncu --page source --csv` export (SASS view), first kernel section only."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hdr[0]]
end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
ix = {n: i for i, n in enumerate(h)}
data = []
for r in rows[hdr[0] + 1:end]:
    if len(r) < len(h):
        continue
    s = float(r[ix['# Samples']] or 0)
    data.append((s, r[ix['Source']].strip(), r[ix['stall_long_sb']], r[ix['stall_lg']], r[ix['stall_short_sb']], r[ix['stall_barrier']], r[ix['Instructions Executed']], r[ix['Avg. Threads Executed']]))
tot = sum(d[0] for d in data) or 1
print(rows[0][1][:100])
print("share  long_sb lg short barrier  inst_exec avg_thr  SASS")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/tot:5.3f}  {d[2]:>6s} {d[3]:>4s} {d[4]:>5s} {d[5]:>5s} {d[6]:>10s} {d[7]:>5s}   {d[1][:90]}")
