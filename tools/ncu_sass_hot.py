"""Top stalled SASS instructions of an `ncu --page source --csv --print-source sass` export, with the two main stall
reasons of each and the kernel-wide stall-reason shares.  usage: ncu_sass_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if r and r[0] == 'Address')
idx = {n: i for i, n in enumerate(hdr)}
def num(r, n):
    try:
        return float(r[idx[n]] or 0)
    except (ValueError, IndexError):
        return 0.0
data = [r for r in rows if len(r) > 40 and r[0].startswith('0x')]
tot = sum(num(r, '# Samples') for r in data) or 1
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
print('total samples', int(tot), ' instructions', len(data))
agg = {n: sum(num(r, n) for r in data) for n in stalls}
for n, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f'  {n:26s} {100 * v / tot:6.2f}%')
for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:top_n]:
    s = num(r, '# Samples')
    why = sorted(((num(r, n), n) for n in stalls), reverse=True)[:2]
    print(f"{100 * s / tot:5.2f}%  {r[idx['Source']].strip()[:72]:72s} thr {r[idx['Avg. Threads Executed']]:>5s}  "
          + ', '.join(f'{n[6:]}={100 * v / max(s, 1):.0f}%' for v, n in why))
