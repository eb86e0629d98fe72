"""Per-source-line totals of an `ncu --page source --csv --print-source cuda,sass` export:
share of stall samples, warp instructions, average active threads.  usage: ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ''
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or r[0] in ('Function Name',) or not r[0].isdigit():
        continue
    try:
        samples = float(r[hdr['# Samples']] or 0); inst = float(r[hdr['Instructions Executed']] or 0); tinst = float(r[hdr['Thread Instructions Executed']] or 0)
    except (ValueError, IndexError):
        continue
    out.append((samples, inst, tinst, fname, r[0], r[1].strip()))
ts = sum(o[0] for o in out) or 1; ti = sum(o[1] for o in out) or 1
print(f"total samples {ts:.0f}  warp instructions {ti:.0f}")
print("samples%  inst%   thr  file:line  source")
for s, i, t, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*s/ts:6.2f}  {100*i/ti:6.2f}  {t/max(i,1):4.1f}  {f}:{ln}  {src[:110]}")
