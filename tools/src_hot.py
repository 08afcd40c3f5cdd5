"""Per-source-line instruction counts and stall samples from `ncu --page source --csv --print-source sass,cuda`.
   python tools/src_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
lines = []
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        lines.append((int(r[0]), r[1].strip(), int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0)))
    except ValueError:
        pass
ti = sum(l[3] for l in lines); ts = sum(l[2] for l in lines)
print("total warp-instructions", ti, "samples", ts)
print("--- by instructions executed")
for l in sorted(lines, key=lambda l: -l[3])[:N]:
    print(f"{l[0]:5d} {100*l[3]/ti:5.1f}% inst {100*l[2]/max(ts,1):5.1f}% samp  {l[1][:110]}")
print("--- by stall samples")
for l in sorted(lines, key=lambda l: -l[2])[:15]:
    print(f"{l[0]:5d} {100*l[3]/ti:5.1f}% inst {100*l[2]/max(ts,1):5.1f}% samp  {l[1][:110]}")
