"""Per-source-line instruction counts, active lanes and stall samples from `ncu --page source --csv --print-source sass,cuda`
   (handles reports with several kernels).   python tools/src_hot2.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
if not starts:
    starts = [0]
for n, i in enumerate(starts):
    end = starts[n + 1] - 1 if n + 1 < len(starts) else len(rows)
    hi = i + 1
    while hi < end and not (rows[hi] and rows[hi][0] in ("Line No", "#", "Address")):
        hi += 1
    hdr = rows[hi]; ix = {}
    for j, h in enumerate(hdr): ix.setdefault(h, j)
    per = {}
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr) or not r[0]: continue
        try:
            ln = int(r[0]); a = per.setdefault(ln, [r[1].strip(), 0, 0, 0])
            a[1] += int(r[ix['# Samples']] or 0); a[2] += int(r[ix['Instructions Executed']] or 0); a[3] += int(r[ix['Thread Instructions Executed']] or 0)
        except ValueError:
            pass
    ti = sum(v[2] for v in per.values()); ts = sum(v[1] for v in per.values()); tt = sum(v[3] for v in per.values())
    print("==", rows[i][1] if len(rows[i]) > 1 else "", "warp-instructions", ti, "samples", ts, "avg active lanes %.1f" % (tt / max(ti, 1)))
    for ln, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:N]:
        print(f"{ln:5d} {100*v[2]/max(ti,1):5.1f}% inst {100*v[1]/max(ts,1):5.1f}% samp lanes {v[3]/max(v[2],1):4.1f}  {v[0][:110]}")
    print("-- by stall samples")
    for ln, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"{ln:5d} {100*v[2]/max(ti,1):5.1f}% inst {100*v[1]/max(ts,1):5.1f}% samp lanes {v[3]/max(v[2],1):4.1f}  {v[0][:110]}")
