"""Top stalled SASS instructions of an `ncu --page source --csv` export: python tools/top_stalls.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "ninstr", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:N]
for i in sorted(top):
    r = data[i]
    stalls = {h: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and '(Not' not in h and int(r[ix[h]]) > 0}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(i, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:80], main)
