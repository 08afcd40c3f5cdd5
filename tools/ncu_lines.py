"""Attribute an ncu SASS source page (ncu -i rep --page source --csv) to CUDA source lines using nvdisasm -g of the same kernel compiled
here (KQ_JIT_DUMP prefix): python tools/ncu_lines.py <src.csv> <dump prefix> [kernel] [N]"""
import csv, re, subprocess, sys, collections
csvp, prefix = sys.argv[1], sys.argv[2]
kernel = sys.argv[3] if len(sys.argv) > 3 else "kq_group_aggregate"
N = int(sys.argv[4]) if len(sys.argv) > 4 else 45
src = open(prefix + ".cu").read().splitlines()
dis = subprocess.run(["nvdisasm", "-g", "-c", prefix + ".cubin"], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur = None; infunc = False
for l in dis:
    if l.startswith(".text."): infunc = l.startswith(".text." + kernel)
    if not infunc: continue
    m = re.match(r'\s*//## File "[^"]*", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+', l)
    if m and cur: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvp)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = None
inst = collections.Counter(); samp = collections.Counter(); wf = collections.Counter()
stalls = collections.defaultdict(collections.Counter)
scols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[0], 16)
    if base is None: base = a
    ln = line_of.get(a - base, 0)
    inst[ln] += float(r[ix["Instructions Executed"]]); samp[ln] += float(r[ix["# Samples"]]); wf[ln] += float(r[ix["L1 Wavefronts Shared"]] or 0)
    for c in scols: stalls[ln][c] += float(r[ix[c]] or 0)
T, S, W = sum(inst.values()), sum(samp.values()), sum(wf.values())
print(f"warp instructions {T:.4g}, samples {S:.0f}, shared wavefronts {W:.4g}")
for ln, v in sorted(samp.items(), key=lambda kv: -kv[1])[:N]:
    top = ", ".join(f"{k[6:]} {int(x)}" for k, x in stalls[ln].most_common(2))
    print(f"{100*v/S:5.1f}% samp {100*inst[ln]/T:5.1f}% inst {100*wf[ln]/max(W,1):5.1f}% wf  L{ln}: {src[ln-1].strip()[:95] if ln else '?'}   [{top}]")
