"""Instruction count per source line (and per inlined call site) of a JIT-dumped kernel: nvdisasm -g output -> histogram.
   python tools/sass_lines.py /tmp/prefix   (needs /tmp/prefix.cu and /tmp/prefix.cubin from KQ_JIT_DUMP)"""
import re, subprocess, sys, collections
prefix = sys.argv[1]
src = open(prefix + ".cu").read().splitlines()
dis = subprocess.run(["nvdisasm", "-g", "-c", prefix + ".cubin"], capture_output=True, text=True).stdout.splitlines()
cur = None
cnt = collections.Counter()
ops = collections.defaultdict(collections.Counter)
infunc = False
want = sys.argv[2] if len(sys.argv) > 2 else "kq_group_aggregate"
for l in dis:
    if l.startswith(".text."):
        infunc = l.startswith(".text." + want)
    if not infunc:
        continue
    m = re.match(r'\s*//## File "[^"]*", line (\d+)(.*)', l)
    if m:
        # outermost call site when inlined: 'inlined at "file", line N' chains; keep the innermost line + the outermost
        chain = [int(m.group(1))] + [int(x) for x in re.findall(r'line (\d+)', m.group(2))]
        cur = tuple(chain)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)', l)
    if m and cur:
        cnt[cur] += 1
        ops[cur][m.group(2)] += 1
# aggregate by OUTERMOST line (position in the kernel body)
by_outer = collections.Counter()
for k, v in cnt.items():
    by_outer[k[-1]] += v
tot = sum(cnt.values())
print("total instructions:", tot)
for line, v in sorted(by_outer.items()):
    print(f"{v:6d}  L{line}: {src[line-1].strip()[:150]}")
