"""Top stalled SASS instructions + per-opcode stall samples from an .ncu-rep source page (first kernel)."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == 'Address':
        if data: break   # next kernel
        continue
    data.append(r)
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data)
print('instructions', len(data), 'total samples', tot)
for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
    s = {h: int(r[hdr.index(h)]) for h in stalls if int(r[hdr.index(h)]) > 0}
    dom = sorted(s.items(), key=lambda kv: -kv[1])[:3]
    print(r[isamp].rjust(6), r[iex].rjust(8), r[0][-5:], r[isrc].strip()[:56].ljust(56), dom)
c = Counter(); e = Counter(); st = Counter()
for r in data:
    parts = r[isrc].split()
    op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    c[op] += int(r[isamp]); e[op] += int(r[iex])
    for h in stalls: st[h] += int(r[hdr.index(h)])
print('--- by opcode: samples, executed')
for op, v in c.most_common(16): print(op.ljust(10), v, e[op])
print('--- stall totals')
for h, v in st.most_common(10): print(h, v)
