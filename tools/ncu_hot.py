"""Top stall PCs of an ncu source-page CSV (ncu -i rep --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ai, xi = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [(int(r[si] or 0), int(r[xi] or 0), r[ai].strip(), i) for i, r in enumerate(rows[2:]) if len(r) > si and r[si].isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
# opcode histogram by samples
from collections import Counter
c = Counter(); ce = Counter()
for s, x, src, i in data:
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    c[op] += s; ce[op] += x
print("by opcode (samples%, executed):", [(k, round(100*v/tot,1), ce[k]) for k, v in c.most_common(14)])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, x, src, i in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:6.2f}%  exec={x:9d}  line {i:5d}  {src}")
