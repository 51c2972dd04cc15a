"""Correlates an ncu SASS-level sampling export (--page source --csv) with CUDA source lines via nvdisasm -g line info.
usage: sass_lines.py <all.sass from nvdisasm -g -c> <kernel mangled-name substring> <ncu csv>"""
import csv, re, sys, collections
sass, kname, ncsv = sys.argv[1:4]
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
cur = ("?", 0)
instrs = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        instrs.append((cur, m.group(2)))
rows = list(csv.reader(open(ncsv)))
hdr = rows[1]
si = hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
n = len(instrs)
data = data[:n]  # first launch only
assert len(data) == n, (len(data), n)
agg = collections.Counter()
tot = 0
for (loc, _), r in zip(instrs, data):
    agg[loc] += int(r[si]); tot += int(r[si])
print("total samples", tot, "instructions", n)
for loc, v in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    print(f"{100*v/tot:5.1f}%  {loc[0]}:{loc[1]}")
