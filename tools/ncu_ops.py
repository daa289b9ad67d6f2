import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
ci = hdr.index("Instructions Executed"); si = hdr.index("Source"); ti = hdr.index("Thread Instructions Executed")
sti = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ci]) for r in data)
print("total warp inst", tot, "sass lines", len(data))
ops = collections.Counter(); stalls = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(r[ci]); stalls[op] += int(r[sti])
ts = sum(stalls.values()) or 1
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s %6.2f%%  stall-samples %6.2f%%" % (op, 100*c/tot, 100*stalls[op]/ts))
