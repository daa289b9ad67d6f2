"""Per-CUDA-source-line executed-instruction and stall-sample shares from
`ncu --page source --print-source cuda,sass --csv` output."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst = collections.Counter(); stall = collections.Counter(); text = {}
cur_file = None
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; li = 0; si = 1; ci = hdr.index("Instructions Executed"); wi = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():   # source line row (no SASS address)
        key = (cur_file, int(r[0]))
        try:
            inst[key] += int(r[ci]); stall[key] += int(r[wi])
        except ValueError:
            pass
        text[key] = r[1].strip()
tot = sum(inst.values()) or 1; ts = sum(stall.values()) or 1
print("total", tot)
for key, c in inst.most_common(top):
    print("%5.2f%% stall %5.2f%%  %s:%d  %s" % (100*c/tot, 100*stall[key]/ts, key[0], key[1], text[key][:90]))
