"""SASS evidence for profiles/: per kernel of libsphb200.so, the counts of the instructions the design
rests on -- packed FP32 (FFMA2 / FADD2 / FMUL2), asynchronous copies (UBLKCP = cp.async.bulk / TMA,
LDGSTS = cp.async, SYNCS = mbarrier), texture fetches, 128-bit loads.

    python tools/sass_counts.py > profiles/r02_sass_counts.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "smoothed_particle_hydrodynamics_b200", "libsphb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WANT = ["FFMA2", "FADD2", "FMUL2", "FFMA", "UBLKCP", "LDGSTS", "SYNCS", "TLD", "TEX", "LDG.E.128", "LDS.128", "STS.128",
        "STG.E.128", "MUFU", "BAR", "ATOM", "RED"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op == w or op.startswith(w + ".") or (("." in w) and op.startswith(w)):
                counts[cur][w] += 1
print("# cuobjdump -sass smoothed_particle_hydrodynamics_b200/libsphb200.so (sm_100a): static instruction counts per kernel")
print("# FFMA2/FADD2/FMUL2 = fma/add/mul.rn.f32x2; UBLKCP = cp.async.bulk (TMA bulk copy); SYNCS = mbarrier ops")
print("%-46s %6s  %s" % ("kernel", "total", "selected opcodes"))
for k, c in counts.items():
    if not k.startswith("k_"):
        continue
    print("%-46s %6d  %s" % (k[:46], total[k], " ".join("%s=%d" % (w, c[w]) for w in WANT if c[w])))
