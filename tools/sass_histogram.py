"""SASS opcode histogram of the shipped library, per kernel: `python tools/sass_histogram.py > profiles/r2_sass_histogram.txt`.
Counts the mnemonics that identify the hardware paths used (DMMA = FP64 tensor core mma.sync.m8n8k4, DFMA/DADD/DMUL = FP64
vector pipe, UBLKCP = cp.async.bulk (TMA engine, non-tensor), UTMALDG = tensor-map TMA, LDGSTS = cp.async, SYNCS = mbarrier,
REDUX/CREDUX = warp reductions, MUFU = special function unit)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gcdyn.jl_b200", "lib", "libgcdyn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = ("DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "REDUX", "CREDUX", "MUFU", "LDS", "STS",
         "LDG", "STG", "ATOM", "RED", "BAR", "ACQBULK", "UTC", "LDTM")
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        hist[cur]["_total"] += 1
        for n in names:
            if op == n or (n in ("UTC", "LDTM") and op.startswith(n)):
                hist[cur][n] += 1
print("kernel".ljust(60), "total", " ".join(n.rjust(7) for n in names))
for k, h in hist.items():
    if h["_total"] < 40:
        continue
    print(k[:60].ljust(60), str(h["_total"]).rjust(5), " ".join(str(h[n]).rjust(7) for n in names))
