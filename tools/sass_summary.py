#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libxped_b200.so (run here, no GPU needed):  python tools/sass_summary.py > profiles/r02_sass_summary.txt
DMMA = FP64 tensor core (mma.sync m8n8k4), UBLKCP = cp.async.bulk (1-D TMA), LDGSTS = cp.async, UCGABAR / CCTL... = cluster barriers, SYNCS = mbarrier ops;
UTMALDG / UTMASTG (tensor-map TMA) and UTCxMMA / LDTM / STTM (tcgen05 + TMEM) are listed to show that they are absent: FP64 has no tcgen05 kind."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "xped_b200/lib/libxped_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["DMMA", "DFMA", "UBLKCP", "LDGSTS", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "SYNCS", "UCGABAR", "BAR.SYNC", "SHFL", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "MEMBAR"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["instr"] += 1
    for k in KEYS:
        if op.startswith(k):
            counts[cur][k] += 1
            total[k] += 1
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0][-70:]
    except Exception:
        return n[-70:]
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
for n, c in counts.items():
    print(f"{demangle(n):70s} {c['instr']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
print(f"{'TOTAL':70s} {sum(c['instr'] for c in counts.values()):7d} " + " ".join(f"{total[k]:8d}" for k in KEYS))
